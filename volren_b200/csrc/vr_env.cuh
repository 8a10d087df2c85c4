// Environment map: bilinear REPEAT fetch, importance map (shader/env_setup.glsl:18-34,
// src/environment.cpp:11-33) and its box-filter pyramid (glGenerateMipmap), tonemapping
// (shader/tonemap.glsl, tonemap.fs, blit.fs).
#pragma once

#include "vr_common.cuh"

namespace vr {

constexpr int IMP_DIM = 512;    // environment.cpp:6
constexpr int IMP_LEVELS = 10;  // levels 0..9, env_imp_base_mip = 9 (renderer.cpp:131)

VR_HD constexpr uint32_t imp_offset(int level) {
    // sum_{l<level} (512>>l)^2
    uint32_t n = 0;
    for (int l = 0; l < level; ++l) n += uint32_t(IMP_DIM >> l) * uint32_t(IMP_DIM >> l);
    return n;
}

struct EnvView {
    const float4* rgb;   // w*h texels, bottom-up, .w unused
    int w, h;
    const float* impmap; // pyramid, levels concatenated
    const float4* split; // per 2x2 quad of every pyramid level: the split probabilities of sample_environment and their
                         // reciprocals (k_env_split), 3 x float4 per quad; nullptr: evaluate them per sample
};

// quads of level `mip` (its texels are (512 >> mip)^2) start at quad (4^(8 - mip) - 1) / 3
VR_HD constexpr uint32_t split_offset(int mip) { return ((1u << (2 * (8 - mip))) - 1u) / 3u; }
constexpr uint32_t SPLIT_QUADS = ((1u << 18) - 1u) / 3u;   // levels 8 ... 0: 1 + 4 + ... + 65536

VR_DEV int wrap_repeat(int i, int n) { int r = i % n; return r < 0 ? r + n : r; }

// texture(env_envmap, uv): GL_LINEAR, GL_REPEAT on both axes, LOD 0 (cppgl texture.cpp:27-60)
VR_DEV float3 env_texture(const EnvView& e, float u, float v) {
    const float x = u * float(e.w) - 0.5f, y = v * float(e.h) - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    const int x0 = wrap_repeat(int(fx), e.w), y0 = wrap_repeat(int(fy), e.h);
    const int x1 = x0 + 1 == e.w ? 0 : x0 + 1, y1 = y0 + 1 == e.h ? 0 : y0 + 1;
    const float4 t00 = __ldg(e.rgb + size_t(y0) * e.w + x0), t10 = __ldg(e.rgb + size_t(y0) * e.w + x1);
    const float4 t01 = __ldg(e.rgb + size_t(y1) * e.w + x0), t11 = __ldg(e.rgb + size_t(y1) * e.w + x1);
    return f3(mixf(mixf(t00.x, t10.x, ax), mixf(t01.x, t11.x, ax), ay),
              mixf(mixf(t00.y, t10.y, ax), mixf(t01.y, t11.y, ax), ay),
              mixf(mixf(t00.z, t10.z, ax), mixf(t01.z, t11.z, ax), ay));
}

VR_GLOBAL void k_env_pad(const float* __restrict__ rgb, float4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        out[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 0.f);
}

// env_setup.glsl:18-34 with num_samples = (8,8), output_size_samples = 4096, inv_samples = 1/64
VR_GLOBAL void __launch_bounds__(256) k_env_impmap(EnvView e, float* __restrict__ impmap) {
    const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (px >= IMP_DIM || py >= IMP_DIM) return;
    float importance = 0.f;
    for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) {
            const float u = (float(px * 8) + (float(x) + .5f)) / 4096.f;
            const float v = (float(py * 8) + (float(y) + .5f)) / 4096.f;
            importance += luma(env_texture(e, u, v));
        }
    impmap[size_t(py) * IMP_DIM + px] = importance * (1.f / 64.f);
}

// glGenerateMipmap on the R32F importance map: 2x2 box filter, pinned as 0.25f*((a+b)+(c+d)).
// One warp reduces a 2x2 quad per lane; levels are produced one launch per level (tiny, one-off).
VR_GLOBAL void k_env_mip(const float* __restrict__ src, int sdim, float* __restrict__ dst, int ddim) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= ddim || y >= ddim) return;
    const float2 r0 = *reinterpret_cast<const float2*>(src + size_t(2 * y) * sdim + 2 * x);
    const float2 r1 = *reinterpret_cast<const float2*>(src + size_t(2 * y + 1) * sdim + 2 * x);
    dst[size_t(y) * ddim + x] = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(r0.x, r0.y), __fadd_rn(r1.x, r1.y)));
}

// ------------------------------------------------------------------------------------------------
// tonemapping

VR_DEV float hable(float x) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
VR_DEV float hable_tonemap(float x, float exposure) { return hable(exposure * x) / hable(11.2f); }
VR_DEV uint32_t to_unorm8(float x) {  // GL float -> unorm8 conversion
    if (isnan(x)) return 0u;
    return uint32_t(__float2int_rn(saturate(x) * 255.f));
}

// shader/tonemap.glsl:29-36 (in place on the RGBA32F colour buffer)
VR_GLOBAL void k_tonemap_inplace(float4* __restrict__ color, size_t n, float exposure, float inv_gamma) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        float4 c = color[i];
        c.x = sanitize(powf(hable_tonemap(c.x, exposure), inv_gamma));
        c.y = sanitize(powf(hable_tonemap(c.y, exposure), inv_gamma));
        c.z = sanitize(powf(hable_tonemap(c.z, exposure), inv_gamma));
        c.w = sanitize(c.w);
        color[i] = c;
    }
}
// RendererOpenGL::draw (renderer.cpp:147-153): tonemap.fs or blit.fs into the RGBA8 framebuffer
VR_GLOBAL void k_draw(const float4* __restrict__ color, uchar4* __restrict__ fb, size_t n, float exposure, float inv_gamma, int tonemapping) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        float4 c = color[i];
        if (tonemapping) {
            c.x = powf(hable_tonemap(c.x, exposure), inv_gamma);
            c.y = powf(hable_tonemap(c.y, exposure), inv_gamma);
            c.z = powf(hable_tonemap(c.z, exposure), inv_gamma);
        }
        fb[i] = make_uchar4((unsigned char)to_unorm8(c.x), (unsigned char)to_unorm8(c.y), (unsigned char)to_unorm8(c.z), (unsigned char)to_unorm8(c.w));
    }
}
VR_GLOBAL void k_color_to_ldr(const float4* __restrict__ color, uchar4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const float4 c = color[i];
        out[i] = make_uchar4((unsigned char)to_unorm8(c.x), (unsigned char)to_unorm8(c.y), (unsigned char)to_unorm8(c.z), (unsigned char)to_unorm8(c.w));
    }
}
// packed tile coordinates (ty << 16 | tx) in raster order
VR_GLOBAL void k_tile_coords(uint32_t* __restrict__ out, size_t n, uint32_t tiles_x) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        out[i] = (uint32_t(i / tiles_x) << 16) | uint32_t(i % tiles_x);
}
VR_GLOBAL void k_scale(float4* __restrict__ color, size_t n, float s) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        float4 c = color[i];
        c.x *= s; c.y *= s; c.z *= s; c.w *= s;
        color[i] = c;
    }
}
VR_GLOBAL void k_add(float4* __restrict__ dst, const float4* __restrict__ src, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        float4 a = dst[i];
        const float4 b = src[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        dst[i] = a;
    }
}

}  // namespace vr
