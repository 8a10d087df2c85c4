// Shared device helpers for libvrb200: small vector math, the TEA/LCG random streams and the
// voldata number formats (fp16 round-half-up, 10/10/10 brick pointers, unorm8 voxels).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define VR_DEV __device__ __forceinline__
#define VR_HD __host__ __device__ __forceinline__
// Non-template kernels defined in headers: the strict translation unit (vrb200_strict.cu, compiled with -fmad=false for the
// IEEE cross-check kernels) includes the same headers, so there they get internal linkage and are dropped when unused.
#ifdef VR_STRICT_TU
#define VR_GLOBAL static __global__
#else
#define VR_GLOBAL __global__
#endif

namespace vr {

// ------------------------------------------------------------------------------------------------
// float3 helpers

VR_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
VR_HD float3 f3(float s) { return make_float3(s, s, s); }
VR_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
VR_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
VR_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
VR_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
VR_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
VR_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
VR_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
VR_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
VR_HD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
VR_DEV float3 normalize(float3 a) { return a * (1.f / sqrtf(dot(a, a))); }   // v * inversesqrt(dot(v, v)), as glm states GLSL's normalize
VR_HD float sqr(float x) { return x * x; }
VR_HD float luma(float3 c) { return dot(c, f3(0.212671f, 0.715160f, 0.072169f)); }            // common.glsl:21
VR_DEV float sanitize(float x) { return (isnan(x) || isinf(x)) ? 0.f : x; }                    // common.glsl:17
VR_DEV float saturate(float x) { return fminf(fmaxf(x, 0.f), 1.f); }                           // common.glsl:23
VR_DEV float mixf(float x, float y, float a) { return x * (1.f - a) + y * a; }                 // GLSL mix()
VR_DEV float power_heuristic(float a, float b) { return sqr(a) / (sqr(a) + sqr(b)); }          // common.glsl:35

// column-major 3x3 / 4x4 (glm layout)
struct Mat3 { float m[9]; };
struct Mat4 { float m[16]; };
VR_HD float3 mul(const Mat3& M, float3 v) {
    return f3(M.m[0] * v.x + M.m[3] * v.y + M.m[6] * v.z, M.m[1] * v.x + M.m[4] * v.y + M.m[7] * v.z,
              M.m[2] * v.x + M.m[5] * v.y + M.m[8] * v.z);
}
// mat4 * vec4(v, 1) and mat4 * vec4(v, 0): the four column products are summed pairwise, (c0 x + c1 y) + (c2 z + c3 w), the
// order of glm's operator* (glm/detail/type_mat4x4.inl) that the compiled-GLSL reference, which the compiled-GLSL reference of the test suite pins
VR_HD float3 mul_point(const Mat4& M, float3 v) {
    return f3((M.m[0] * v.x + M.m[4] * v.y) + (M.m[8] * v.z + M.m[12]), (M.m[1] * v.x + M.m[5] * v.y) + (M.m[9] * v.z + M.m[13]),
              (M.m[2] * v.x + M.m[6] * v.y) + (M.m[10] * v.z + M.m[14]));
}
VR_HD float3 mul_dir(const Mat4& M, float3 v) {
    // (c3 * 0 is kept: its signed zero decides the sign of an exactly-zero component, hence of 1 / idir)
    return f3((M.m[0] * v.x + M.m[4] * v.y) + (M.m[8] * v.z + M.m[12] * 0.f), (M.m[1] * v.x + M.m[5] * v.y) + (M.m[9] * v.z + M.m[13] * 0.f),
              (M.m[2] * v.x + M.m[6] * v.y) + (M.m[10] * v.z + M.m[14] * 0.f));
}

constexpr float PI_F = 3.14159265358979323846f;  // common.glsl:4
constexpr float INV_4PI = 1.f / (4 * PI_F);

// ------------------------------------------------------------------------------------------------
// RNG: TEA-32 seeding + 32-bit LCG stream (common.glsl:40-67)

VR_HD uint32_t tea32(uint32_t v0, uint32_t v1) {
    uint32_t s0 = 0;
#pragma unroll 4
    for (int n = 0; n < 32; ++n) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xA341316Cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4u);
        v1 += ((v0 << 4) + 0xAD90777Du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761Eu);
    }
    return v0;
}

constexpr uint32_t LCG_A = 1664525u, LCG_C = 1013904223u;
VR_HD float rng(uint32_t& s) {
    s = s * LCG_A + LCG_C;
    return float(s & 0x00FFFFFFu) * (1.f / 16777216.f);  // exact: both 24-bit int -> float and the 2^-24 scale
}
// closed-form k-step jump of the LCG: s -> A^k s + C (A^(k-1)+...+1)  (mod 2^32)
VR_HD constexpr uint32_t lcg_pow_a(int k) { uint32_t a = 1; for (int i = 0; i < k; ++i) a *= LCG_A; return a; }
VR_HD constexpr uint32_t lcg_sum_c(int k) { uint32_t c = 0; for (int i = 0; i < k; ++i) c = c * LCG_A + LCG_C; return c; }
template <int K> VR_HD void rng_skip(uint32_t& s) {
    constexpr uint32_t A = lcg_pow_a(K), C = lcg_sum_c(K);
    s = s * A + C;
}

// ------------------------------------------------------------------------------------------------
// voldata number formats

// glm::detail::toFloat16 (glm/detail/type_half.inl:105-238): round-half-UP on the magnitude.
VR_HD uint32_t float_to_half_rhu(float f) {
#ifdef __CUDA_ARCH__
    const int32_t bits = __float_as_int(f);
#else
    int32_t bits; memcpy(&bits, &f, 4);
#endif
    const int32_t sign = (bits >> 16) & 0x8000;
    int32_t e = ((bits >> 23) & 0xff) - 112;
    int32_t m = bits & 0x007fffff;
    if (e <= 0) {
        if (e < -10) return uint32_t(sign);
        m = (m | 0x00800000) >> (1 - e);
        if (m & 0x1000) m += 0x2000;
        return uint32_t(sign | (m >> 13));
    }
    if (e == 0xff - 112) {
        if (m == 0) return uint32_t(sign | 0x7c00);
        m >>= 13;
        return uint32_t(sign | 0x7c00 | m | (m == 0));
    }
    if (m & 0x1000) {
        m += 0x2000;
        if (m & 0x00800000) { m = 0; e += 1; }
    }
    if (e > 30) return uint32_t(sign | 0x7c00);
    return uint32_t(sign | (e << 10) | (m >> 13));
}

// encode_range (grid_brick.cpp:24-26). glm's hdata is a signed short, so uint32_t(hdata) sign-extends:
// a negative minimum ORs 0xffff into the majorant half. Reproduced bit for bit.
VR_HD uint32_t encode_range(float lo, float hi) {
    const uint32_t l = uint32_t(int32_t(int16_t(float_to_half_rhu(lo))));
    const uint32_t h = uint32_t(int32_t(int16_t(float_to_half_rhu(hi))));
    return l | (h << 16);
}
VR_DEV float half_bits_to_float(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)h)); }
VR_DEV float range_lo(uint32_t w) { return half_bits_to_float(w & 0xffffu); }
VR_DEV float range_hi(uint32_t w) { return half_bits_to_float(w >> 16); }

// encode_ptr / decode_ptr (grid_brick.cpp:32-43)
VR_HD uint32_t encode_ptr(uint32_t x, uint32_t y, uint32_t z) { return (x << 22) | (y << 12) | (z << 2); }
VR_HD uint3 decode_ptr(uint32_t d) { return make_uint3((d >> 22) & 1023u, (d >> 12) & 1023u, (d >> 2) & 1023u); }

}  // namespace vr
