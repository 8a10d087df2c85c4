// The tracking code: shader/pathtracer_brick.glsl + pathtracer_brick_tf.glsl + the parts of
// shader/common.glsl they reach (USE_DDA): brick-DDA majorant delta tracking, envmap NEE with MIS,
// Russian roulette, Henyey-Greenstein, LUT transfer function.
//
// Every helper is a template over a math policy MT:
//   StrictMath : IEEE division / sqrt and the accurate libdevice log / sincos -- the arithmetic of the CPU oracle
//                up to FMA contraction; used by the cross-check kernels (vrb_set_kernel 1 and 2);
//   FastMath   : MUFU-based rcp / rsqrt / lg2 / sin / cos (what a GLSL compiler emits for these built-ins on a GPU);
//                used by the production kernel. Besides speed this is about CODE SIZE: with IEEE sequences the
//                persistent kernel is 73 KB of SASS and stalls on instruction fetch (profiles/r01_v3_*).
// This file also holds the straightforward one-thread-per-pixel kernel (k_trace_pixels) and the deterministic
// transmittance-only mode; the production persistent kernel lives in vr_trace2.cuh.
#pragma once

#include "vr_common.cuh"
#include "vr_env.cuh"
#include "../../include/vrb200.h"

namespace vr {

#ifndef VR_DECODED
#define VR_DECODED 1        // 0: the production kernel fetches through the canonical records + u8 atlas as well (A/B builds)
#endif
struct StrictMath {
    static constexpr bool fast = false;
    static constexpr bool decoded = false;      // cross-check kernels: canonical records + u8 atlas
    static VR_DEV float div(float a, float b) { return a / b; }
    static VR_DEV float rcp(float a) { return 1.f / a; }
    static VR_DEV float sqrt(float a) { return sqrtf(a); }
    static VR_DEV float log(float a) { return logf(a); }
    static VR_DEV void sincos(float a, float* s, float* c) { sincosf(a, s, c); }
    static VR_DEV float3 normalize(float3 v) { return v * (1.f / sqrtf(dot(v, v))); }   // glm: v * inversesqrt(dot(v, v))
};
struct FastMath {
    static constexpr bool fast = true;
    static constexpr bool decoded = VR_DECODED != 0;
    static VR_DEV float rcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
    static VR_DEV float div(float a, float b) { return a * rcp(b); }
    static VR_DEV float sqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
    static VR_DEV float log(float a) { return __logf(a); }
    static VR_DEV void sincos(float a, float* s, float* c) { __sincosf(a, s, c); }
    static VR_DEV float3 normalize(float3 v) { return v * rsqrtf(dot(v, v)); }
};

struct GridView {
    uint3 nb;                  // n_bricks (level 0)
    const uint2* rec;          // {atlas slot, range word} per brick
    const uint32_t* mips[3];   // range words of levels 1..3 (dims nb >> level)
    const uint8_t* atlas_lin;  // slot * 512 + z*64 + y*8 + x  (+ one all-zero brick behind the last slot)
    const uint2* recp;         // records padded by one brick per side: ((bz+1) * (nby+2) + (by+1)) * (nbx+2) + (bx+1)
    uint32_t psx, psxy;        // strides of recp: nbx + 2, (nbx + 2) * (nby + 2)
    // decoded apron bricks (production trilinear fetch): cell (bx+1, by+1, bz+1) of the (nb+1)^3 lattice, bx in [-1, nb-1],
    // owns the 9^3 DECODED fp32 voxels 8b ... 8b+8 per axis, so that the 2x2x2 footprint of any sample point whose base
    // voxel lies in brick b is inside ONE block: datlas[cslot[cell] * 729 + z*81 + y*9 + x]; block 0 is all zeros and
    // shared by every cell whose 8 bricks are all (0, 0)-range or outside the grid
    const uint32_t* cslot;
    const float* datlas;
};
constexpr uint32_t DBRICK = 729u;

struct TraceArgs {
    vrb_params p;
    GridView density, emission;
    EnvView env;
    const float4* lut;
    uint32_t tf_size;
    float4* color;
    int x0, y0, x1, y1;
    int first_sample, n_samples, accum_mode;
    unsigned long long* counters;  // 7 x u64 (vrb_counters order) or nullptr
    Mat4 emis_from_density;        // vol_emission_inv_transform * vol_density_transform (common.glsl:325)
    float cam_z;                   // view_dir's z = -.5f / tan(.5f * M_PI * cam_fov / 180.f) (common.glsl:78), host libm
    // persistent kernel only
    const float* maj[4];           // per-level majorant tables (final value used by the tracking loop)
    const float* maj_oob;          // majorant of an out-of-bounds fetch (one float)
    uint32_t maj_off[4];           // the same tables as element offsets from maj[0] (they share one allocation); maj_off_oob: the out-of-bounds entry
    uint32_t maj_off_oob;
    unsigned int* job_counter;     // block ticket: block b = (tile slot b >> sample_bits, sample b & mask), 32 samples each
    int tiles_x, n_jobs;           // n_jobs = number of block ids = tiles << sample_bits (ids with sample >= n_samples are padding)
    int sample_bits;               // ceil(log2(n_samples))
    float4* lbuf;                  // per-launch sample buffer: lbuf[(s - first_sample) * lbuf_stride + y * W + x]
    size_t lbuf_stride;
    // tile slot k is the tile with packed coordinates tile_order[k] = (ty << 16 | tx): natural order, or heaviest first
    // when the previous launch of the same view left costs; finished samples add the cycles they occupied their lane to
    // tile_cost[ty * tiles_x + tx] (nullptr: not measured)
    const uint32_t* tile_order;
    unsigned int* tile_cost;
    // screen-space brick mask (hidden environment only): tile slots >= *n_live hold tiles no non-empty brick projects onto;
    // nullptr: every slot is live and n_jobs is the host's count
    const unsigned int* n_live;
};

template <bool COUNT> struct Cnt;
template <> struct Cnt<false> {
    VR_DEV void maj() {} VR_DEV void dens() {} VR_DEV void emis() {} VR_DEV void nee() {} VR_DEV void env() {} VR_DEV void real() {} VR_DEV void samp() {}
};
template <> struct Cnt<true> {
    uint32_t n_samp = 0, n_maj = 0, n_dens = 0, n_emis = 0, n_nee = 0, n_env = 0, n_real = 0;
    VR_DEV void maj() { ++n_maj; } VR_DEV void dens() { ++n_dens; } VR_DEV void emis() { ++n_emis; } VR_DEV void nee() { ++n_nee; }
    VR_DEV void env() { ++n_env; } VR_DEV void real() { ++n_real; } VR_DEV void samp() { ++n_samp; }
};
VR_DEV void flush_counters(const TraceArgs&, const Cnt<false>&) {}
VR_DEV void flush_counters(const TraceArgs& a, const Cnt<true>& c) {
    const uint32_t v[7] = { c.n_samp, c.n_maj, c.n_dens, c.n_emis, c.n_nee, c.n_env, c.n_real };
#pragma unroll
    for (int i = 0; i < 7; ++i) atomicAdd(a.counters + i, (unsigned long long)v[i]);
}

// exact u8 / 255.f (GL unorm8 -> float) without a divide: reciprocal multiply + one FMA correction step
VR_DEV float unorm8_to_float(uint32_t u) {
    const float x = float(u);
    const float r = 1.f / 255.f;
    const float q = x * r;
    const float rem = fmaf(-q, 255.f, x);
    return fmaf(rem, r, q);
}

// texelFetch chain of lookup_density_brick / lookup_temperature_brick (common.glsl:268-275, :314-321);
// out-of-bounds fetches return 0 (robust access), which makes the value 0 + 0 * (0 - 0).
VR_DEV float brick_value(const GridView& g, int x, int y, int z) {
    const int bx = x >> 3, by = y >> 3, bz = z >> 3;
    if (unsigned(bx) >= g.nb.x || unsigned(by) >= g.nb.y || unsigned(bz) >= g.nb.z) return 0.f;
    const uint2 r = __ldg(g.rec + (uint32_t(bz) * g.nb.y + uint32_t(by)) * g.nb.x + uint32_t(bx));   // n_bricks < 2^30: 32-bit index math
    const float lo = range_lo(r.y), hi = range_hi(r.y);
    float unorm = 0.f;
    if (r.x != 0xffffffffu)
        unorm = unorm8_to_float(__ldg(g.atlas_lin + size_t(r.x) * 512u + uint32_t(((z & 7) << 6) | ((y & 7) << 3) | (x & 7))));
#ifdef VR_STRICT_TU
    return lo + unorm * (hi - lo);        // -fmad=false: two roundings, the arithmetic of the oracle / the compiled GLSL
#else
    return fmaf(unorm, hi - lo, lo);      // explicit: the same rounding at every call site (tracer, decoded apron blocks)
#endif
}

// lookup_majorant (common.glsl:278-281) without the density_scale factor
VR_DEV float brick_majorant(const GridView& g, float3 ipos, int mip) {
    const int bx = int(floorf(ipos.x)) >> (3 + mip), by = int(floorf(ipos.y)) >> (3 + mip), bz = int(floorf(ipos.z)) >> (3 + mip);
    const uint32_t nx = g.nb.x >> mip, ny = g.nb.y >> mip, nz = g.nb.z >> mip;
    if (unsigned(bx) >= nx || unsigned(by) >= ny || unsigned(bz) >= nz) return 0.f;
    const size_t i = (size_t(bz) * ny + by) * nx + bx;
    const uint32_t w = mip == 0 ? __ldg(&g.rec[i].y) : __ldg(g.mips[mip - 1] + i);
    return range_hi(w);
}

// stochastic_tricubic_filter (common.glsl:221-244): 9 draws, weighted reservoir over the 4 B-spline taps.
// FastMath tests `r * sum < w` instead of `r < w / max(1e-3, sum)` (no division; the running sums of the B-spline weights
// are w0 + w1 in [1/6, 5/6], 1 - w3 >= 5/6 and 1, so the max() never acts) and evaluates the weights in Horner form with
// the 1/6 folded into the coefficients (21 instead of 34 instructions per axis).
#ifndef VR_DECODED_TAP
#define VR_DECODED_TAP 0    // 1: the non-TF kernel's single tricubic tap reads the decoded blocks too
#endif
#ifndef VR_ENV_SPLIT
#define VR_ENV_SPLIT 1      // sample_environment reads the precomputed split tables (k_env_split)
#endif
#ifndef VR_LEAN_FILTER
#define VR_LEAN_FILTER 1
#endif
template <class MT>
VR_DEV int3 stochastic_tricubic_filter(float3 ipos, uint32_t& seed) {
    const float qx = ipos.x - 0.5f, qy = ipos.y - 0.5f, qz = ipos.z - 0.5f;
    const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
    const float t[3] = { qx - fx, qy - fy, qz - fz };
    int idx[3];
    float r[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = rng(seed);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float t1 = t[a], t2 = t1 * t1;
        int k = 0;
        if (MT::fast && VR_LEAN_FILTER) {
            const float w0 = fmaf(fmaf(fmaf(-1.f / 6.f, t1, 0.5f), t1, -0.5f), t1, 1.f / 6.f);
            const float w1 = fmaf(fmaf(0.5f, t1, -1.f), t2, 2.f / 3.f);
            const float w2 = fmaf(fmaf(fmaf(-0.5f, t1, 0.5f), t1, 0.5f), t1, 1.f / 6.f);
            const float w3 = (1.f / 6.f) * t1 * t2;
            const float s1 = w0 + w1, s2 = s1 + w2, s3 = s2 + w3;
            if (r[a] * s1 < w1) k = 1;
            if (r[3 + a] * s2 < w2) k = 2;
            if (r[6 + a] * s3 < w3) k = 3;
        } else {
            float w = (1.f / 6.f) * (-t1 * t2 + 3 * t2 - 3 * t1 + 1);
            float sum = w;
            w = (1.f / 6.f) * (3 * t1 * t2 - 6 * t2 + 4);
            sum = w + sum;
            if (MT::fast ? (r[a] * fmaxf(1e-3f, sum) < w) : (r[a] < w / fmaxf(1e-3f, sum))) k = 1;
            w = (1.f / 6.f) * (-3 * t1 * t2 + 3 * t2 + 3 * t1 + 1);
            sum = w + sum;
            if (MT::fast ? (r[3 + a] * fmaxf(1e-3f, sum) < w) : (r[3 + a] < w / fmaxf(1e-3f, sum))) k = 2;
            w = (1.f / 6.f) * t1 * t2;
            sum = w + sum;
            if (MT::fast ? (r[6 + a] * fmaxf(1e-3f, sum) < w) : (r[6 + a] < w / fmaxf(1e-3f, sum))) k = 3;
        }
        idx[a] = k;
    }
    return make_int3(int(fx) + idx[0] - 1, int(fy) + idx[1] - 1, int(fz) + idx[2] - 1);
}

// tf_window + tf_lookup (common.glsl:203-212)
template <class MT>
VR_DEV float4 tf_lookup(const TraceArgs& a, float d) {
    const float tc = fminf(fmaxf(MT::div(d - a.p.tf_window_left, a.p.tf_window_width), 0.0f), 1.0f - 1e-6f);
    const float s = tc * float(a.tf_size);
    const float fl = floorf(s);
    const int idx = int(fl);
    const float f = s - fl;
    const uint32_t idx1 = min(uint32_t(idx + 1), a.tf_size - 1u);
    const float4 A = __ldg(a.lut + idx), B = __ldg(a.lut + idx1);
    return make_float4(mixf(A.x, B.x, f), mixf(A.y, B.y, f), mixf(A.z, B.z, f), mixf(A.w, B.w, f));
}
// only the alpha channel (majorant mapping, common.glsl:425/472); always IEEE: it fills the majorant tables
VR_DEV float tf_lookup_alpha(const TraceArgs& a, float d) {
    const float tc = fminf(fmaxf((d - a.p.tf_window_left) / a.p.tf_window_width, 0.0f), 1.0f - 1e-6f);
    const float s = tc * float(a.tf_size);
    const float fl = floorf(s);
    const int idx = int(fl);
    const float f = s - fl;
    const uint32_t idx1 = min(uint32_t(idx + 1), a.tf_size - 1u);
    return mixf(__ldg(&a.lut[idx].w), __ldg(&a.lut[idx1].w), f);
}

// lookup_density_trilinear (common.glsl:289-297) without density_scale: eight nearest decodes at ipos - 0.5.
// Branch-free: the footprint touches at most two bricks per axis; every tap reads its record from the padded record
// array (border / out-of-atlas entries point at an all-zero brick with range 0, i.e. texelFetch out of bounds -> 0)
// and its byte from the brick-linear atlas. All lanes of a warp run the same ~170 instructions whether or not the
// footprint straddles bricks (the previous in-brick fast path + rolled slow path cost ~600 issue slots per collision
// event because almost every warp had lanes in both, profiles/r01_v3_trace_tf_lines.txt).
VR_DEV float density_trilinear(const GridView& g, float3 ipos) {
    const float qx = ipos.x - 0.5f, qy = ipos.y - 0.5f, qz = ipos.z - 0.5f;
    const float flx = floorf(qx), fly = floorf(qy), flz = floorf(qz);
    const float fx = qx - flx, fy = qy - fly, fz = qz - flz;
    const int x = int(flx), y = int(fly), z = int(flz);
    const int bx = x >> 3, by = y >> 3, bz = z >> 3;
    // base brick in [-1, nb - 1]: otherwise all eight taps are outside the grid
    if (unsigned(bx + 1) > g.nb.x || unsigned(by + 1) > g.nb.y || unsigned(bz + 1) > g.nb.z) return 0.f;
    const uint32_t lx = x & 7, ly = y & 7, lz = z & 7;
    const uint32_t sx = lx == 7u, sy = ly == 7u ? g.psx : 0u, sz = lz == 7u ? g.psxy : 0u;   // record strides taken by the +1 taps
    const uint32_t base = uint32_t(bz + 1) * g.psxy + uint32_t(by + 1) * g.psx + uint32_t(bx + 1);
    const uint32_t lx1 = (lx + 1u) & 7u;
    float row[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t dy = k & 1, dz = k >> 1;
        const uint32_t ridx = base + (dy ? sy : 0u) + (dz ? sz : 0u);
        const uint2 ra = __ldg(g.recp + ridx), rb = __ldg(g.recp + ridx + sx);
        const uint32_t voff = (((lz + dz) & 7u) << 6) | (((ly + dy) & 7u) << 3);
        const uint32_t a = __ldg(g.atlas_lin + size_t(ra.x) * 512u + (voff | lx));
        const uint32_t b = __ldg(g.atlas_lin + size_t(rb.x) * 512u + (voff | lx1));
        const float loa = range_lo(ra.y), lob = range_lo(rb.y);
        const float va = loa + unorm8_to_float(a) * (range_hi(ra.y) - loa);
        const float vb = lob + unorm8_to_float(b) * (range_hi(rb.y) - lob);
        row[k] = mixf(va, vb, fx);
    }
    return mixf(mixf(row[0], row[1], fy), mixf(row[2], row[3], fy), fz);
}

// The same value from the decoded apron bricks: one slot load + 8 independent fp32 loads at fixed offsets from one
// base address (~55 instructions instead of ~170; 8 dependent record -> byte chains become 1 -> 8). Every stored voxel is
// brick_value() evaluated by k_decode_cells with the expression above, and the lerp order is the same, so the result is
// bit-identical to density_trilinear (checked by tests through vrb_debug_sample_density).
VR_DEV float density_trilinear_decoded(const GridView& g, float3 ipos) {
    const float qx = ipos.x - 0.5f, qy = ipos.y - 0.5f, qz = ipos.z - 0.5f;
    const float flx = floorf(qx), fly = floorf(qy), flz = floorf(qz);
    const float fx = qx - flx, fy = qy - fly, fz = qz - flz;
    const int x = int(flx), y = int(fly), z = int(flz);
    const int bx = x >> 3, by = y >> 3, bz = z >> 3;
    if (unsigned(bx + 1) > g.nb.x || unsigned(by + 1) > g.nb.y || unsigned(bz + 1) > g.nb.z) return 0.f;
    const uint32_t cell = (uint32_t(bz + 1) * (g.nb.y + 1u) + uint32_t(by + 1)) * (g.nb.x + 1u) + uint32_t(bx + 1);
    const float* b = g.datlas + size_t(__ldg(g.cslot + cell)) * DBRICK + uint32_t((z & 7) * 81 + (y & 7) * 9 + (x & 7));
    const float v000 = __ldg(b), v100 = __ldg(b + 1), v010 = __ldg(b + 9), v110 = __ldg(b + 10);
    const float v001 = __ldg(b + 81), v101 = __ldg(b + 82), v011 = __ldg(b + 90), v111 = __ldg(b + 91);
    return mixf(mixf(mixf(v000, v100, fx), mixf(v010, v110, fx), fy), mixf(mixf(v001, v101, fx), mixf(v011, v111, fx), fy), fz);
}

// One decoded voxel: the value brick_value() returns, from the block of the brick that contains it (1 slot load + 1 fp32
// load instead of record load + range decode + byte load + unorm conversion).
VR_DEV float decoded_value(const GridView& g, int x, int y, int z) {
    const int bx = x >> 3, by = y >> 3, bz = z >> 3;
    if (unsigned(bx) >= g.nb.x || unsigned(by) >= g.nb.y || unsigned(bz) >= g.nb.z) return 0.f;
    const uint32_t cell = (uint32_t(bz + 1) * (g.nb.y + 1u) + uint32_t(by + 1)) * (g.nb.x + 1u) + uint32_t(bx + 1);
    return __ldg(g.datlas + size_t(__ldg(g.cslot + cell)) * DBRICK + uint32_t((z & 7) * 81 + (y & 7) * 9 + (x & 7)));
}

// ---- building the decoded apron bricks -----------------------------------------------------------------
// flags[cell] = 1 when any of the cell's 8 bricks (b + {0,1}^3) lies in the grid with a range other than (0, 0)
VR_GLOBAL void k_cell_flags(const uint2* __restrict__ rec, uint3 nb, uint32_t* __restrict__ flags) {
    const uint32_t cx = nb.x + 1, cy = nb.y + 1, cz = nb.z + 1;
    const size_t n = size_t(cx) * cy * cz;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const int bx = int(i % cx) - 1, by = int((i / cx) % cy) - 1, bz = int(i / (size_t(cx) * cy)) - 1;
        uint32_t f = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int x = bx + (k & 1), y = by + ((k >> 1) & 1), z = bz + (k >> 2);
            if (unsigned(x) < nb.x && unsigned(y) < nb.y && unsigned(z) < nb.z) {
                const uint32_t w = rec[(size_t(z) * nb.y + y) * nb.x + x].y;
                if (range_lo(w) != 0.f || range_hi(w) != 0.f) f = 1;
            }
        }
        flags[i] = f;
    }
}
// cslot = 1 + (number of flagged cells before this one) for flagged cells, 0 (the shared zero block) otherwise
VR_GLOBAL void k_cell_slots(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ excl, size_t n, uint32_t* __restrict__ cslot) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) cslot[i] = flags[i] ? excl[i] + 1u : 0u;
}
// one warp per cell with a block of its own: 729 decoded voxels
VR_GLOBAL void __launch_bounds__(256) k_decode_cells(const GridView g, const uint32_t* __restrict__ cslot, float* __restrict__ datlas) {
    const uint32_t cx = g.nb.x + 1, cy = g.nb.y + 1, cz = g.nb.z + 1;
    const size_t n = size_t(cx) * cy * cz;
    const int lane = threadIdx.x & 31;
    for (size_t i = (blockIdx.x * size_t(blockDim.x) + threadIdx.x) >> 5; i < n; i += (size_t(gridDim.x) * blockDim.x) >> 5) {
        const uint32_t s = cslot[i];
        if (s == 0u) continue;
        const int x0 = (int(i % cx) - 1) * 8, y0 = (int((i / cx) % cy) - 1) * 8, z0 = (int(i / (size_t(cx) * cy)) - 1) * 8;
        for (uint32_t v = lane; v < DBRICK; v += 32) {
            const int lx = int(v % 9u), ly = int((v / 9u) % 9u), lz = int(v / 81u);
            datlas[size_t(s) * DBRICK + v] = brick_value(g, x0 + lx, y0 + ly, z0 + lz);
        }
    }
}

// lookup_emission (common.glsl:324-328). Without an emission grid the samplers are unbound (value 0) but the
// tricubic filter still consumes 9 draws: that case is an O(1) LCG jump.
template <class MT>
VR_DEV float3 lookup_emission(const TraceArgs& a, float3 ipos, uint32_t& seed, bool& fetched) {
    fetched = false;
    if (!a.p.has_emission) { rng_skip<9>(seed); return f3(0.f); }
    const float3 ipos_e = mul_point(a.emis_from_density, ipos);
    const int3 tap = stochastic_tricubic_filter<MT>(ipos_e, seed);
    fetched = true;
    const float t = brick_value(a.emission, tap.x, tap.y, tap.z) * a.p.vol_emission_norm;
    return a.p.vol_emission_scale * f3(sqr(t), sqr(sqr(t)), sqr(sqr(sqr(t))));
}

// intersect_box (common.glsl:157-165)
template <class MT>
VR_DEV bool intersect_box(float3 pos, float3 dir, const float* bb_min, const float* bb_max, float& tnear, float& tfar) {
    const float3 inv = f3(MT::rcp(dir.x), MT::rcp(dir.y), MT::rcp(dir.z));
    const float3 lo = (f3(bb_min[0], bb_min[1], bb_min[2]) - pos) * inv;
    const float3 hi = (f3(bb_max[0], bb_max[1], bb_max[2]) - pos) * inv;
    const float3 tmin = f3(fminf(lo.x, hi.x), fminf(lo.y, hi.y), fminf(lo.z, hi.z));
    const float3 tmax = f3(fmaxf(lo.x, hi.x), fmaxf(lo.y, hi.y), fmaxf(lo.z, hi.z));
    tnear = fmaxf(0.f, fmaxf(tmin.x, fmaxf(tmin.y, tmin.z)));
    tfar = fminf(tmax.x, fminf(tmax.y, tmax.z));
    return tnear <= tfar;
}

// stepDDA (common.glsl:404-409)
VR_DEV float step_dda(float3 pos, float3 ri, int mip) {
    // dim = 8 << mip and 1/dim as exact powers of two, straight from the exponent bits
    const float dim = __uint_as_float(0x41000000u + (uint32_t(mip) << 23)), inv_dim = __uint_as_float(0x3e000000u - (uint32_t(mip) << 23));
    const float ox = ri.x >= 0.f ? dim + 0.5f : -0.5f;
    const float oy = ri.y >= 0.f ? dim + 0.5f : -0.5f;
    const float oz = ri.z >= 0.f ? dim + 0.5f : -0.5f;
    const float tx = (floorf(pos.x * inv_dim) * dim + ox - pos.x) * ri.x;
    const float ty = (floorf(pos.y * inv_dim) * dim + oy - pos.y) * ri.y;
    const float tz = (floorf(pos.z * inv_dim) * dim + oz - pos.z) * ri.z;
    return fminf(tx, fminf(ty, tz));
}

// GLSL round() is implementation-defined at .5; Mesa lowers it to round-half-even
VR_DEV int round_mip(float mip) { return __float2int_rn(mip); }

template <class MT>
VR_DEV float phase_hg(float cos_t, float g) {  // common.glsl:172-175
    const float denom = 1 + sqr(g) + 2 * g * cos_t;
    return MT::div(INV_4PI * (1 - sqr(g)), denom * MT::sqrt(denom));
}
template <class MT>
VR_DEV float3 align_to(float3 N, float3 v) {  // common.glsl:25-33
    float3 T;
    if (fabsf(N.x) > fabsf(N.y)) T = MT::fast ? f3(-N.z, 0.f, N.x) * rsqrtf(N.x * N.x + N.z * N.z) : f3(-N.z, 0.f, N.x) / sqrtf(N.x * N.x + N.z * N.z);
    else T = MT::fast ? f3(0.f, N.z, -N.y) * rsqrtf(N.y * N.y + N.z * N.z) : f3(0.f, N.z, -N.y) / sqrtf(N.y * N.y + N.z * N.z);
    const float3 B = cross(N, T);
    return MT::normalize(v.x * T + v.y * B + v.z * N);
}
template <class MT>
VR_DEV float3 sample_phase_hg(float3 dir, float g, float s0, float s1) {  // common.glsl:184-190
    const float cos_t = fabsf(g) < 1e-4f ? 1.f - 2.f * s0 : MT::div(1 + sqr(g) - sqr(MT::div(1 - sqr(g), 1 - g + 2 * g * s0)), 2 * g);
    const float sin_t = MT::sqrt(fmaxf(0.f, 1.f - sqr(cos_t)));
    const float phi = 2.f * PI_F * s1;
    float sp, cp;
    MT::sincos(phi, &sp, &cp);
    return align_to<MT>(dir, f3(sin_t * cp, sin_t * sp, cos_t));
}

// bilinear REPEAT fetch with a cheap wrap (uv in [-1, 2] never needs the integer modulo)
VR_DEV int wrap_fast(int i, int n) {
    if (i < 0) i += n; else if (i >= n) i -= n;
    if (unsigned(i) >= unsigned(n)) { i %= n; if (i < 0) i += n; }   // only for |uv| > 2: never taken by the tracer's own lookups
    return i;
}
VR_DEV float3 env_texture_fw(const EnvView& e, float u, float v) {
    const float x = u * float(e.w) - 0.5f, y = v * float(e.h) - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    const int x0 = wrap_fast(int(fx), e.w), y0 = wrap_fast(int(fy), e.h);
    const int x1 = x0 + 1 == e.w ? 0 : x0 + 1, y1 = y0 + 1 == e.h ? 0 : y0 + 1;
    const float4 t00 = __ldg(e.rgb + size_t(y0) * e.w + x0), t10 = __ldg(e.rgb + size_t(y0) * e.w + x1);
    const float4 t01 = __ldg(e.rgb + size_t(y1) * e.w + x0), t11 = __ldg(e.rgb + size_t(y1) * e.w + x1);
    return f3(mixf(mixf(t00.x, t10.x, ax), mixf(t01.x, t11.x, ax), ay),
              mixf(mixf(t00.y, t10.y, ax), mixf(t01.y, t11.y, ax), ay),
              mixf(mixf(t00.z, t10.z, ax), mixf(t01.z, t11.z, ax), ay));
}

// lookup_environment (common.glsl:93-98)
VR_DEV float3 lookup_environment(const TraceArgs& a, float3 dir) {
    const float3 idir = mul(*reinterpret_cast<const Mat3*>(a.p.env_inv_transform), dir);
    const float u = atan2f(idir.z, idir.x) / (2 * PI_F) + 0.5f;
    const float v = 1.f - acosf(fminf(fmaxf(idir.y, -1.f), 1.f)) / PI_F;
    return a.p.env_strength * env_texture_fw(a.env, u, v);
}
// pdf_environment (common.glsl:148-152), given Le = lookup_environment(dir)
template <class MT>
VR_DEV float pdf_environment(const TraceArgs& a, float3 Le_dir) {
    const float avg_w = __ldg(a.env.impmap + imp_offset(9));
    return MT::div(luma(Le_dir), avg_w) * INV_4PI;
}

// sample_environment (common.glsl:100-146): hierarchical 2x2 warping down the importance pyramid (rolled: code size)
// The per-level quantities dsplit, e and the reciprocals the remaps multiply with depend on the quad only: k_env_split
// evaluates them once per environment with these very expressions (FastMath: a * rcp.approx(b)), so the table path
// returns the same bits with 2 dependent 16-byte loads and no MUFU per level instead of 2 loads + 4 MUFU.RCP.
VR_GLOBAL void k_env_split(const float* __restrict__ impmap, float4* __restrict__ split) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < SPLIT_QUADS; i += gridDim.x * blockDim.x) {
        int mip = 8;
        while (mip > 0 && i >= split_offset(mip - 1)) --mip;
        const uint32_t q = i - split_offset(mip), half = (IMP_DIM >> mip) >> 1, d = IMP_DIM >> mip;
        const uint32_t qx = q % half, qy = q / half;
        const float* base = impmap + imp_offset(mip) + size_t(2 * qy) * d + 2 * qx;
        const float w0 = base[0], w1 = base[1], w2 = base[d], w3 = base[d + 1];
        const float q0 = w0 + w2, q1 = w1 + w3;
        const float dsplit = FastMath::div(q0, fmaxf(1e-8f, q0 + q1));
        const float el = FastMath::div(w0, q0), er = FastMath::div(w1, q1);
        split[3 * size_t(i)] = make_float4(dsplit, FastMath::rcp(dsplit), FastMath::rcp(1.f - dsplit), 0.f);
        split[3 * size_t(i) + 1] = make_float4(el, FastMath::rcp(el), FastMath::rcp(1.f - el), 0.f);
        split[3 * size_t(i) + 2] = make_float4(er, FastMath::rcp(er), FastMath::rcp(1.f - er), 0.f);
    }
}

template <class MT>
VR_DEV float4 sample_environment(const TraceArgs& a, float px, float py, float3& w_i) {
    int posx = 0, posy = 0;
    uint32_t off = imp_offset(9);
    if (MT::fast && VR_ENV_SPLIT && a.env.split) {
        uint32_t qoff = 0, half = 1;          // quad offset and quads per row of the level
#pragma unroll 1
        for (int mip = 8; mip >= 0; --mip) {
            const float4* q = a.env.split + 3 * size_t(qoff + uint32_t(posy) * half + uint32_t(posx));
            const float4 s = __ldg(q);
            const bool right = !(px < s.x);
            px = right ? (px - s.x) * s.z : px * s.y;
            const float4 e = __ldg(q + (right ? 2 : 1));
            const bool top = !(py < e.x);
            py = top ? (py - e.x) * e.z : py * e.y;
            posx = 2 * posx + (right ? 1 : 0);
            posy = 2 * posy + (top ? 1 : 0);
            qoff += half * half;
            half *= 2;
        }
    } else
#pragma unroll 1
    for (int mip = 8; mip >= 0; --mip) {
        posx *= 2; posy *= 2;
        const int d = IMP_DIM >> mip;
        off -= uint32_t(d) * uint32_t(d);   // offset of level `mip`
        const float* base = a.env.impmap + off + size_t(posy) * d + posx;
        const float2 r0 = __ldg(reinterpret_cast<const float2*>(base));        // w[0], w[1]
        const float2 r1 = __ldg(reinterpret_cast<const float2*>(base + d));    // w[2], w[3]
        const float q0 = r0.x + r1.x, q1 = r0.y + r1.y;
        const float dsplit = MT::div(q0, fmaxf(1e-8f, q0 + q1));
        const bool right = !(px < dsplit);
        px = right ? MT::div(px - dsplit, 1.f - dsplit) : MT::div(px, dsplit);
        posx += right ? 1 : 0;
        const float e = MT::div(right ? r0.y : r0.x, right ? q1 : q0);
        const bool top = !(py < e);
        py = top ? MT::div(py - e, 1.f - e) : MT::div(py, e);
        posy += top ? 1 : 0;
    }
    const float uvx = (float(posx) + px) * (1.f / IMP_DIM), uvy = (float(posy) + py) * (1.f / IMP_DIM);
    const float theta = saturate(1.f - uvy) * PI_F;
    const float phi = (saturate(uvx) * 2.f - 1.f) * PI_F;
    float st, ct, sp, cp;
    MT::sincos(theta, &st, &ct);
    MT::sincos(phi, &sp, &cp);
    w_i = mul(*reinterpret_cast<const Mat3*>(a.p.env_transform), f3(st * cp, ct, st * sp));
    const float3 Le = a.p.env_strength * env_texture_fw(a.env, uvx, uvy);
    const float avg_w = __ldg(a.env.impmap + imp_offset(9));
    const float pdf = MT::div(__ldg(a.env.impmap + size_t(posy) * IMP_DIM + posx), avg_w);
    return make_float4(Le.x, Le.y, Le.z, pdf * INV_4PI);
}

struct Ray {
    float3 ipos, idir, ri;
    float tnear, tfar;
};
template <class MT>
VR_DEV bool setup_ray(const TraceArgs& a, float3 wpos, float3 wdir, Ray& r) {
    if (!intersect_box<MT>(wpos, wdir, a.p.vol_bb_min, a.p.vol_bb_max, r.tnear, r.tfar)) return false;
    const Mat4& M = *reinterpret_cast<const Mat4*>(a.p.vol_density_inv_transform);
    r.ipos = mul_point(M, wpos);
    r.idir = mul_dir(M, wdir);  // non-normalised
    r.ri = f3(MT::rcp(r.idir.x), MT::rcp(r.idir.y), MT::rcp(r.idir.z));
    return true;
}

constexpr int MAX_DDA_ITERS = 1 << 22;  // hang guard only; never reached by finite rays

template <bool TF, bool COUNT>
VR_DEV float majorant_at(const TraceArgs& a, float3 curr, int mip, Cnt<COUNT>& cnt) {
    cnt.maj();
    const float m = a.p.vol_density_scale * brick_majorant(a.density, curr, mip);
    if (TF) return a.p.vol_majorant * tf_lookup_alpha(a, m * a.p.vol_inv_majorant);
    return m;
}

// transmittanceDDA (common.glsl:412-455)
template <bool TF, bool COUNT, class MT>
VR_DEV float transmittance_dda(const TraceArgs& a, float3 wpos, float3 wdir, uint32_t& seed, Cnt<COUNT>& cnt) {
    Ray r;
    if (!setup_ray<MT>(a, wpos, wdir, r)) return 1.f;
    float t = r.tnear + 1e-6f, Tr = 1.f, tau = -MT::log(1.f - rng(seed)), mip = 3.f;
    for (int it = 0; t < r.tfar && it < MAX_DDA_ITERS; ++it) {
        const float3 curr = r.ipos + t * r.idir;
        const int m = round_mip(mip);
        const float majorant = majorant_at<TF, COUNT>(a, curr, m, cnt);
        const float dt = step_dda(curr, r.ri, m);
        t += dt;
        tau -= majorant * dt;
        mip = fminf(mip + 0.25f, 3.f);
        if (tau > 0) continue;
        t += MT::div(tau, majorant);
        if (t >= r.tfar) break;
        cnt.dens();
        float d;
        if (TF) d = a.p.vol_majorant * tf_lookup<MT>(a, a.p.vol_density_scale * density_trilinear(a.density, r.ipos + t * r.idir) * a.p.vol_inv_majorant).w;
        else {
            const int3 tap = stochastic_tricubic_filter<MT>(r.ipos + t * r.idir, seed);
            d = a.p.vol_density_scale * brick_value(a.density, tap.x, tap.y, tap.z);
        }
        if (rng(seed) * majorant < d) {
            Tr *= fmaxf(0.f, 1.f - MT::div(a.p.vol_majorant, majorant));
            if (Tr < .1f) {
                const float prob = 1 - Tr;
                if (rng(seed) < prob) return 0.f;
                Tr = MT::div(Tr, 1 - prob);
            }
        }
        tau = -MT::log(1.f - rng(seed));
        mip = fmaxf(0.f, mip - 2.f);
    }
    return Tr;
}

// sample_volumeDDA (common.glsl:458-501)
template <bool TF, bool COUNT, class MT>
VR_DEV bool sample_volume_dda(const TraceArgs& a, float3 wpos, float3 wdir, float& t_out, float3& throughput, float3& Le, uint32_t& seed, Cnt<COUNT>& cnt) {
    Ray r;
    if (!setup_ray<MT>(a, wpos, wdir, r)) return false;
    const float3 albedo = f3(a.p.vol_albedo[0], a.p.vol_albedo[1], a.p.vol_albedo[2]);
    float t = r.tnear + 1e-6f, tau = -MT::log(1.f - rng(seed)), mip = 3.f;
    for (int it = 0; t < r.tfar && it < MAX_DDA_ITERS; ++it) {
        const float3 curr = r.ipos + t * r.idir;
        const int m = round_mip(mip);
        const float majorant = majorant_at<TF, COUNT>(a, curr, m, cnt);
        const float dt = step_dda(curr, r.ri, m);
        t += dt;
        tau -= majorant * dt;
        mip = fminf(mip + 0.25f, 3.f);
        if (tau > 0) continue;
        t += MT::div(tau, majorant);
        if (t >= r.tfar) break;
        cnt.dens();
        const float3 at = r.ipos + t * r.idir;
        float d;
        float3 tf_rgb = f3(1.f);
        if (TF) {
            const float4 rgba = tf_lookup<MT>(a, a.p.vol_density_scale * density_trilinear(a.density, at) * a.p.vol_inv_majorant);
            d = a.p.vol_majorant * rgba.w;
            tf_rgb = f3(rgba.x, rgba.y, rgba.z);
        } else {
            const int3 tap = stochastic_tricubic_filter<MT>(at, seed);
            d = a.p.vol_density_scale * brick_value(a.density, tap.x, tap.y, tap.z);
        }
        bool fetched;
        const float3 em = lookup_emission<MT>(a, at, seed, fetched);
        if (fetched) {
            cnt.emis();
            Le = Le + throughput * (f3(1.f) - albedo) * em * d * a.p.vol_inv_majorant;
        }
        if (rng(seed) * majorant < d) {
            throughput = throughput * albedo;
            if (TF) throughput = throughput * tf_rgb;
            t_out = t;
            return true;
        }
        tau = -MT::log(1.f - rng(seed));
        mip = fmaxf(0.f, mip - 2.f);
    }
    return false;
}

// trace_path (common.glsl:599-652)
template <bool TF, bool COUNT, class MT>
VR_DEV float4 trace_path(const TraceArgs& a, float3 pos, float3 dir, uint32_t& seed, Cnt<COUNT>& cnt) {
    float3 L = f3(0.f), throughput = f3(1.f);
    bool free_path = true;
    uint32_t n_paths = 0;
    float t = 0.f, f_p = 0.f;
    while (sample_volume_dda<TF, COUNT, MT>(a, pos, dir, t, throughput, L, seed, cnt)) {
        cnt.real();
        pos = pos + t * dir;
        float3 w_i;
        const float r0 = rng(seed), r1 = rng(seed);
        cnt.nee();
        const float4 Le_pdf = sample_environment<MT>(a, r0, r1, w_i);
        if (Le_pdf.w > 0) {
            f_p = phase_hg<MT>(dot(-dir, w_i), a.p.vol_phase_g);
            const float mis_weight = a.p.show_environment > 0 ? MT::div(sqr(Le_pdf.w), sqr(Le_pdf.w) + sqr(f_p)) : 1.f;
            const float Tr = transmittance_dda<TF, COUNT, MT>(a, pos, w_i, seed, cnt);
            const float3 c = throughput * mis_weight * f_p * Tr * f3(Le_pdf.x, Le_pdf.y, Le_pdf.z);
            L = L + f3(MT::div(c.x, Le_pdf.w), MT::div(c.y, Le_pdf.w), MT::div(c.z, Le_pdf.w));
        }
        if (++n_paths >= uint32_t(a.p.bounces)) { free_path = false; break; }
        const float rr_val = luma(throughput);
        if (rr_val < .1f) {
            const float prob = 1 - rr_val;
            if (rng(seed) < prob) { free_path = false; break; }
            const float k = 1 - prob;
            throughput = f3(MT::div(throughput.x, k), MT::div(throughput.y, k), MT::div(throughput.z, k));
        }
        const float s0 = rng(seed), s1 = rng(seed);
        const float3 scatter_dir = sample_phase_hg<MT>(dir, a.p.vol_phase_g, s0, s1);
        f_p = phase_hg<MT>(dot(-dir, scatter_dir), a.p.vol_phase_g);
        dir = scatter_dir;
    }
    if (free_path && a.p.show_environment > 0) {
        cnt.env();
        const float3 Le = lookup_environment(a, dir);
        const float pe = pdf_environment<MT>(a, Le);
        const float mis_weight = n_paths > 0 ? MT::div(sqr(f_p), sqr(f_p) + sqr(pe)) : 1.f;
        L = L + throughput * mis_weight * Le;
    }
    return make_float4(L.x, L.y, L.z, fminf(float(n_paths), 1.f));
}

// view_dir (common.glsl:76-80)
template <class MT>
VR_DEV float3 view_dir(const TraceArgs& a, int x, int y, float sx, float sy) {
    const float w = float(a.p.resolution[0]), h = float(a.p.resolution[1]);
    const float px = MT::div(float(x) + sx - w * .5f, h), py = MT::div(float(y) + sy - h * .5f, h);
    return MT::normalize(mul(*reinterpret_cast<const Mat3*>(a.p.cam_transform), MT::normalize(f3(px, py, a.cam_z))));
}

// running mean of pathtracer_brick.glsl:36, unfused so that it matches mix() = x*(1-a) + y*a bit for bit
VR_DEV float mix_rn(float x, float y, float a) { return __fadd_rn(__fmul_rn(x, __fsub_rn(1.f, a)), __fmul_rn(y, a)); }

// main() of pathtracer_brick.glsl:23-37, one thread per pixel, all requested samples of that pixel in
// sample order (so the running mean is evaluated exactly like n successive dispatches).
// Warps cover 8x4 pixel tiles for ray coherence. Cross-check kernel (StrictMath), not the production path.
template <bool TF, bool COUNT>
VR_GLOBAL void __launch_bounds__(256) k_trace_pixels(const __grid_constant__ TraceArgs a) {
    using MT = StrictMath;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = a.x0 + blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = a.y0 + blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= a.x1 || y >= a.y1) return;
    Cnt<COUNT> cnt;
    const int W = a.p.resolution[0];
    float4* px = a.color + size_t(y) * W + x;
    float4 acc = *px;
    const float3 cam = f3(a.p.cam_pos[0], a.p.cam_pos[1], a.p.cam_pos[2]);
    for (int s = a.first_sample; s < a.first_sample + a.n_samples; ++s) {
        uint32_t seed = tea32(uint32_t(a.p.seed) * uint32_t(y * W + x), uint32_t(s));
        const float jx = rng(seed), jy = rng(seed);
        const float3 dir = view_dir<MT>(a, x, y, jx, jy);
        float4 L = trace_path<TF, COUNT, MT>(a, cam, dir, seed, cnt);
        L.x = sanitize(L.x); L.y = sanitize(L.y); L.z = sanitize(L.z); L.w = sanitize(L.w);
        cnt.samp();
        if (a.accum_mode == VRB_ACCUM_MEAN) {
            const float w = 1.f / float(s);
            acc.x = mix_rn(acc.x, L.x, w); acc.y = mix_rn(acc.y, L.y, w); acc.z = mix_rn(acc.z, L.z, w); acc.w = mix_rn(acc.w, L.w, w);
        } else {
            acc.x += L.x; acc.y += L.y; acc.z += L.z; acc.w += L.w;
        }
    }
    *px = acc;
    flush_counters(a, cnt);
}

// ------------------------------------------------------------------------------------------------
// deterministic transmittance-only mode ("T1"): centre ray, exact voxel DDA with empty-brick
// skipping, color = (Tr * Le_env(dir), 1 - Tr). fp32 on the device; the oracle evaluates it in fp64.
VR_GLOBAL void __launch_bounds__(256) k_trace_deterministic(const __grid_constant__ TraceArgs a) {
    using MT = StrictMath;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= a.p.resolution[0] || y >= a.p.resolution[1]) return;
    const float3 dir = view_dir<MT>(a, x, y, .5f, .5f);
    const float3 pos = f3(a.p.cam_pos[0], a.p.cam_pos[1], a.p.cam_pos[2]);
    Ray r;
    float tau = 0.f;
    if (setup_ray<MT>(a, pos, dir, r)) {
        float t = r.tnear;
        const GridView& g = a.density;
        for (int it = 0; t < r.tfar && it < MAX_DDA_ITERS; ++it) {
            const float tp = t + 1e-6f * (1.f + fabsf(t));
            const float3 q = r.ipos + tp * r.idir;
            const int vx = int(floorf(q.x)), vy = int(floorf(q.y)), vz = int(floorf(q.z));
            // brick-level skip when the whole brick is empty (or out of bounds): cell size 8, else 1
            const int bx = vx >> 3, by = vy >> 3, bz = vz >> 3;
            bool empty = true;
            uint2 rec = make_uint2(0xffffffffu, 0u);
            if (unsigned(bx) < g.nb.x && unsigned(by) < g.nb.y && unsigned(bz) < g.nb.z) {
                rec = __ldg(g.rec + (size_t(bz) * g.nb.y + by) * g.nb.x + bx);
                empty = range_lo(rec.y) == 0.f && range_hi(rec.y) == 0.f;
            }
            const int cs = empty ? 8 : 1;
            const int cx = empty ? bx * 8 : vx, cy = empty ? by * 8 : vy, cz = empty ? bz * 8 : vz;
            float tnext = r.tfar;
            if (r.idir.x > 0.f) tnext = fminf(tnext, (float(cx + cs) - r.ipos.x) * r.ri.x); else if (r.idir.x < 0.f) tnext = fminf(tnext, (float(cx) - r.ipos.x) * r.ri.x);
            if (r.idir.y > 0.f) tnext = fminf(tnext, (float(cy + cs) - r.ipos.y) * r.ri.y); else if (r.idir.y < 0.f) tnext = fminf(tnext, (float(cy) - r.ipos.y) * r.ri.y);
            if (r.idir.z > 0.f) tnext = fminf(tnext, (float(cz + cs) - r.ipos.z) * r.ri.z); else if (r.idir.z < 0.f) tnext = fminf(tnext, (float(cz) - r.ipos.z) * r.ri.z);
            if (tnext <= t) tnext = tp;
            if (!empty) {
                const float lo = range_lo(rec.y), hi = range_hi(rec.y);
                float unorm = 0.f;
                if (rec.x != 0xffffffffu) unorm = unorm8_to_float(__ldg(g.atlas_lin + size_t(rec.x) * 512u + uint32_t(((vz & 7) << 6) | ((vy & 7) << 3) | (vx & 7))));
                tau += (lo + unorm * (hi - lo)) * (tnext - t);
            }
            t = tnext;
        }
    }
    const float Tr = expf(-a.p.vol_density_scale * tau);
    const float3 Le = lookup_environment(a, dir);
    a.color[size_t(y) * a.p.resolution[0] + x] = make_float4(Tr * Le.x, Tr * Le.y, Tr * Le.z, 1.f - Tr);
}

}  // namespace vr
