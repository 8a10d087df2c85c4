/*
 * vr_oracle.c -- TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's hot path.
 *
 * Never linked, loaded or called by the product (libvrb200.so / C++ host / volpy). Allowed users:
 * tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).
 *
 * What is restated (all paths relative to the reference checkout, nihofm/volren @ e8aea40):
 *   voldata layer   submodules/voldata/src/grid_dense.cpp:57-103, grid_brick.cpp:24-154,
 *                   glm/detail/type_half.inl:105-238            -> PINNED bit-exactly vs oracle/_ref
 *   LUT             src/transferfunc.cpp:33-58
 *   env setup       shader/env_setup.glsl:18-34, src/environment.cpp:11-33 (+ glGenerateMipmap box filter)
 *   path tracer     shader/pathtracer_brick.glsl:23-37, pathtracer_brick_tf.glsl:24-38,
 *                   shader/common.glsl:10-67,72-80,85-152,157-190,195-212,221-244,249-328,399-501,596-652
 *   tonemap         shader/tonemap.glsl:13-36, shader/tonemap.fs:10-28, shader/blit.fs
 * The shader layer is PARITY UNPINNED: the GLSL programs cannot execute here (no GL), so there is no
 * reference output to pin against; GL fixed-function behaviour is pinned BY DECISION as:
 *   bilinear with REPEAT on both axes, float weights, texel centres at (i+.5)/N, LOD 0;
 *   glGenerateMipmap = 0.25f*((a+b)+(c+d)); unorm8 -> float = c/255.f; float -> unorm8 = rint(clamp*255);
 *   out-of-bounds texelFetch = 0 (robust buffer access); GLSL round() = round-half-even (Mesa lowers
 *   round() to fround_even; llvmpipe is the reference's stated CPU path);
 *   min/max/clamp/mix exactly as the GLSL spec formulas; acos argument clamped to [-1,1].
 * All arithmetic is fp32 without FMA contraction (-ffp-contract=off), except the deterministic mode
 * which is fp64 by definition (DESIGN.md "T1").
 */
#include "vr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ */
/* small vector helpers (GLSL semantics)                                                            */

typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;
typedef struct { float x, y, z, w; } v4;
typedef struct { int x, y, z; } i3;

static const float PI_F = (float)3.14159265358979323846; /* common.glsl:4 */
#define INV_PI (1.f / PI_F)
#define INV_2PI (1.f / (2 * PI_F))
#define INV_4PI (1.f / (4 * PI_F))

static inline v3 V3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scale3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross3(v3 a, v3 b) { return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
/* normalize(v) = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x): GLSL 4.50 spec 8.5 as glm states it
 * (glm/detail/func_geometric.inl compute_normalize, func_exponential.inl inversesqrt) -- pinned by oracle/_ref/libglsl_ref.so */
static inline v3 normalize3(v3 a) { const float r = 1.f / sqrtf(dot3(a, a)); return V3(a.x * r, a.y * r, a.z * r); }
static inline float gl_min(float x, float y) { return y < x ? y : x; }   /* GLSL 4.50 spec 8.3 */
static inline float gl_max(float x, float y) { return x < y ? y : x; }
static inline float gl_clamp(float x, float lo, float hi) { return gl_min(gl_max(x, lo), hi); }
static inline float gl_mix(float x, float y, float a) { return x * (1.f - a) + y * a; }
static inline float gl_fract(float x) { return x - floorf(x); }
static inline float sqr(float x) { return x * x; }                      /* common.glsl:10 */
static inline float luma(v3 c) { return dot3(c, V3(0.212671f, 0.715160f, 0.072169f)); } /* :21 */
static inline float sanitize1(float x) { return (isnan(x) || isinf(x)) ? 0.f : x; }     /* :17-19 */
static inline float saturate(float x) { return gl_clamp(x, 0.f, 1.f); }                 /* :23 */
static inline float power_heuristic(float a, float b) { return sqr(a) / (sqr(a) + sqr(b)); } /* :35 */
/* float -> int as the GPU does it (saturating, NaN -> 0). GLSL leaves out-of-range conversions undefined and the
 * reference reaches them: rng() == 0 with a zero majorant gives t = 0/0 (common.glsl:434/481), after which lookups
 * run on NaN positions; robust texel/buffer access makes that harmless on a GPU, a plain C cast would index wildly. */
static inline int f2i(float x) {
    if (isnan(x)) return 0;
    if (x >= 2147483648.f) return 2147483647;
    if (x <= -2147483648.f) return -2147483647 - 1;
    return (int)x;
}

/* column-major mat3 * vec3, mat4 * vec4 */
static inline v3 m3mul(const float* m, v3 v) {
    return V3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z, m[2] * v.x + m[5] * v.y + m[8] * v.z);
}
/* mat4 * vec4 sums the four column products pairwise, (c0 x + c1 y) + (c2 z + c3 w): glm/detail/type_mat4x4.inl operator* */
static inline v3 m4mul_xyz(const float* m, v3 v, float w) {
    return V3((m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * w), (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * w),
              (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w));
}
static void m4mul(const float* a, const float* b, float* out) { /* out = a * b */
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
            out[c * 4 + r] = a[0 * 4 + r] * b[c * 4 + 0] + a[1 * 4 + r] * b[c * 4 + 1] + a[2 * 4 + r] * b[c * 4 + 2] + a[3 * 4 + r] * b[c * 4 + 3];
}

/* common.glsl:25-33 */
static v3 align_to(v3 N, v3 v) {
    v3 T;
    if (fabsf(N.x) > fabsf(N.y)) {
        const float l = sqrtf(N.x * N.x + N.z * N.z);
        T = V3(-N.z / l, 0.f / l, N.x / l);
    } else {
        const float l = sqrtf(N.y * N.y + N.z * N.z);
        T = V3(0.f / l, N.z / l, -N.y / l);
    }
    const v3 B = cross3(N, T);
    return normalize3(add3(add3(scale3(T, v.x), scale3(B, v.y)), scale3(N, v.z)));
}

/* ------------------------------------------------------------------------------------------------ */
/* RNG (common.glsl:40-67)                                                                          */

uint32_t vro_tea(uint32_t val0, uint32_t val1, uint32_t n_rounds) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < n_rounds; ++n) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xA341316Cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4u);
        v1 += ((v0 << 4) + 0xAD90777Du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761Eu);
    }
    return v0;
}

float vro_rng(uint32_t* previous) {
    *previous = *previous * 1664525u + 1013904223u;
    return (float)(*previous & 0x00FFFFFFu) / (float)0x01000000u;
}

void vro_rng_stream(uint32_t seed, int n, float* out, uint32_t* states_out) {
    for (int i = 0; i < n; ++i) {
        out[i] = vro_rng(&seed);
        if (states_out) states_out[i] = seed;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* half floats (glm/detail/type_half.inl:105-238): round-half-UP on the magnitude, not ties-to-even  */

uint16_t vro_to_half(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    const int32_t bits = (int32_t)u;
    const int32_t sign = (bits >> 16) & 0x8000;
    int32_t expo = ((bits >> 23) & 0xff) - (127 - 15);
    int32_t mant = bits & 0x007fffff;
    if (expo <= 0) {
        if (expo < -10) return (uint16_t)sign;               /* flushes to signed zero */
        mant = (mant | 0x00800000) >> (1 - expo);             /* denormal half */
        if (mant & 0x1000) mant += 0x2000;
        return (uint16_t)(sign | (mant >> 13));
    }
    if (expo == 0xff - (127 - 15)) {
        if (mant == 0) return (uint16_t)(sign | 0x7c00);      /* inf */
        mant >>= 13;
        return (uint16_t)(sign | 0x7c00 | mant | (mant == 0)); /* nan */
    }
    if (mant & 0x1000) {
        mant += 0x2000;
        if (mant & 0x00800000) { mant = 0; expo += 1; }
    }
    if (expo > 30) return (uint16_t)(sign | 0x7c00);          /* overflow -> inf */
    return (uint16_t)(sign | (expo << 10) | (mant >> 13));
}

float vro_from_half(uint16_t h) { /* exact IEEE binary16 -> binary32 (type_half.inl:31-103) */
    const uint32_t s = (uint32_t)(h >> 15) << 31;
    const int e = (h >> 10) & 0x1f;
    const uint32_t m = h & 0x3ff;
    uint32_t u;
    if (e == 0) {
        if (m == 0) u = s;
        else {
            float v = ldexpf((float)m, -24);
            memcpy(&u, &v, 4);
            u |= s;
        }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((uint32_t)(e + 112) << 23) | (m << 13);
    float r;
    memcpy(&r, &u, 4);
    return r;
}

uint32_t vro_encode_range(float lo, float hi) { /* grid_brick.cpp:24-26 */
    /* glm's hdata is a SIGNED short: uint32_t(hdata) sign-extends, so a negative minimum sets all 16
     * upper bits of the word (the majorant half then reads back as 0xffff). Reproduced deliberately. */
    return (uint32_t)(int32_t)(int16_t)vro_to_half(lo) | ((uint32_t)(int32_t)(int16_t)vro_to_half(hi) << 16);
}

/* ------------------------------------------------------------------------------------------------ */
/* DenseGrid (grid_dense.cpp:57-103)                                                                */

void vro_dense_from_float(const float* data, const uint32_t dim[3], uint8_t* out_u8, float out_minmax[2]) {
    const size_t n = (size_t)dim[0] * dim[1] * dim[2];
    float lo = FLT_MAX, hi = FLT_MIN; /* sic: max starts at the smallest POSITIVE float (grid_dense.cpp:61) */
    for (size_t i = 0; i < n; ++i) {
        lo = data[i] < lo ? data[i] : lo; /* std::min(a,b) = b<a ? b : a */
        hi = hi < data[i] ? data[i] : hi; /* std::max(a,b) = a<b ? b : a */
    }
    for (size_t i = 0; i < n; ++i)
        out_u8[i] = (uint8_t)roundf(255 * (data[i] - lo) / (hi - lo)); /* grid_dense.cpp:91 */
    out_minmax[0] = lo;
    out_minmax[1] = hi;
}

float vro_dense_lookup(const uint8_t* vox, const uint32_t dim[3], float vmin, float vmax, uint32_t x, uint32_t y, uint32_t z) {
    if (x >= dim[0] || y >= dim[1] || z >= dim[2]) return 0.f; /* grid_dense.cpp:100 */
    const size_t idx = (size_t)z * dim[0] * dim[1] + (size_t)y * dim[0] + x;
    return vmin + (vox[idx] / 255.f) * (vmax - vmin);
}

/* ------------------------------------------------------------------------------------------------ */
/* BrickGrid (grid_brick.cpp:60-154), serial (raster) allocation order                              */

int vro_brick_dims(const uint32_t dim[3], uint32_t n_bricks[3]) {
    for (int a = 0; a < 3; ++a) {
        /* div_round_up goes through float: ceil(float(num) / float(denom)) (grid_brick.cpp:54-56) */
        const uint32_t b = (uint32_t)ceilf((float)dim[a] / 8.f);
        const uint32_t c = (uint32_t)ceilf((float)b / 8.f);
        n_bricks[a] = (c * 1u) << 3;
        if (n_bricks[a] >= 1024u) return VRB_ERR_TOO_MANY_BRICKS; /* :66-67 */
    }
    return 0;
}

static inline void decode_ptr(uint32_t d, uint32_t p[3]) { /* grid_brick.cpp:39-43 */
    p[0] = (d >> 22) & 1023u; p[1] = (d >> 12) & 1023u; p[2] = (d >> 2) & 1023u;
}

/* Grid::lookup as BrickGrid(const Grid&) sees it: a virtual call with a uvec3 (negative window coordinates wrapped) */
typedef float (*vro_lookup_fn)(const void* src, uint32_t x, uint32_t y, uint32_t z);
typedef struct { const uint8_t* vox; const uint32_t* dim; float vmin, vmax; } vro_dense_src;
static float lookup_dense(const void* p, uint32_t x, uint32_t y, uint32_t z) {
    const vro_dense_src* s = (const vro_dense_src*)p;
    return vro_dense_lookup(s->vox, s->dim, s->vmin, s->vmax, x, y, z);
}
/* any other Grid (e.g. NanoVDBGrid::lookup, grid_nvdb.cpp:64-67): its values tabulated on the padded lattice
 * [-2, 8 nb + 2)^3, the set of voxels the constructor addresses (x fastest, origin (-2, -2, -2)) */
typedef struct { const float* val; uint32_t px, py; } vro_values_src;
static float lookup_values(const void* p, uint32_t x, uint32_t y, uint32_t z) {
    const vro_values_src* s = (const vro_values_src*)p;
    return s->val[((size_t)((int32_t)z + 2) * s->py + (size_t)((int32_t)y + 2)) * s->px + (size_t)((int32_t)x + 2)];
}
static int brick_build_any(vro_lookup_fn lookup, const void* grid, const uint32_t dim[3], vrb_brick_view* out);

int vro_brick_build(const uint8_t* vox, const uint32_t dim[3], float vmin, float vmax, vrb_brick_view* out) {
    const vro_dense_src s = { vox, dim, vmin, vmax };
    return brick_build_any(lookup_dense, &s, dim, out);
}
int vro_brick_build_values(const float* padded_values, const uint32_t extent[3], vrb_brick_view* out) {
    uint32_t nb[3];
    const int st = vro_brick_dims(extent, nb);
    if (st) return st;
    const vro_values_src s = { padded_values, nb[0] * 8 + 4, nb[1] * 8 + 4 };
    return brick_build_any(lookup_values, &s, extent, out);
}

static int brick_build_any(vro_lookup_fn lookup, const void* grid, const uint32_t dim[3], vrb_brick_view* out) {
    uint32_t nb[3];
    const int st = vro_brick_dims(dim, nb);
    if (st) return st;
    memcpy(out->n_bricks, nb, sizeof(nb));
    const size_t n_total = (size_t)nb[0] * nb[1] * nb[2];
    const uint32_t ax = nb[0] * 8, ay = nb[1] * 8;
    uint64_t counter = 0;
    for (uint32_t bz = 0; bz < nb[2]; ++bz) for (uint32_t by = 0; by < nb[1]; ++by) for (uint32_t bx = 0; bx < nb[0]; ++bx) {
        const size_t bi = ((size_t)bz * nb[1] + by) * nb[0] + bx;
        out->indirection[bi] = 0;
        float lmin = FLT_MAX, lmax = -FLT_MAX;
        for (int z = -2; z < 10; ++z) for (int y = -2; y < 10; ++y) for (int x = -2; x < 10; ++x) {
            /* negative coordinates wrap to huge unsigned values -> DenseGrid::lookup returns 0 */
            const float v = lookup(grid, (uint32_t)((int)(bx * 8) + x), (uint32_t)((int)(by * 8) + y), (uint32_t)((int)(bz * 8) + z));
            lmin = v < lmin ? v : lmin;
            lmax = lmax < v ? v : lmax;
        }
        out->range[bi] = vro_encode_range(lmin, lmax);
        if (lmax == lmin) continue;
        const uint64_t id = counter++;
        const uint32_t px = (uint32_t)(id % nb[0]), py = (uint32_t)((id / nb[0]) % nb[1]), pz = (uint32_t)(id / ((uint64_t)nb[0] * nb[1]));
        out->indirection[bi] = (px << 22) | (py << 12) | (pz << 2);
        const float lo = vro_from_half((uint16_t)(out->range[bi] & 0xffff)), hi = vro_from_half((uint16_t)(out->range[bi] >> 16));
        for (uint32_t z = 0; z < 8; ++z) for (uint32_t y = 0; y < 8; ++y) for (uint32_t x = 0; x < 8; ++x) {
            const float v = lookup(grid, bx * 8 + x, by * 8 + y, bz * 8 + z);
            const float vn = gl_clamp((v - lo) / (hi - lo), 0.f, 1.f); /* glm::clamp = min(max(x,lo),hi) */
            /* uint8_t(std::round(255 * NaN)) is UB in the reference (fp16 range collapse, 0/0); x86 yields 0 */
            const float r = roundf(255 * vn);
            const uint8_t q = isnan(r) ? 0 : (uint8_t)r;
            out->atlas[((size_t)(pz * 8 + z) * ay + (py * 8 + y)) * ax + (px * 8 + x)] = q;
        }
    }
    out->brick_count = counter;
    out->atlas_dim[0] = ax;
    out->atlas_dim[1] = ay;
    out->atlas_dim[2] = 8u * (uint32_t)roundf(ceilf((float)counter / (float)(nb[0] * nb[1]))); /* :112 */
    /* min/max mips (:114-141) */
    const uint32_t* src = out->range;
    uint32_t sd[3] = { nb[0], nb[1], nb[2] };
    for (int i = 0; i < 3; ++i) {
        const uint32_t md[3] = { nb[0] >> (i + 1), nb[1] >> (i + 1), nb[2] >> (i + 1) };
        uint32_t* dst = out->range_mips[i];
        for (uint32_t bz = 0; bz < md[2]; ++bz) for (uint32_t by = 0; by < md[1]; ++by) for (uint32_t bx = 0; bx < md[0]; ++bx) {
            float rmin = FLT_MAX, rmax = -FLT_MAX;
            for (uint32_t z = 0; z < 2; ++z) for (uint32_t y = 0; y < 2; ++y) for (uint32_t x = 0; x < 2; ++x) {
                const uint32_t w = src[((size_t)(2 * bz + z) * sd[1] + (2 * by + y)) * sd[0] + (2 * bx + x)];
                const float lo = vro_from_half((uint16_t)(w & 0xffff)), hi = vro_from_half((uint16_t)(w >> 16));
                rmin = lo < rmin ? lo : rmin;
                rmax = rmax < hi ? hi : rmax;
            }
            dst[((size_t)bz * md[1] + by) * md[0] + bx] = vro_encode_range(rmin, rmax);
        }
        src = dst;
        sd[0] = md[0]; sd[1] = md[1]; sd[2] = md[2];
    }
    (void)n_total;
    return 0;
}

float vro_brick_lookup(const vro_grid* g, uint32_t x, uint32_t y, uint32_t z) { /* grid_brick.cpp:148-154 */
    const size_t bi = ((size_t)(z >> 3) * g->n_bricks[1] + (y >> 3)) * g->n_bricks[0] + (x >> 3);
    uint32_t p[3];
    decode_ptr(g->indirection[bi], p);
    const float lo = vro_from_half((uint16_t)(g->range[bi] & 0xffff)), hi = vro_from_half((uint16_t)(g->range[bi] >> 16));
    const uint8_t d = g->atlas[((size_t)((p[2] << 3) + (z & 7)) * g->atlas_dim[1] + ((p[1] << 3) + (y & 7))) * g->atlas_dim[0] + ((p[0] << 3) + (x & 7))];
    return lo + d * (1.f / 255.f) * (hi - lo);
}

/* ------------------------------------------------------------------------------------------------ */
/* transfer function upload (transferfunc.cpp:33-58); returns 1 if the CDF rewrite was applied       */

int vro_lut_upload(const float* rgba, uint32_t n, float* out) {
    int needs_cdf = 0;
    for (uint32_t i = 1; i < n; ++i)
        if (rgba[(i - 1) * 4 + 3] > rgba[i * 4 + 3]) { needs_cdf = 1; break; }
    memcpy(out, rgba, (size_t)n * 16);
    if (!needs_cdf) return 0;
    for (uint32_t i = 1; i < n; ++i) out[i * 4 + 3] += out[(i - 1) * 4 + 3];
    const float integral = out[(n - 1) * 4 + 3];
    for (uint32_t i = 0; i < n; ++i)
        out[i * 4 + 3] = integral <= 0.f ? (i + 1) / (float)n : out[i * 4 + 3] / integral;
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* environment: bilinear REPEAT fetch, importance map + box pyramid                                  */

#define IMP_DIM 512      /* environment.cpp:6 */
#define IMP_LEVELS 10    /* levels 0..9; env_imp_base_mip = 9 (renderer.cpp:131) */

size_t vro_env_pyramid_floats(void) {
    size_t n = 0;
    for (int l = 0; l < IMP_LEVELS; ++l) n += (size_t)(IMP_DIM >> l) * (IMP_DIM >> l);
    return n;
}
static inline size_t imp_offset(int level) {
    size_t n = 0;
    for (int l = 0; l < level; ++l) n += (size_t)(IMP_DIM >> l) * (IMP_DIM >> l);
    return n;
}

static inline int wrap_repeat(int i, int n) { int r = i % n; return r < 0 ? r + n : r; }

static v3 env_texture(const float* rgb, int w, int h, float u, float v) { /* texture(env_envmap, uv), LOD 0 */
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    const int x0 = wrap_repeat(f2i(fx), w), y0 = wrap_repeat(f2i(fy), h);
    const int x1 = wrap_repeat(x0 + 1, w), y1 = wrap_repeat(y0 + 1, h);
    const float* t00 = rgb + ((size_t)y0 * w + x0) * 3;
    const float* t10 = rgb + ((size_t)y0 * w + x1) * 3;
    const float* t01 = rgb + ((size_t)y1 * w + x0) * 3;
    const float* t11 = rgb + ((size_t)y1 * w + x1) * 3;
    v3 r;
    r.x = gl_mix(gl_mix(t00[0], t10[0], ax), gl_mix(t01[0], t11[0], ax), ay);
    r.y = gl_mix(gl_mix(t00[1], t10[1], ax), gl_mix(t01[1], t11[1], ax), ay);
    r.z = gl_mix(gl_mix(t00[2], t10[2], ax), gl_mix(t01[2], t11[2], ax), ay);
    return r;
}

void vro_env_build(const float* rgb, int w, int h, float* pyr) { /* env_setup.glsl:18-34, environment.cpp:19-32 */
    const int ns = 8; /* sqrt(SAMPLES = 64) */
    const float inv_samples = 1.f / (float)(ns * ns);
    const float out_samples = (float)(IMP_DIM * ns);
#pragma omp parallel for schedule(static)
    for (int py = 0; py < IMP_DIM; ++py)
        for (int px = 0; px < IMP_DIM; ++px) {
            float importance = 0.f;
            for (int y = 0; y < ns; ++y)
                for (int x = 0; x < ns; ++x) {
                    const float u = ((float)(px * ns) + ((float)x + .5f)) / out_samples;
                    const float v = ((float)(py * ns) + ((float)y + .5f)) / out_samples;
                    importance += luma(env_texture(rgb, w, h, u, v));
                }
            pyr[(size_t)py * IMP_DIM + px] = importance * inv_samples;
        }
    for (int l = 1; l < IMP_LEVELS; ++l) {
        const int d = IMP_DIM >> l, sd = d * 2;
        const float* src = pyr + imp_offset(l - 1);
        float* dst = pyr + imp_offset(l);
        for (int y = 0; y < d; ++y)
            for (int x = 0; x < d; ++x) {
                const float a = src[(size_t)(2 * y) * sd + 2 * x], b = src[(size_t)(2 * y) * sd + 2 * x + 1];
                const float c = src[(size_t)(2 * y + 1) * sd + 2 * x], e = src[(size_t)(2 * y + 1) * sd + 2 * x + 1];
                dst[(size_t)y * d + x] = 0.25f * ((a + b) + (c + e));
            }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* path tracer state                                                                                */

typedef struct {
    const vro_scene* sc;
    const vrb_params* p;
    uint64_t n_maj, n_dens, n_emis, n_nee, n_env, n_real;
} tctx;

static inline float imp_fetch(const tctx* c, int x, int y, int level) { /* texelFetch(env_impmap, ...) */
    const int d = IMP_DIM >> level;
    if (x < 0 || y < 0 || x >= d || y >= d) return 0.f;
    return c->sc->impmap[imp_offset(level) + (size_t)y * d + x];
}

/* common.glsl:93-98 */
static v3 lookup_environment(const tctx* c, v3 dir) {
    const v3 idir = m3mul(c->p->env_inv_transform, dir);
    const float u = atan2f(idir.z, idir.x) / (2 * PI_F) + 0.5f;
    const float v = 1.f - acosf(gl_clamp(idir.y, -1.f, 1.f)) / PI_F;
    return scale3(env_texture(c->sc->env_rgb, c->sc->env_w, c->sc->env_h, u, v), c->p->env_strength);
}

/* common.glsl:100-146 */
static v4 sample_environment(const tctx* c, v2 rnd, v3* w_i) {
    int posx = 0, posy = 0;
    v2 p = rnd;
    for (int mip = 9 - 1; mip >= 0; mip--) {
        posx *= 2; posy *= 2;
        float w[4];
        w[0] = imp_fetch(c, posx, posy, mip);
        w[1] = imp_fetch(c, posx + 1, posy, mip);
        w[2] = imp_fetch(c, posx, posy + 1, mip);
        w[3] = imp_fetch(c, posx + 1, posy + 1, mip);
        float q[2];
        q[0] = w[0] + w[2];
        q[1] = w[1] + w[3];
        int off_x;
        const float d = q[0] / gl_max(1e-8f, q[0] + q[1]);
        if (p.x < d) { off_x = 0; p.x = p.x / d; }
        else { off_x = 1; p.x = (p.x - d) / (1.f - d); }
        posx += off_x;
        const float e = w[off_x] / q[off_x];
        if (p.y < e) { p.y = p.y / e; }
        else { posy += 1; p.y = (p.y - e) / (1.f - e); }
    }
    const float inv_dim = 1.f / (float)IMP_DIM;
    const float uvx = ((float)posx + p.x) * inv_dim, uvy = ((float)posy + p.y) * inv_dim;
    const float theta = saturate(1.f - uvy) * PI_F;
    const float phi = (saturate(uvx) * 2.f - 1.f) * PI_F;
    const float sin_t = sinf(theta);
    *w_i = m3mul(c->p->env_transform, V3(sin_t * cosf(phi), cosf(theta), sin_t * sinf(phi)));
    const v3 Le = scale3(env_texture(c->sc->env_rgb, c->sc->env_w, c->sc->env_h, uvx, uvy), c->p->env_strength);
    const float avg_w = imp_fetch(c, 0, 0, 9);
    const float pdf = imp_fetch(c, posx, posy, 0) / avg_w;
    v4 r = { Le.x, Le.y, Le.z, pdf * INV_4PI };
    return r;
}

/* common.glsl:148-152 */
static float pdf_environment(const tctx* c, v3 dir) {
    const float avg_w = imp_fetch(c, 0, 0, 9);
    const float pdf = luma(lookup_environment(c, dir)) / avg_w;
    return pdf * INV_4PI;
}

/* common.glsl:157-165 */
static int intersect_box(v3 pos, v3 dir, v3 bb_min, v3 bb_max, v2* near_far) {
    const v3 inv_dir = V3(1.f / dir.x, 1.f / dir.y, 1.f / dir.z);
    const v3 lo = mul3(sub3(bb_min, pos), inv_dir);
    const v3 hi = mul3(sub3(bb_max, pos), inv_dir);
    const v3 tmin = V3(gl_min(lo.x, hi.x), gl_min(lo.y, hi.y), gl_min(lo.z, hi.z));
    const v3 tmax = V3(gl_max(lo.x, hi.x), gl_max(lo.y, hi.y), gl_max(lo.z, hi.z));
    near_far->x = gl_max(0.f, gl_max(tmin.x, gl_max(tmin.y, tmin.z)));
    near_far->y = gl_min(tmax.x, gl_min(tmax.y, tmax.z));
    return near_far->x <= near_far->y;
}

/* common.glsl:172-175 */
static float phase_hg(float cos_t, float g) {
    const float denom = 1 + sqr(g) + 2 * g * cos_t;
    return INV_4PI * (1 - sqr(g)) / (denom * sqrtf(denom));
}

/* common.glsl:184-190 */
static v3 sample_phase_hg(v3 dir, float g, v2 s) {
    const float cos_t = fabsf(g) < 1e-4f ? 1.f - 2.f * s.x : (1 + sqr(g) - sqr((1 - sqr(g)) / (1 - g + 2 * g * s.x))) / (2 * g);
    const float sin_t = sqrtf(gl_max(0.f, 1.f - sqr(cos_t)));
    const float phi = 2.f * PI_F * s.y;
    return align_to(dir, V3(sin_t * cosf(phi), sin_t * sinf(phi), cos_t));
}

/* common.glsl:203-212 */
static v4 tf_lookup(const tctx* c, float d) {
    const float tc = gl_clamp((d - c->p->tf_window_left) / c->p->tf_window_width, 0.0f, 1.0f - 1e-6f);
    const uint32_t n = c->sc->tf_size;
    const float s = tc * (float)n;
    const int idx = f2i(floorf(s));
    const float f = gl_fract(s);
    uint32_t idx1 = (uint32_t)(idx + 1);
    if (idx1 > n - 1) idx1 = n - 1;
    const float* a = c->sc->tf_lut + (size_t)idx * 4;
    const float* b = c->sc->tf_lut + (size_t)idx1 * 4;
    v4 r = { gl_mix(a[0], b[0], f), gl_mix(a[1], b[1], f), gl_mix(a[2], b[2], f), gl_mix(a[3], b[3], f) };
    return r;
}

/* common.glsl:221-244 */
static i3 stochastic_tricubic_filter(v3 ipos, uint32_t* seed) {
    const v3 q = V3(ipos.x - 0.5f, ipos.y - 0.5f, ipos.z - 0.5f);
    const i3 ii = { f2i(floorf(q.x)), f2i(floorf(q.y)), f2i(floorf(q.z)) };
    const float t[3] = { q.x - (float)ii.x, q.y - (float)ii.y, q.z - (float)ii.z };
    int idx[3] = { 0, 0, 0 };
    float sum[3], w[3], t2[3];
    for (int a = 0; a < 3; ++a) {
        t2[a] = t[a] * t[a];
        w[a] = (1.f / 6.f) * (-t[a] * t2[a] + 3 * t2[a] - 3 * t[a] + 1);
        sum[a] = w[a];
    }
    float r[3];
    /* second tap */
    for (int a = 0; a < 3; ++a) { w[a] = (1.f / 6.f) * (3 * t[a] * t2[a] - 6 * t2[a] + 4); sum[a] = w[a] + sum[a]; }
    r[0] = vro_rng(seed); r[1] = vro_rng(seed); r[2] = vro_rng(seed);
    for (int a = 0; a < 3; ++a) if (r[a] < w[a] / gl_max(1e-3f, sum[a])) idx[a] = 1;
    /* third tap */
    for (int a = 0; a < 3; ++a) { w[a] = (1.f / 6.f) * (-3 * t[a] * t2[a] + 3 * t2[a] + 3 * t[a] + 1); sum[a] = w[a] + sum[a]; }
    r[0] = vro_rng(seed); r[1] = vro_rng(seed); r[2] = vro_rng(seed);
    for (int a = 0; a < 3; ++a) if (r[a] < w[a] / gl_max(1e-3f, sum[a])) idx[a] = 2;
    /* fourth tap */
    for (int a = 0; a < 3; ++a) { w[a] = (1.f / 6.f) * t[a] * t2[a]; sum[a] = w[a] + sum[a]; }
    r[0] = vro_rng(seed); r[1] = vro_rng(seed); r[2] = vro_rng(seed);
    for (int a = 0; a < 3; ++a) if (r[a] < w[a] / gl_max(1e-3f, sum[a])) idx[a] = 3;
    const i3 out = { ii.x + idx[0] - 1, ii.y + idx[1] - 1, ii.z + idx[2] - 1 };
    return out;
}

/* texelFetch on the brick textures; out of bounds -> 0 */
static inline int brick_index(const vro_grid* g, int bx, int by, int bz, int mip, size_t* out) {
    const int nx = (int)(g->n_bricks[0] >> mip), ny = (int)(g->n_bricks[1] >> mip), nz = (int)(g->n_bricks[2] >> mip);
    if (bx < 0 || by < 0 || bz < 0 || bx >= nx || by >= ny || bz >= nz) return 0;
    *out = ((size_t)bz * ny + by) * nx + bx;
    return 1;
}

/* common.glsl:268-275 / :314-321 */
static float lookup_brick_value(const vro_grid* g, i3 ii) {
    size_t bi;
    if (!brick_index(g, ii.x >> 3, ii.y >> 3, ii.z >> 3, 0, &bi)) return 0.f; /* ptr=0, range=(0,0) -> 0 + u*(0-0) */
    uint32_t ptr[3];
    decode_ptr(g->indirection[bi], ptr);
    const uint32_t rw = g->range[bi];
    const float lo = vro_from_half((uint16_t)(rw & 0xffff)), hi = vro_from_half((uint16_t)(rw >> 16));
    const uint32_t ax = (ptr[0] << 3) + (uint32_t)(ii.x & 7), ay = (ptr[1] << 3) + (uint32_t)(ii.y & 7), az = (ptr[2] << 3) + (uint32_t)(ii.z & 7);
    float unorm = 0.f;
    if (ax < g->atlas_dim[0] && ay < g->atlas_dim[1] && az < g->atlas_dim[2])
        unorm = (float)g->atlas[((size_t)az * g->atlas_dim[1] + ay) * g->atlas_dim[0] + ax] / 255.f;
    return lo + unorm * (hi - lo);
}

/* common.glsl:278-281 */
static float lookup_majorant(tctx* c, v3 ipos, int mip) {
    c->n_maj++;
    const vro_grid* g = &c->sc->density;
    const int bx = f2i(floorf(ipos.x)) >> (3 + mip), by = f2i(floorf(ipos.y)) >> (3 + mip), bz = f2i(floorf(ipos.z)) >> (3 + mip);
    size_t bi;
    if (!brick_index(g, bx, by, bz, mip, &bi)) return c->p->vol_density_scale * 0.f;
    const uint32_t rw = mip == 0 ? g->range[bi] : g->range_mips[mip - 1][bi];
    return c->p->vol_density_scale * vro_from_half((uint16_t)(rw >> 16));
}

static inline i3 floor3(v3 p) { i3 r = { f2i(floorf(p.x)), f2i(floorf(p.y)), f2i(floorf(p.z)) }; return r; }

/* common.glsl:289-297 */
static float lookup_density_trilinear(tctx* c, v3 ipos) {
    const v3 q = V3(ipos.x - 0.5f, ipos.y - 0.5f, ipos.z - 0.5f);
    const v3 f = V3(gl_fract(q.x), gl_fract(q.y), gl_fract(q.z));
    const i3 ii = floor3(q);
    const vro_grid* g = &c->sc->density;
#define TAP(dx, dy, dz) lookup_brick_value(g, (i3){ ii.x + dx, ii.y + dy, ii.z + dz })
    const float lx0 = gl_mix(TAP(0, 0, 0), TAP(1, 0, 0), f.x);
    const float lx1 = gl_mix(TAP(0, 1, 0), TAP(1, 1, 0), f.x);
    const float hx0 = gl_mix(TAP(0, 0, 1), TAP(1, 0, 1), f.x);
    const float hx1 = gl_mix(TAP(0, 1, 1), TAP(1, 1, 1), f.x);
#undef TAP
    return c->p->vol_density_scale * gl_mix(gl_mix(lx0, lx1, f.y), gl_mix(hx0, hx1, f.y), f.z);
}

/* common.glsl:300-304 */
static float lookup_density_stochastic(tctx* c, v3 ipos, uint32_t* seed) {
    const i3 tap = stochastic_tricubic_filter(ipos, seed);
    /* lookup_density(vec3(tap)) -> floor(float(int)) is the identity */
    return c->p->vol_density_scale * lookup_brick_value(&c->sc->density, tap);
}

/* common.glsl:324-328 */
static v3 lookup_emission(tctx* c, v3 ipos, uint32_t* seed) {
    float m[16];
    m4mul(c->p->vol_emission_inv_transform, c->p->vol_density_transform, m);
    const v3 ipos_e = m4mul_xyz(m, ipos, 1.f);
    const i3 tap = stochastic_tricubic_filter(ipos_e, seed);
    float t = 0.f; /* unbound samplers return 0 */
    if (c->p->has_emission) { t = lookup_brick_value(&c->sc->emission, tap); c->n_emis++; }
    t = t * c->p->vol_emission_norm;
    const float s = c->p->vol_emission_scale;
    return V3(s * sqr(t), s * sqr(sqr(t)), s * sqr(sqr(sqr(t))));
}

/* common.glsl:404-409 */
static float stepDDA(v3 pos, v3 inv_dir, int mip) {
    const float dim = (float)(8 << mip);
    const float ox = inv_dir.x >= 0 ? dim + 0.5f : -0.5f;
    const float oy = inv_dir.y >= 0 ? dim + 0.5f : -0.5f;
    const float oz = inv_dir.z >= 0 ? dim + 0.5f : -0.5f;
    const float tx = (floorf(pos.x * (1.f / dim)) * dim + ox - pos.x) * inv_dir.x;
    const float ty = (floorf(pos.y * (1.f / dim)) * dim + oy - pos.y) * inv_dir.y;
    const float tz = (floorf(pos.z * (1.f / dim)) * dim + oz - pos.z) * inv_dir.z;
    return gl_min(tx, gl_min(ty, tz));
}

static inline int round_mip(float mip) { return (int)rintf(mip); } /* GLSL round(): half-to-even (see header) */

#define MIP_START 3.f
#define MIP_SPEED_UP 0.25f
#define MIP_SPEED_DOWN 2.f

/* common.glsl:412-455 */
static float transmittanceDDA(tctx* c, v3 wpos, v3 wdir, uint32_t* seed) {
    const vrb_params* p = c->p;
    const int TF = p->use_transferfunc;
    v2 nf;
    if (!intersect_box(wpos, wdir, V3(p->vol_bb_min[0], p->vol_bb_min[1], p->vol_bb_min[2]), V3(p->vol_bb_max[0], p->vol_bb_max[1], p->vol_bb_max[2]), &nf)) return 1.f;
    const v3 ipos = m4mul_xyz(p->vol_density_inv_transform, wpos, 1.f);
    const v3 idir = m4mul_xyz(p->vol_density_inv_transform, wdir, 0.f);
    const v3 ri = V3(1.f / idir.x, 1.f / idir.y, 1.f / idir.z);
    float t = nf.x + 1e-6f, Tr = 1.f, tau = -logf(1.f - vro_rng(seed)), mip = MIP_START;
    while (t < nf.y) {
        const v3 curr = add3(ipos, scale3(idir, t));
        float majorant = lookup_majorant(c, curr, round_mip(mip));
        if (TF) majorant = p->vol_majorant * tf_lookup(c, majorant * p->vol_inv_majorant).w;
        const float dt = stepDDA(curr, ri, round_mip(mip));
        t += dt;
        tau -= majorant * dt;
        mip = gl_min(mip + MIP_SPEED_UP, 3.f);
        if (tau > 0) continue;
        t += tau / majorant;
        if (t >= nf.y) break;
        float d;
        c->n_dens++;
        if (TF) {
            const v4 rgba = tf_lookup(c, lookup_density_trilinear(c, add3(ipos, scale3(idir, t))) * p->vol_inv_majorant);
            d = p->vol_majorant * rgba.w;
        } else d = lookup_density_stochastic(c, add3(ipos, scale3(idir, t)), seed);
        if (vro_rng(seed) * majorant < d) {
            Tr *= gl_max(0.f, 1.f - p->vol_majorant / majorant);
            if (Tr < .1f) {
                const float prob = 1 - Tr;
                if (vro_rng(seed) < prob) return 0.f;
                Tr /= 1 - prob;
            }
        }
        tau = -logf(1.f - vro_rng(seed));
        mip = gl_max(0.f, mip - MIP_SPEED_DOWN);
    }
    return Tr;
}

/* common.glsl:458-501 */
static int sample_volumeDDA(tctx* c, v3 wpos, v3 wdir, float* t_out, v3* throughput, v3* Le, uint32_t* seed) {
    const vrb_params* p = c->p;
    const int TF = p->use_transferfunc;
    v2 nf;
    if (!intersect_box(wpos, wdir, V3(p->vol_bb_min[0], p->vol_bb_min[1], p->vol_bb_min[2]), V3(p->vol_bb_max[0], p->vol_bb_max[1], p->vol_bb_max[2]), &nf)) return 0;
    const v3 ipos = m4mul_xyz(p->vol_density_inv_transform, wpos, 1.f);
    const v3 idir = m4mul_xyz(p->vol_density_inv_transform, wdir, 0.f);
    const v3 ri = V3(1.f / idir.x, 1.f / idir.y, 1.f / idir.z);
    const v3 albedo = V3(p->vol_albedo[0], p->vol_albedo[1], p->vol_albedo[2]);
    float t = nf.x + 1e-6f;
    float tau = -logf(1.f - vro_rng(seed)), mip = MIP_START;
    *t_out = t;
    while (t < nf.y) {
        const v3 curr = add3(ipos, scale3(idir, t));
        float majorant = lookup_majorant(c, curr, round_mip(mip));
        if (TF) majorant = p->vol_majorant * tf_lookup(c, majorant * p->vol_inv_majorant).w;
        const float dt = stepDDA(curr, ri, round_mip(mip));
        t += dt;
        tau -= majorant * dt;
        mip = gl_min(mip + MIP_SPEED_UP, 3.f);
        if (tau > 0) continue;
        t += tau / majorant;
        if (t >= nf.y) break;
        float d;
        v4 rgba = { 1, 1, 1, 1 };
        c->n_dens++;
        const v3 at = add3(ipos, scale3(idir, t));
        if (TF) {
            rgba = tf_lookup(c, lookup_density_trilinear(c, at) * p->vol_inv_majorant);
            d = p->vol_majorant * rgba.w;
        } else d = lookup_density_stochastic(c, at, seed);
        {
            const v3 em = lookup_emission(c, at, seed);
            /* Le += throughput * (1 - albedo) * emission * d * inv_majorant (left-to-right) */
            const v3 one_minus = V3(1.f - albedo.x, 1.f - albedo.y, 1.f - albedo.z);
            v3 term = mul3(mul3(*throughput, one_minus), em);
            term = scale3(scale3(term, d), p->vol_inv_majorant);
            *Le = add3(*Le, term);
        }
        if (vro_rng(seed) * majorant < d) {
            *throughput = mul3(*throughput, albedo);
            if (TF) *throughput = mul3(*throughput, V3(rgba.x, rgba.y, rgba.z));
            *t_out = t;
            return 1;
        }
        tau = -logf(1.f - vro_rng(seed));
        mip = gl_max(0.f, mip - MIP_SPEED_DOWN);
    }
    *t_out = t;
    return 0;
}

/* common.glsl:599-652 */
static v4 trace_path(tctx* c, v3 pos, v3 dir, uint32_t* seed) {
    const vrb_params* p = c->p;
    v3 L = V3(0, 0, 0), throughput = V3(1, 1, 1);
    int free_path = 1;
    uint32_t n_paths = 0;
    float t = 0.f, f_p = 0.f;
    while (sample_volumeDDA(c, pos, dir, &t, &throughput, &L, seed)) {
        c->n_real++;
        pos = add3(pos, scale3(dir, t));
        v3 w_i;
        v2 r2;
        r2.x = vro_rng(seed); r2.y = vro_rng(seed);
        c->n_nee++;
        const v4 Le_pdf = sample_environment(c, r2, &w_i);
        if (Le_pdf.w > 0) {
            f_p = phase_hg(dot3(V3(-dir.x, -dir.y, -dir.z), w_i), p->vol_phase_g);
            const float mis_weight = p->show_environment > 0 ? power_heuristic(Le_pdf.w, f_p) : 1.f;
            const float Tr = transmittanceDDA(c, pos, w_i, seed);
            /* L += throughput * mis_weight * f_p * Tr * Le_pdf.rgb / Le_pdf.w */
            v3 term = scale3(scale3(scale3(throughput, mis_weight), f_p), Tr);
            term = mul3(term, V3(Le_pdf.x, Le_pdf.y, Le_pdf.z));
            term = V3(term.x / Le_pdf.w, term.y / Le_pdf.w, term.z / Le_pdf.w);
            L = add3(L, term);
        }
        if (++n_paths >= (uint32_t)p->bounces) { free_path = 0; break; }
        const float rr_val = luma(throughput);
        if (rr_val < .1f) {
            const float prob = 1 - rr_val;
            if (vro_rng(seed) < prob) { free_path = 0; break; }
            const float k = 1 - prob;
            throughput = V3(throughput.x / k, throughput.y / k, throughput.z / k);
        }
        v2 ps;
        ps.x = vro_rng(seed); ps.y = vro_rng(seed);
        const v3 scatter_dir = sample_phase_hg(dir, p->vol_phase_g, ps);
        f_p = phase_hg(dot3(V3(-dir.x, -dir.y, -dir.z), scatter_dir), p->vol_phase_g);
        dir = scatter_dir;
    }
    if (free_path && p->show_environment > 0) {
        c->n_env++;
        const v3 Le = lookup_environment(c, dir);
        const float mis_weight = n_paths > 0 ? power_heuristic(f_p, pdf_environment(c, dir)) : 1.f;
        L = add3(L, mul3(scale3(throughput, mis_weight), Le));
    }
    v4 r = { L.x, L.y, L.z, gl_clamp((float)n_paths, 0.f, 1.f) };
    return r;
}

/* common.glsl:76-80 */
static v3 view_dir(const vrb_params* p, int x, int y, v2 pixel_sample) {
    const float w = (float)p->resolution[0], h = (float)p->resolution[1];
    const float px = ((float)x + pixel_sample.x - w * .5f) / h;
    const float py = ((float)y + pixel_sample.y - h * .5f) / h;
    const float z = -.5f / tanf(.5f * PI_F * p->cam_fov / 180.f);
    return normalize3(m3mul(p->cam_transform, normalize3(V3(px, py, z))));
}

int vro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* pathtracer_brick.glsl:23-37 for samples first..first+n-1 */
void vro_trace(const vro_scene* scene, const vrb_params* p, int first_sample, int n_samples, const int tile[4],
               int accum_mode, float* color, vrb_counters* counters, int n_threads) {
    const int W = p->resolution[0], H = p->resolution[1];
    const int x0 = tile ? tile[0] : 0, y0 = tile ? tile[1] : 0, x1 = tile ? tile[2] : W, y1 = tile ? tile[3] : H;
    uint64_t n_maj = 0, n_dens = 0, n_emis = 0, n_nee = 0, n_env = 0, n_real = 0, n_samp = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : n_maj, n_dens, n_emis, n_nee, n_env, n_real, n_samp)
    for (int y = y0; y < y1; ++y) {
        tctx c;
        memset(&c, 0, sizeof(c));
        c.sc = scene;
        c.p = p;
        for (int x = x0; x < x1; ++x) {
            float* px = color + ((size_t)y * W + x) * 4;
            for (int s = first_sample; s < first_sample + n_samples; ++s) {
                uint32_t seed = vro_tea((uint32_t)p->seed * (uint32_t)(y * W + x), (uint32_t)s, 32);
                v2 jit;
                jit.x = vro_rng(&seed); jit.y = vro_rng(&seed);
                const v3 dir = view_dir(p, x, y, jit);
                const v4 L = trace_path(&c, V3(p->cam_pos[0], p->cam_pos[1], p->cam_pos[2]), dir, &seed);
                const float Ls[4] = { sanitize1(L.x), sanitize1(L.y), sanitize1(L.z), sanitize1(L.w) };
                if (accum_mode == VRB_ACCUM_MEAN) {
                    const float a = 1.f / (float)s;
                    for (int k = 0; k < 4; ++k) px[k] = gl_mix(px[k], Ls[k], a);
                } else
                    for (int k = 0; k < 4; ++k) px[k] += Ls[k];
                n_samp++;
            }
        }
        n_maj += c.n_maj; n_dens += c.n_dens; n_emis += c.n_emis; n_nee += c.n_nee; n_env += c.n_env; n_real += c.n_real;
    }
    if (counters) {
        counters->n_samples += n_samp; counters->n_maj += n_maj; counters->n_dens += n_dens; counters->n_emis += n_emis;
        counters->n_nee += n_nee; counters->n_env += n_env; counters->n_real += n_real;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* deterministic transmittance-only mode ("T1", defined by this build): fp64 exact voxel DDA         */

void vro_trace_deterministic(const vro_scene* scene, const vrb_params* p, float* color, int n_threads) {
    const int W = p->resolution[0], H = p->resolution[1];
    const vro_grid* g = &scene->density;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int y = 0; y < H; ++y) {
        tctx c;
        memset(&c, 0, sizeof(c));
        c.sc = scene;
        c.p = p;
        for (int x = 0; x < W; ++x) {
            const v2 centre = { .5f, .5f };
            const v3 dirf = view_dir(p, x, y, centre);
            const double pos[3] = { p->cam_pos[0], p->cam_pos[1], p->cam_pos[2] };
            const double dir[3] = { dirf.x, dirf.y, dirf.z };
            /* slab test in fp64 */
            double tn = 0.0, tf = INFINITY;
            int hit = 1;
            for (int a = 0; a < 3; ++a) {
                const double inv = 1.0 / dir[a];
                double lo = ((double)p->vol_bb_min[a] - pos[a]) * inv, hi = ((double)p->vol_bb_max[a] - pos[a]) * inv;
                if (hi < lo) { const double tmp = lo; lo = hi; hi = tmp; }
                if (lo > tn) tn = lo;
                if (hi < tf) tf = hi;
            }
            if (!(tn <= tf)) hit = 0;
            double tau = 0.0;
            if (hit) {
                const float* M = p->vol_density_inv_transform;
                double ip[3], id[3];
                for (int r = 0; r < 3; ++r) {
                    ip[r] = (double)M[0 + r] * pos[0] + (double)M[4 + r] * pos[1] + (double)M[8 + r] * pos[2] + (double)M[12 + r];
                    id[r] = (double)M[0 + r] * dir[0] + (double)M[4 + r] * dir[1] + (double)M[8 + r] * dir[2];
                }
                double t = tn;
                int guard = 0;
                while (t < tf && guard++ < (1 << 22)) {
                    /* voxel containing the midpoint of the next tiny step decides the cell */
                    double tnext = tf;
                    int vox[3];
                    const double tp = t + 1e-9 * (1.0 + fabs(t));
                    for (int a = 0; a < 3; ++a) {
                        const double q = ip[a] + tp * id[a];
                        vox[a] = (int)floor(q);
                        if (id[a] > 0) { const double tb = ((double)(vox[a] + 1) - ip[a]) / id[a]; if (tb < tnext) tnext = tb; }
                        else if (id[a] < 0) { const double tb = ((double)vox[a] - ip[a]) / id[a]; if (tb < tnext) tnext = tb; }
                    }
                    if (tnext <= t) tnext = tp; /* guarantee progress */
                    const i3 vi = { vox[0], vox[1], vox[2] };
                    const double sigma = (double)lookup_brick_value(g, vi);
                    tau += sigma * (tnext - t);
                    t = tnext;
                }
            }
            const double Tr = exp(-(double)p->vol_density_scale * tau);
            const v3 Le = lookup_environment(&c, dirf);
            float* px = color + ((size_t)y * W + x) * 4;
            px[0] = (float)(Tr * Le.x); px[1] = (float)(Tr * Le.y); px[2] = (float)(Tr * Le.z); px[3] = (float)(1.0 - Tr);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* tonemap (tonemap.glsl:13-36, tonemap.fs:10-28)                                                    */

static inline float hable(float x) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
}
static inline float hable_tonemap(float x, float exposure) { return hable(exposure * x) / hable(11.2f); }

void vro_tonemap_inplace(float* color, int w, int h, float exposure, float gamma) {
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        float* px = color + i * 4;
        for (int k = 0; k < 3; ++k) px[k] = sanitize1(powf(hable_tonemap(px[k], exposure), 1.f / gamma));
        px[3] = sanitize1(px[3]);
    }
}

static inline uint8_t to_unorm8(float x) { /* GL float -> unorm8: clamp, *255, round to nearest */
    if (isnan(x)) return 0;
    const float c = x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
    return (uint8_t)rintf(c * 255.f);
}

void vro_draw(const float* color, int w, int h, float exposure, float gamma, int tonemapping, uint8_t* rgba8) {
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const float* px = color + i * 4;
        for (int k = 0; k < 3; ++k)
            rgba8[i * 4 + k] = to_unorm8(tonemapping ? powf(hable_tonemap(px[k], exposure), 1.f / gamma) : px[k]);
        rgba8[i * 4 + 3] = to_unorm8(px[3]);
    }
}

void vro_color_to_ldr(const float* color, int w, int h, uint8_t* rgba8) {
    for (size_t i = 0; i < (size_t)w * h * 4; ++i) rgba8[i] = to_unorm8(color[i]);
}
