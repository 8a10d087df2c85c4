/*
 * glsl_shim.h -- TEST INFRASTRUCTURE ONLY.
 *
 * The environment in which the reference's UNMODIFIED GLSL (rewritten lexically by glsl2cpp.py) compiles as C++:
 *   - types and built-in functions = glm 0.9.9.8 of the reference's own submodule (submodules/cppgl/subtrees/glm),
 *     which restates the GLSL specification's definitions (mix = x*(1-a)+y*a, clamp = min(max()), normalize =
 *     v*inversesqrt(dot), ...), with GLM_FORCE_SWIZZLE for `.rgb` / `.xyz`;
 *   - the implicit int -> float promotions of GLSL that C++ templates do not deduce (mixed operators below);
 *   - the GL fixed-function behaviour that NO source states (SURVEY 8(c) "parity unpinned" list), pinned here exactly as
 *     DESIGN.md states it:
 *       texture(sampler2D)  bilinear at LOD 0 (no derivatives in a compute shader), texel centres at (i+.5)/N,
 *                           GL_REPEAT on both axes (cppgl/texture.cpp:45-46), fp32 weights, nested mix
 *       texelFetch          exact texel; out of bounds -> 0 (robust buffer access)
 *       GL_R8 atlas         unorm8 -> float = u8 / 255.f (GL 4.5 spec eq. 2.1); GL_COMPRESSED_RED pinned to uncompressed
 *       GL_RG16F range      half -> float by glm::detail::toFloat32 (exact); .x = low half, .y = high half
 *       GL_RGB10_A2UI       x = bits 31..22, y = 21..12, z = 11..2, w = 1..0 (GL_UNSIGNED_INT_10_10_10_2, renderer.cpp:165-167)
 *       acos(x), |x| > 1    argument clamped to [-1, 1] (the spec says "undefined"; normalize() can return 1 + 1 ulp)
 *       round()             half-to-even (the spec leaves .5 to the implementation; Mesa and NVIDIA lower to roundEven)
 */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

/* glm enables its OPERATOR swizzles (`v.rgb` as a member, assignable) only when it believes the compiler has the
 * MS anonymous-struct extension (detail/setup.hpp:75-81, 459); gcc has it, so claim it. This changes the storage union of
 * vec2/3/4 only: GLM_ARCH keeps no SIMD bit, i.e. no SIMD/intrinsic arithmetic path is switched on. */
#ifndef _MSC_EXTENSIONS
#define _MSC_EXTENSIONS 1
#define GLSL_SHIM_UNDEF_MSC_EXTENSIONS
#endif
#define GLM_FORCE_SWIZZLE
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/detail/type_half.hpp>
#ifdef GLSL_SHIM_UNDEF_MSC_EXTENSIONS
#undef _MSC_EXTENSIONS
#endif
static_assert((GLM_ARCH & GLM_ARCH_SIMD_BIT) == 0, "glm must not use an intrinsic arithmetic path here");
static_assert(GLM_CONFIG_SWIZZLE == GLM_SWIZZLE_OPERATOR, "operator swizzles expected");

namespace glsl {
using namespace glm;
using glm::uint;

/* ---- GLSL implicit conversions (int / uint / double-literal operand next to a float vector) -------------------- */
#define GLSL_MIXED_SCALAR(V)                                                                              \
    inline V operator+(const V& a, int b) { return a + float(b); }                                        \
    inline V operator-(const V& a, int b) { return a - float(b); }                                        \
    inline V operator*(const V& a, int b) { return a * float(b); }                                        \
    inline V operator/(const V& a, int b) { return a / float(b); }                                        \
    inline V operator+(int a, const V& b) { return float(a) + b; }                                        \
    inline V operator-(int a, const V& b) { return float(a) - b; }                                        \
    inline V operator*(int a, const V& b) { return float(a) * b; }                                        \
    inline V operator/(int a, const V& b) { return float(a) / b; }
GLSL_MIXED_SCALAR(vec2)
GLSL_MIXED_SCALAR(vec3)
GLSL_MIXED_SCALAR(vec4)

#define GLSL_MIXED_VECTOR(IV, V)                                                                          \
    inline V operator+(const IV& a, const V& b) { return V(a) + b; }                                      \
    inline V operator-(const IV& a, const V& b) { return V(a) - b; }                                      \
    inline V operator*(const IV& a, const V& b) { return V(a) * b; }                                      \
    inline V operator/(const IV& a, const V& b) { return V(a) / b; }                                      \
    inline V operator+(const V& a, const IV& b) { return a + V(b); }                                      \
    inline V operator-(const V& a, const IV& b) { return a - V(b); }                                      \
    inline V operator*(const V& a, const IV& b) { return a * V(b); }                                      \
    inline V operator/(const V& a, const IV& b) { return a / V(b); }                                      \
    inline V operator*(const IV& a, float b) { return V(a) * b; }                                         \
    inline V operator*(float a, const IV& b) { return a * V(b); }
GLSL_MIXED_VECTOR(ivec2, vec2)
GLSL_MIXED_VECTOR(ivec3, vec3)

/* swizzle on the right of a compound assignment (`throughput *= rgba.rgb`) */
template <int E0, int E1, int E2>
inline vec3& operator*=(vec3& a, const glm::detail::_swizzle<3, float, glm::defaultp, E0, E1, E2, -1>& b) { return a *= vec3(b); }

inline uvec3 operator<<(const uvec3& a, int b) { return a << uint(b); }                  /* ptr << 3 */

/* names overloaded below would otherwise hide glm's for scalar arguments */
using glm::clamp;
using glm::min;
using glm::max;
inline float clamp(uint x, float lo, float hi) { return glm::clamp(float(x), lo, hi); }  /* clamp(n_paths, 0.f, 1.f) */
inline uint min(int a, uint b) { return glm::min(uint(a), b); }                          /* min(idx + 1, tf_size - 1) */
inline float acos(float x) { return std::acos(glm::clamp(x, -1.f, 1.f)); }               /* pinned, see header */
inline float round(float x) { return glm::roundEven(x); }                               /* pinned, see header */

/* ---- images and samplers --------------------------------------------------------------------------------------- */
struct image2D { float* data = nullptr; int w = 0, h = 0, channels = 4; };
inline vec4 imageLoad(const image2D& im, const ivec2& p) {
    const float* t = im.data + (size_t(p.y) * im.w + p.x) * im.channels;
    return im.channels == 4 ? vec4(t[0], t[1], t[2], t[3]) : vec4(t[0], 0.f, 0.f, 1.f);
}
inline void imageStore(const image2D& im, const ivec2& p, const vec4& v) {
    float* t = im.data + (size_t(p.y) * im.w + p.x) * im.channels;
    t[0] = v.x;
    if (im.channels == 4) { t[1] = v.y; t[2] = v.z; t[3] = v.w; }
}

/* 2-D float texture with `channels` interleaved floats per texel and a mip chain given as per-level pointers */
struct sampler2D {
    const float* level[16] = {};
    int w = 0, h = 0, channels = 0, levels = 0;
};
inline ivec2 textureSize(const sampler2D& s, int lod) { return ivec2(glm::max(s.w >> lod, 1), glm::max(s.h >> lod, 1)); }
inline vec4 fetch_texel(const sampler2D& s, int x, int y, int lod) {
    const float* t = s.level[lod] + (size_t(y) * glm::max(s.w >> lod, 1) + x) * s.channels;
    return vec4(t[0], s.channels > 1 ? t[1] : 0.f, s.channels > 2 ? t[2] : 0.f, s.channels > 3 ? t[3] : 1.f);
}
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int lod) {
    if (lod < 0 || lod >= s.levels) return vec4(0);
    const ivec2 d = textureSize(s, lod);
    if (p.x < 0 || p.y < 0 || p.x >= d.x || p.y >= d.y) return vec4(0);
    return fetch_texel(s, p.x, p.y, lod);
}
inline int wrap_repeat(int i, int n) { const int r = i % n; return r < 0 ? r + n : r; }
inline vec4 texture(const sampler2D& s, const vec2& uv) {
    const float x = uv.x * float(s.w) - 0.5f, y = uv.y * float(s.h) - 0.5f;
    const float fx = std::floor(x), fy = std::floor(y);
    const float a = x - fx, b = y - fy;
    const int x0 = wrap_repeat(int(fx), s.w), y0 = wrap_repeat(int(fy), s.h);
    const int x1 = wrap_repeat(x0 + 1, s.w), y1 = wrap_repeat(y0 + 1, s.h);
    return glm::mix(glm::mix(fetch_texel(s, x0, y0, 0), fetch_texel(s, x1, y0, 0), a),
                    glm::mix(fetch_texel(s, x0, y1, 0), fetch_texel(s, x1, y1, 0), a), b);
}

/* brick textures: words of voldata's Buf3D<uint32_t> / Buf3D<uint8_t>, x fastest (voldata/buf3d.h:27-29) */
struct usampler3D { const uint32_t* data = nullptr; ivec3 dim = ivec3(0); };              /* GL_RGB10_A2UI */
struct sampler3D {                                                                         /* GL_RG16F (+3 mips) or GL_R8 */
    const uint32_t* rg16f[4] = {};
    const uint8_t* r8 = nullptr;
    ivec3 dim = ivec3(0);
    int levels = 0;
};
inline uvec4 texelFetch(const usampler3D& s, const ivec3& p, int lod) {
    if (!s.data || lod != 0 || p.x < 0 || p.y < 0 || p.z < 0 || p.x >= s.dim.x || p.y >= s.dim.y || p.z >= s.dim.z) return uvec4(0);
    const uint32_t w = s.data[(size_t(p.z) * s.dim.y + p.y) * s.dim.x + p.x];
    return uvec4(w >> 22, (w >> 12) & 1023u, (w >> 2) & 1023u, w & 3u);
}
inline vec4 texelFetch(const sampler3D& s, const ivec3& p, int lod) {
    if (lod < 0 || lod >= s.levels) return vec4(0);
    const ivec3 d = s.dim >> lod;
    if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= d.x || p.y >= d.y || p.z >= d.z) return vec4(0);
    const size_t i = (size_t(p.z) * d.y + p.y) * d.x + p.x;
    if (s.r8) return vec4(float(s.r8[i]) / 255.f, 0.f, 0.f, 1.f);
    const uint32_t w = s.rg16f[lod][i];
    return vec4(glm::detail::toFloat32(glm::detail::hdata(w & 0xffffu)), glm::detail::toFloat32(glm::detail::hdata(w >> 16)), 0.f, 1.f);
}

struct invocation_id { uvec2 xy; };           /* only `.xy` is read (ivec2(gl_GlobalInvocationID.xy)) */
extern thread_local invocation_id gl_GlobalInvocationID;

/* R10 of glsl2cpp.py: `int(x)` of the shader text. float -> int as the GPU converts (saturating, NaN -> 0); integers pass through. */
inline int glsl_int(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.f) return 2147483647;
    if (x <= -2147483648.f) return -2147483647 - 1;
    return static_cast<int>(x);
}
inline int glsl_int(int x) { return x; }
inline int glsl_int(uint x) { return static_cast<int>(x); }
}  // namespace glsl
