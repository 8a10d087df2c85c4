// vmath.h -- the small slice of glm the VolRen host API exposes (vec2/3/4, ivec*, uvec*, mat3, mat4, quat),
// written from scratch so that the host needs no third-party headers. Conventions are glm's defaults, which
// the reference relies on: column-major matrices (m[c][r]), right-handed lookAt, radians, quat
// memory order x,y,z,w. `namespace glm` aliases this namespace so reference-style call sites
// (`renderer->albedo = glm::vec3(0.8f)`) keep compiling.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>

namespace vmath {

template <typename T> struct tvec2 {
    T x, y;
    constexpr tvec2() : x(0), y(0) {}
    constexpr explicit tvec2(T s) : x(s), y(s) {}
    constexpr tvec2(T x, T y) : x(x), y(y) {}
    template <typename U> constexpr explicit tvec2(const tvec2<U>& o) : x(T(o.x)), y(T(o.y)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct tvec4;
template <typename T> struct tvec3 {
    T x, y, z;
    constexpr tvec3() : x(0), y(0), z(0) {}
    constexpr explicit tvec3(T s) : x(s), y(s), z(s) {}
    constexpr tvec3(T x, T y, T z) : x(x), y(y), z(z) {}
    template <typename U> constexpr explicit tvec3(const tvec3<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
    constexpr explicit tvec3(const tvec4<T>& o);
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct tvec4 {
    T x, y, z, w;
    constexpr tvec4() : x(0), y(0), z(0), w(0) {}
    constexpr explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
    constexpr tvec4(T x, T y, T z, T w) : x(x), y(y), z(z), w(w) {}
    constexpr tvec4(const tvec3<T>& v, T w) : x(v.x), y(v.y), z(v.z), w(w) {}
    template <typename U> constexpr explicit tvec4(const tvec4<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(o.w)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    // glm names the colour swizzle too (transferfunc.cpp uses .a)
    T& a() { return w; }
};
template <typename T> constexpr tvec3<T>::tvec3(const tvec4<T>& o) : x(o.x), y(o.y), z(o.z) {}

using vec2 = tvec2<float>;   using vec3 = tvec3<float>;   using vec4 = tvec4<float>;
using ivec2 = tvec2<int32_t>; using ivec3 = tvec3<int32_t>; using ivec4 = tvec4<int32_t>;
using uvec2 = tvec2<uint32_t>; using uvec3 = tvec3<uint32_t>; using uvec4 = tvec4<uint32_t>;

// component-wise arithmetic (vector op vector, vector op scalar, scalar op vector, compound, unary minus)
#define VMATH_OPS(V, ...)                                                                                         \
    template <typename T> constexpr V<T> operator+(const V<T>& a, const V<T>& b) { return V<T>(__VA_ARGS__(+)); } \
    template <typename T> constexpr V<T> operator-(const V<T>& a, const V<T>& b) { return V<T>(__VA_ARGS__(-)); } \
    template <typename T> constexpr V<T> operator*(const V<T>& a, const V<T>& b) { return V<T>(__VA_ARGS__(*)); } \
    template <typename T> constexpr V<T> operator/(const V<T>& a, const V<T>& b) { return V<T>(__VA_ARGS__(/)); } \
    template <typename T> constexpr V<T> operator+(const V<T>& a, T s) { return a + V<T>(s); }                    \
    template <typename T> constexpr V<T> operator-(const V<T>& a, T s) { return a - V<T>(s); }                    \
    template <typename T> constexpr V<T> operator*(const V<T>& a, T s) { return a * V<T>(s); }                    \
    template <typename T> constexpr V<T> operator/(const V<T>& a, T s) { return a / V<T>(s); }                    \
    template <typename T> constexpr V<T> operator+(T s, const V<T>& a) { return V<T>(s) + a; }                    \
    template <typename T> constexpr V<T> operator-(T s, const V<T>& a) { return V<T>(s) - a; }                    \
    template <typename T> constexpr V<T> operator*(T s, const V<T>& a) { return V<T>(s) * a; }                    \
    template <typename T> constexpr V<T> operator/(T s, const V<T>& a) { return V<T>(s) / a; }                    \
    template <typename T> V<T>& operator+=(V<T>& a, const V<T>& b) { return a = a + b; }                          \
    template <typename T> V<T>& operator-=(V<T>& a, const V<T>& b) { return a = a - b; }                          \
    template <typename T> V<T>& operator*=(V<T>& a, const V<T>& b) { return a = a * b; }                          \
    template <typename T> V<T>& operator/=(V<T>& a, const V<T>& b) { return a = a / b; }                          \
    template <typename T> V<T>& operator+=(V<T>& a, T s) { return a = a + s; }                                    \
    template <typename T> V<T>& operator-=(V<T>& a, T s) { return a = a - s; }                                    \
    template <typename T> V<T>& operator*=(V<T>& a, T s) { return a = a * s; }                                    \
    template <typename T> V<T>& operator/=(V<T>& a, T s) { return a = a / s; }
#define VMATH_E2(op) a.x op b.x, a.y op b.y
#define VMATH_E3(op) a.x op b.x, a.y op b.y, a.z op b.z
#define VMATH_E4(op) a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w
VMATH_OPS(tvec2, VMATH_E2)
VMATH_OPS(tvec3, VMATH_E3)
VMATH_OPS(tvec4, VMATH_E4)
#undef VMATH_OPS
// unary minus (for unsigned: modular negation, as in glm)
template <typename T> constexpr tvec2<T> operator-(const tvec2<T>& a) { return tvec2<T>(T(0) - a.x, T(0) - a.y); }
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a) { return tvec3<T>(T(0) - a.x, T(0) - a.y, T(0) - a.z); }
template <typename T> constexpr tvec4<T> operator-(const tvec4<T>& a) { return tvec4<T>(T(0) - a.x, T(0) - a.y, T(0) - a.z, T(0) - a.w); }
template <typename T> constexpr bool operator==(const tvec3<T>& a, const tvec3<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
template <typename V> inline float length(const V& v) { return std::sqrt(dot(v, v)); }
template <typename V> inline V normalize(const V& v) { return v * (1.f / std::sqrt(dot(v, v))); }   // glm: v * inversesqrt(dot(v, v))
inline vec3 min(const vec3& a, const vec3& b) { return vec3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }
inline constexpr float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

// ---- matrices (column-major: m[c] is column c) ----
struct mat3 {
    vec3 c[3];
    mat3() : mat3(1.f) {}
    explicit mat3(float d) : c{ vec3(d, 0, 0), vec3(0, d, 0), vec3(0, 0, d) } {}
    mat3(const vec3& c0, const vec3& c1, const vec3& c2) : c{ c0, c1, c2 } {}
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() : mat4(1.f) {}
    explicit mat4(float d) : c{ vec4(d, 0, 0, 0), vec4(0, d, 0, 0), vec4(0, 0, d, 0), vec4(0, 0, 0, d) } {}
    mat4(const vec4& c0, const vec4& c1, const vec4& c2, const vec4& c3) : c{ c0, c1, c2, c3 } {}
    mat4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3, float c0, float c1, float c2, float c3, float d0, float d1,
         float d2, float d3)
        : c{ vec4(a0, a1, a2, a3), vec4(b0, b1, b2, b3), vec4(c0, c1, c2, c3), vec4(d0, d1, d2, d3) } {}
    explicit mat4(const mat3& m) : c{ vec4(m[0], 0), vec4(m[1], 0), vec4(m[2], 0), vec4(0, 0, 0, 1) } {}
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
    const float* data() const { return &c[0].x; }
};
inline mat3 to_mat3(const mat4& m) { return mat3(vec3(m[0]), vec3(m[1]), vec3(m[2])); }   // glm::mat3(mat4)

inline vec3 operator*(const mat3& m, const vec3& v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z; }
inline vec4 operator*(const mat4& m, const vec4& v) { return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w; }
inline mat3 operator*(const mat3& a, const mat3& b) { return mat3(a * b[0], a * b[1], a * b[2]); }
inline mat4 operator*(const mat4& a, const mat4& b) { return mat4(a * b[0], a * b[1], a * b[2], a * b[3]); }
inline mat3 operator*(const mat3& a, float s) { return mat3(a[0] * s, a[1] * s, a[2] * s); }
inline mat3 operator*(float s, const mat3& a) { return a * s; }
inline mat4 operator*(const mat4& a, float s) { return mat4(a[0] * s, a[1] * s, a[2] * s, a[3] * s); }
inline mat4 operator*(float s, const mat4& a) { return a * s; }
inline mat3 operator+(const mat3& a, const mat3& b) { return mat3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline mat3 operator-(const mat3& a, const mat3& b) { return mat3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline mat4 operator+(const mat4& a, const mat4& b) { return mat4(a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]); }
inline mat4 operator-(const mat4& a, const mat4& b) { return mat4(a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]); }
inline mat3 operator-(const mat3& a) { return mat3(-a[0], -a[1], -a[2]); }
inline mat4 operator-(const mat4& a) { return mat4(-a[0], -a[1], -a[2], -a[3]); }
inline mat3& operator+=(mat3& a, const mat3& b) { return a = a + b; }
inline mat3& operator-=(mat3& a, const mat3& b) { return a = a - b; }
inline mat3& operator*=(mat3& a, const mat3& b) { return a = a * b; }
inline mat3& operator*=(mat3& a, float s) { return a = a * s; }
inline mat4& operator+=(mat4& a, const mat4& b) { return a = a + b; }
inline mat4& operator-=(mat4& a, const mat4& b) { return a = a - b; }
inline mat4& operator*=(mat4& a, const mat4& b) { return a = a * b; }
inline mat4& operator*=(mat4& a, float s) { return a = a * s; }

inline mat3 transpose(const mat3& m) { return mat3(vec3(m[0].x, m[1].x, m[2].x), vec3(m[0].y, m[1].y, m[2].y), vec3(m[0].z, m[1].z, m[2].z)); }
inline mat4 transpose(const mat4& m) {
    return mat4(vec4(m[0].x, m[1].x, m[2].x, m[3].x), vec4(m[0].y, m[1].y, m[2].y, m[3].y), vec4(m[0].z, m[1].z, m[2].z, m[3].z),
                vec4(m[0].w, m[1].w, m[2].w, m[3].w));
}
// adjugate / determinant (fp32, like glm::inverse)
inline mat3 inverse(const mat3& m) {
    const float det = m[0].x * (m[1].y * m[2].z - m[2].y * m[1].z) - m[1].x * (m[0].y * m[2].z - m[2].y * m[0].z) +
                      m[2].x * (m[0].y * m[1].z - m[1].y * m[0].z);
    const float id = 1.f / det;
    mat3 r;
    r[0].x = +(m[1].y * m[2].z - m[2].y * m[1].z) * id;
    r[1].x = -(m[1].x * m[2].z - m[2].x * m[1].z) * id;
    r[2].x = +(m[1].x * m[2].y - m[2].x * m[1].y) * id;
    r[0].y = -(m[0].y * m[2].z - m[2].y * m[0].z) * id;
    r[1].y = +(m[0].x * m[2].z - m[2].x * m[0].z) * id;
    r[2].y = -(m[0].x * m[2].y - m[2].x * m[0].y) * id;
    r[0].z = +(m[0].y * m[1].z - m[1].y * m[0].z) * id;
    r[1].z = -(m[0].x * m[1].z - m[1].x * m[0].z) * id;
    r[2].z = +(m[0].x * m[1].y - m[1].x * m[0].y) * id;
    return r;
}
inline mat4 inverse(const mat4& m) {
    const float* a = m.data();   // a[c * 4 + r]
    float inv[16];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    const float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    const float id = 1.f / det;
    mat4 r;
    for (int i = 0; i < 16; ++i) (&r[0].x)[i] = inv[i] * id;
    return r;
}

inline mat4 translate(const mat4& m, const vec3& v) {
    mat4 r = m;
    r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
    return r;
}
inline mat4 scale(const mat4& m, const vec3& v) { return mat4(m[0] * v.x, m[1] * v.y, m[2] * v.z, m[3]); }
inline mat4 rotate(const mat4& m, float angle, const vec3& v) {
    const float c = std::cos(angle), s = std::sin(angle);
    const vec3 axis = normalize(v), temp = axis * (1.f - c);
    mat3 R;
    R[0] = vec3(c + temp.x * axis.x, temp.x * axis.y + s * axis.z, temp.x * axis.z - s * axis.y);
    R[1] = vec3(temp.y * axis.x - s * axis.z, c + temp.y * axis.y, temp.y * axis.z + s * axis.x);
    R[2] = vec3(temp.z * axis.x + s * axis.y, temp.z * axis.y - s * axis.x, c + temp.z * axis.z);
    return mat4(m[0] * R[0].x + m[1] * R[0].y + m[2] * R[0].z, m[0] * R[1].x + m[1] * R[1].y + m[2] * R[1].z,
                m[0] * R[2].x + m[1] * R[2].y + m[2] * R[2].z, m[3]);
}
// right-handed, like glm::lookAt without GLM_FORCE_LEFT_HANDED
inline mat4 lookAt(const vec3& eye, const vec3& center, const vec3& up) {
    const vec3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
    mat4 r(1.f);
    r[0].x = s.x; r[1].x = s.y; r[2].x = s.z;
    r[0].y = u.x; r[1].y = u.y; r[2].y = u.z;
    r[0].z = -f.x; r[1].z = -f.y; r[2].z = -f.z;
    r[3].x = -dot(s, eye); r[3].y = -dot(u, eye); r[3].z = dot(f, eye);
    return r;
}
// right-handed, depth -1..1 (glm::perspectiveRH_NO)
inline mat4 perspective(float fovy, float aspect, float z_near, float z_far) {
    const float t = std::tan(fovy / 2.f);
    mat4 r(0.f);
    r[0].x = 1.f / (aspect * t);
    r[1].y = 1.f / t;
    r[2].z = -(z_far + z_near) / (z_far - z_near);
    r[2].w = -1.f;
    r[3].z = -(2.f * z_far * z_near) / (z_far - z_near);
    return r;
}

// ---- quaternion (memory order x, y, z, w as in glm 0.9.9 without GLM_FORCE_QUAT_DATA_WXYZ) ----
struct quat {
    float x, y, z, w;
    quat() : x(0), y(0), z(0), w(1) {}
    quat(float w, float x, float y, float z) : x(x), y(y), z(z), w(w) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
// glm::quat_cast
inline quat toQuat(const mat3& m) {
    const float fx = m[0].x - m[1].y - m[2].z, fy = m[1].y - m[0].x - m[2].z, fz = m[2].z - m[0].x - m[1].y, fw = m[0].x + m[1].y + m[2].z;
    int big = 0;
    float four = fw;
    if (fx > four) { four = fx; big = 1; }
    if (fy > four) { four = fy; big = 2; }
    if (fz > four) { four = fz; big = 3; }
    const float v = std::sqrt(four + 1.f) * 0.5f, mult = 0.25f / v;
    switch (big) {
        case 0: return quat(v, (m[1].z - m[2].y) * mult, (m[2].x - m[0].z) * mult, (m[0].y - m[1].x) * mult);
        case 1: return quat((m[1].z - m[2].y) * mult, v, (m[0].y + m[1].x) * mult, (m[2].x + m[0].z) * mult);
        case 2: return quat((m[2].x - m[0].z) * mult, (m[0].y + m[1].x) * mult, v, (m[1].z + m[2].y) * mult);
        default: return quat((m[0].y - m[1].x) * mult, (m[2].x + m[0].z) * mult, (m[1].z + m[2].y) * mult, v);
    }
}
inline quat toQuat(const mat4& m) { return toQuat(to_mat3(m)); }
// glm::quat(vec3 eulerAngles): pitch (x), yaw (y), roll (z)
inline quat quat_from_euler(const vec3& e) {
    const vec3 c(std::cos(e.x * .5f), std::cos(e.y * .5f), std::cos(e.z * .5f)), s(std::sin(e.x * .5f), std::sin(e.y * .5f), std::sin(e.z * .5f));
    return quat(c.x * c.y * c.z + s.x * s.y * s.z, s.x * c.y * c.z - c.x * s.y * s.z, c.x * s.y * c.z + s.x * c.y * s.z, c.x * c.y * s.z - s.x * s.y * c.z);
}
inline quat normalize(const quat& q) {
    const float len = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    if (len <= 0.f) return quat(1, 0, 0, 0);
    const float il = 1.f / len;
    return quat(q.w * il, q.x * il, q.y * il, q.z * il);
}
inline quat operator+(const quat& a, const quat& b) { return quat(a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z); }
inline quat operator-(const quat& a, const quat& b) { return quat(a.w - b.w, a.x - b.x, a.y - b.y, a.z - b.z); }
inline quat operator-(const quat& a) { return quat(-a.w, -a.x, -a.y, -a.z); }
inline quat operator*(const quat& p, const quat& q) {
    return quat(p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z, p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
                p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z, p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x);
}
inline quat operator*(const quat& a, float s) { return quat(a.w * s, a.x * s, a.y * s, a.z * s); }
inline quat operator*(float s, const quat& a) { return a * s; }
inline quat& operator+=(quat& a, const quat& b) { return a = a + b; }
inline quat& operator-=(quat& a, const quat& b) { return a = a - b; }
inline quat& operator*=(quat& a, const quat& b) { return a = a * b; }
inline quat& operator*=(quat& a, float s) { return a = a * s; }

// ---- glm::to_string look-alikes ("vec3(1.000000, 2.000000, 3.000000)") ----
namespace detail {
inline std::string fmt(float v) { char b[64]; snprintf(b, sizeof b, "%f", v); return b; }
inline std::string fmt(int32_t v) { return std::to_string(v); }
inline std::string fmt(uint32_t v) { return std::to_string(v); }
template <typename T> constexpr const char* prefix() { return ""; }
template <> constexpr const char* prefix<int32_t>() { return "i"; }
template <> constexpr const char* prefix<uint32_t>() { return "u"; }
}  // namespace detail
template <typename T> std::string to_string(const tvec2<T>& v) { return std::string(detail::prefix<T>()) + "vec2(" + detail::fmt(v.x) + ", " + detail::fmt(v.y) + ")"; }
template <typename T> std::string to_string(const tvec3<T>& v) {
    return std::string(detail::prefix<T>()) + "vec3(" + detail::fmt(v.x) + ", " + detail::fmt(v.y) + ", " + detail::fmt(v.z) + ")";
}
template <typename T> std::string to_string(const tvec4<T>& v) {
    return std::string(detail::prefix<T>()) + "vec4(" + detail::fmt(v.x) + ", " + detail::fmt(v.y) + ", " + detail::fmt(v.z) + ", " + detail::fmt(v.w) + ")";
}
inline std::string to_string(const mat3& m) {
    std::string s = "mat3x3(";
    for (int c = 0; c < 3; ++c) s += std::string(c ? ", " : "") + "(" + detail::fmt(m[c].x) + ", " + detail::fmt(m[c].y) + ", " + detail::fmt(m[c].z) + ")";
    return s + ")";
}
inline std::string to_string(const mat4& m) {
    std::string s = "mat4x4(";
    for (int c = 0; c < 4; ++c)
        s += std::string(c ? ", " : "") + "(" + detail::fmt(m[c].x) + ", " + detail::fmt(m[c].y) + ", " + detail::fmt(m[c].z) + ", " + detail::fmt(m[c].w) + ")";
    return s + ")";
}
inline std::string to_string(const quat& q) { return "quat(" + detail::fmt(q.w) + ", {" + detail::fmt(q.x) + ", " + detail::fmt(q.y) + ", " + detail::fmt(q.z) + "})"; }

}  // namespace vmath

#ifndef VOLREN_NO_GLM_ALIAS
namespace glm = vmath;
#endif
