// Minimal stand-in for the subset of glm 0.9.9.8 the reference's scene / transform code uses
// (README.md:8 pins glm 0.9.9.8; the library is not vendored in the reference and not installed here).
// Only what Scenes.cpp, TransformComponent.cpp, RTModel.cpp and RaytraceScene.cpp need: vec2/3/4, mat4
// (column-major, m[col][row]), radians, pi, sin, cos, max.  All binary32, no implicit contraction.
#pragma once

#include <cmath>
#include <cstddef>

namespace glm {

struct vec2 {
    float x{}, y{};
    constexpr vec2() = default;
    constexpr vec2(float s) : x(s), y(s) {}
    constexpr vec2(float x_, float y_) : x(x_), y(y_) {}
    constexpr bool operator==(const vec2& o) const { return x == o.x && y == o.y; }
};

struct vec3 {
    float x{}, y{}, z{};
    constexpr vec3() = default;
    constexpr vec3(float s) : x(s), y(s), z(s) {}
    constexpr vec3(int s) : x(float(s)), y(float(s)), z(float(s)) {}
    template <class A, class B, class C> constexpr vec3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
    constexpr bool operator==(const vec3& o) const { return x == o.x && y == o.y && z == o.z; }
    constexpr float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float length() const { return 3; }   // glm::vec3::length() is the component count (the reference calls it, Scenes.cpp:33)
};
constexpr vec3 operator+(vec3 a, vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
constexpr vec3 operator-(vec3 a, vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
constexpr vec3 operator*(vec3 a, vec3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
constexpr vec3 operator*(float s, vec3 a) { return { s * a.x, s * a.y, s * a.z }; }
constexpr vec3 operator*(vec3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
constexpr vec3 operator/(float s, vec3 a) { return { s / a.x, s / a.y, s / a.z }; }
constexpr vec3 operator/(vec3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
constexpr vec3 operator-(vec3 a) { return { -a.x, -a.y, -a.z }; }

struct vec4 {
    float x{}, y{}, z{}, w{};
    constexpr vec4() = default;
    constexpr vec4(float s) : x(s), y(s), z(s), w(s) {}
    constexpr vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    constexpr vec4(vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    constexpr bool operator==(const vec4& o) const { return x == o.x && y == o.y && z == o.z && w == o.w; }
    constexpr float& operator[](int i) { return (&x)[i]; }
    constexpr const float& operator[](int i) const { return (&x)[i]; }
};

struct mat3 {
    vec3 c[3];
    constexpr mat3() = default;
    constexpr mat3(vec3 a, vec3 b, vec3 d) : c{ a, b, d } {}
    constexpr vec3& operator[](int i) { return c[i]; }
    constexpr const vec3& operator[](int i) const { return c[i]; }
};

struct mat4 {
    vec4 c[4];   // columns
    constexpr mat4() = default;
    constexpr explicit mat4(float d) : c{ vec4(d, 0, 0, 0), vec4(0, d, 0, 0), vec4(0, 0, d, 0), vec4(0, 0, 0, d) } {}
    constexpr mat4(vec4 a, vec4 b, vec4 d, vec4 e) : c{ a, b, d, e } {}
    constexpr vec4& operator[](int i) { return c[i]; }
    constexpr const vec4& operator[](int i) const { return c[i]; }
    constexpr bool operator==(const mat4& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2] && c[3] == o.c[3]; }
};
// M * v = ((c0*x + c1*y) + c2*z) + c3*w  -- the operation order pinned for ModelSpaceToWorldSpace.comp:37-45
constexpr vec4 operator*(const mat4& m, vec4 v) {
    vec4 r;
    for (int i = 0; i < 4; i++) r[i] = ((m.c[0][i] * v.x + m.c[1][i] * v.y) + m.c[2][i] * v.z) + m.c[3][i] * v.w;
    return r;
}

template <class T> constexpr T pi() { return T(3.14159265358979323846264338327950288); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float cos(float a) { return std::cos(a); }
inline float sin(float a) { return std::sin(a); }
template <class T> constexpr T max(T a, T b) { return a < b ? b : a; }
template <class T> constexpr T min(T a, T b) { return b < a ? b : a; }

}  // namespace glm
