// kfglm.hpp -- the small subset of the glm API that the kuafu.hpp surface exposes
// (glm::vec2/3/4, ivec3, mat4, translate/rotate/scale/lookAt/perspective/inverse/transpose ...).
//
// glm is not available in this build environment, so the facade ships this drop-in subset with
// glm's conventions: column-major mat4 indexed m[col][row], right-handed lookAt, [-1,1] clip depth
// perspective (the reference defines no GLM_FORCE_* macro).  Define KUAFU_USE_SYSTEM_GLM to compile
// the facade against the real glm instead (what a SAPIEN build would do).
#pragma once

#ifdef KUAFU_USE_SYSTEM_GLM
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <glm/gtc/constants.hpp>
#else

#include <cmath>
#include <cstddef>

namespace glm {

struct vec2 {
  float x = 0, y = 0;
  vec2() = default;
  explicit vec2(float s) : x(s), y(s) {}
  template <typename A, typename B>
  vec2(A a, B b) : x(float(a)), y(float(b)) {}
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};

struct vec3 {
  float x = 0, y = 0, z = 0;
  vec3() = default;
  explicit vec3(float s) : x(s), y(s), z(s) {}
  template <typename A, typename B, typename C>
  vec3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
  vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct ivec3 {
  int x = 0, y = 0, z = 0;
  ivec3() = default;
  ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
  int& operator[](int i) { return (&x)[i]; }
  const int& operator[](int i) const { return (&x)[i]; }
};

struct vec4 {
  float x = 0, y = 0, z = 0, w = 0;
  vec4() = default;
  explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
  template <typename A, typename B, typename C, typename D>
  vec4(A a, B b, C c, D d) : x(float(a)), y(float(b)), z(float(c)), w(float(d)) {}
  template <typename D>
  vec4(const vec3& v, D d) : x(v.x), y(v.y), z(v.z), w(float(d)) {}
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};

inline vec3 operator+(const vec3& a, const vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(const vec3& a, const vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(const vec3& a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(const vec3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, const vec3& a) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(const vec3& a, const vec3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(const vec3& a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const vec3& a, const vec3& b) { return !(a == b); }
inline bool operator==(const vec2& a, const vec2& b) { return a.x == b.x && a.y == b.y; }
inline vec2 operator+(const vec2& a, const vec2& b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator*(const vec2& a, float s) { return {a.x * s, a.y * s}; }
inline vec4 operator+(const vec4& a, const vec4& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator*(const vec4& a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline bool operator==(const vec4& a, const vec4& b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) {
  return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / std::sqrt(dot(a, a))); }

template <typename T = float>
constexpr T pi() { return T(3.14159265358979323846264338327950288); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float cos(float a) { return std::cos(a); }
inline float sin(float a) { return std::sin(a); }

struct mat4 {
  vec4 c[4];
  mat4() : mat4(1.0f) {}
  explicit mat4(float d) {
    c[0] = vec4(d, 0, 0, 0);
    c[1] = vec4(0, d, 0, 0);
    c[2] = vec4(0, 0, d, 0);
    c[3] = vec4(0, 0, 0, d);
  }
  mat4(const vec4& a, const vec4& b, const vec4& cc, const vec4& d) { c[0] = a; c[1] = b; c[2] = cc; c[3] = d; }
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};

inline vec4 operator*(const mat4& m, const vec4& v) {
  return m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w;
}
inline mat4 operator*(const mat4& a, const mat4& b) {
  mat4 r(0.0f);
  for (int j = 0; j < 4; j++) r[j] = a[0] * b[j].x + a[1] * b[j].y + a[2] * b[j].z + a[3] * b[j].w;
  return r;
}
inline bool operator==(const mat4& a, const mat4& b) {
  return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3];
}
inline mat4 transpose(const mat4& m) {
  mat4 r(0.0f);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r[i][j] = m[j][i];
  return r;
}
inline mat4 translate(const mat4& m, const vec3& v) {
  mat4 r = m;
  r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
  return r;
}
inline mat4 scale(const mat4& m, const vec3& v) {
  mat4 r = m;
  r[0] = m[0] * v.x;
  r[1] = m[1] * v.y;
  r[2] = m[2] * v.z;
  return r;
}
inline mat4 rotate(const mat4& m, float angle, const vec3& v) {
  const float c = std::cos(angle), s = std::sin(angle);
  const vec3 axis = normalize(v);
  const vec3 t = axis * (1.0f - c);
  float R[3][3];
  R[0][0] = c + t.x * axis.x;
  R[0][1] = t.x * axis.y + s * axis.z;
  R[0][2] = t.x * axis.z - s * axis.y;
  R[1][0] = t.y * axis.x - s * axis.z;
  R[1][1] = c + t.y * axis.y;
  R[1][2] = t.y * axis.z + s * axis.x;
  R[2][0] = t.z * axis.x + s * axis.y;
  R[2][1] = t.z * axis.y - s * axis.x;
  R[2][2] = c + t.z * axis.z;
  mat4 r(0.0f);
  r[0] = m[0] * R[0][0] + m[1] * R[0][1] + m[2] * R[0][2];
  r[1] = m[0] * R[1][0] + m[1] * R[1][1] + m[2] * R[1][2];
  r[2] = m[0] * R[2][0] + m[1] * R[2][1] + m[2] * R[2][2];
  r[3] = m[3];
  return r;
}
inline mat4 lookAt(const vec3& eye, const vec3& center, const vec3& up) {
  const vec3 f = normalize(center - eye);
  const vec3 s = normalize(cross(f, up));
  const vec3 u = cross(s, f);
  mat4 r(1.0f);
  r[0][0] = s.x; r[1][0] = s.y; r[2][0] = s.z;
  r[0][1] = u.x; r[1][1] = u.y; r[2][1] = u.z;
  r[0][2] = -f.x; r[1][2] = -f.y; r[2][2] = -f.z;
  r[3][0] = -dot(s, eye);
  r[3][1] = -dot(u, eye);
  r[3][2] = dot(f, eye);
  return r;
}
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
  const float t = std::tan(fovy / 2.0f);
  mat4 r(0.0f);
  r[0][0] = 1.0f / (aspect * t);
  r[1][1] = 1.0f / t;
  r[2][2] = -(zFar + zNear) / (zFar - zNear);
  r[2][3] = -1.0f;
  r[3][2] = -(2.0f * zFar * zNear) / (zFar - zNear);
  return r;
}
// General 4x4 inverse by cofactors (2x2 sub-determinants of the lower rows first).
inline mat4 inverse(const mat4& m) {
  const float a00 = m[0][0], a01 = m[0][1], a02 = m[0][2], a03 = m[0][3];
  const float a10 = m[1][0], a11 = m[1][1], a12 = m[1][2], a13 = m[1][3];
  const float a20 = m[2][0], a21 = m[2][1], a22 = m[2][2], a23 = m[2][3];
  const float a30 = m[3][0], a31 = m[3][1], a32 = m[3][2], a33 = m[3][3];
  const float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10;
  const float b03 = a01 * a12 - a02 * a11, b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
  const float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30, b08 = a20 * a33 - a23 * a30;
  const float b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
  const float det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
  const float id = 1.0f / det;
  mat4 r(0.0f);
  r[0][0] = (a11 * b11 - a12 * b10 + a13 * b09) * id;
  r[0][1] = (a02 * b10 - a01 * b11 - a03 * b09) * id;
  r[0][2] = (a31 * b05 - a32 * b04 + a33 * b03) * id;
  r[0][3] = (a22 * b04 - a21 * b05 - a23 * b03) * id;
  r[1][0] = (a12 * b08 - a10 * b11 - a13 * b07) * id;
  r[1][1] = (a00 * b11 - a02 * b08 + a03 * b07) * id;
  r[1][2] = (a32 * b02 - a30 * b05 - a33 * b01) * id;
  r[1][3] = (a20 * b05 - a22 * b02 + a23 * b01) * id;
  r[2][0] = (a10 * b10 - a11 * b08 + a13 * b06) * id;
  r[2][1] = (a01 * b08 - a00 * b10 - a03 * b06) * id;
  r[2][2] = (a30 * b04 - a31 * b02 + a33 * b00) * id;
  r[2][3] = (a21 * b02 - a20 * b04 - a23 * b00) * id;
  r[3][0] = (a11 * b07 - a10 * b09 - a12 * b06) * id;
  r[3][1] = (a00 * b09 - a01 * b07 + a02 * b06) * id;
  r[3][2] = (a31 * b01 - a30 * b03 - a32 * b00) * id;
  r[3][3] = (a20 * b03 - a21 * b01 + a22 * b00) * id;
  return r;
}

}  // namespace glm
#endif  // KUAFU_USE_SYSTEM_GLM
