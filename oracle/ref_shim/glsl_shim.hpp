// glsl_shim.hpp -- TEST INFRASTRUCTURE (part of the oracle, never linked into the product).
//
// The GLSL the reference's ray-tracing shaders are written in, as a C++ vocabulary: vector / matrix
// types, built-in functions, ray-tracing built-in variables and the resource kinds a shader stage
// can declare.  Together with glsl2cpp.py this lets g++ compile resources/shaders/PathTrace.{rgen,
// rchit,rahit,rmiss} and PathTraceShadow.rmiss from /root/reference unmodified and run them on the
// CPU (oracle/_ref/libkf_ref.so), which is what pins the hand-written restatement in kf_oracle.cpp.
//
// What GLSL leaves to the implementation is fixed here the same way the oracle's arithmetic contract
// fixes it (kf_oracle.cpp header): IEEE binary32, no FMA contraction,
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z
//   length(v)     = sqrt(dot(v,v));  normalize(v) = v * (1 / sqrt(dot(v,v)))
//   mat * vec     = columns summed left to right;  vec * mat = one dot() per column
//   mix(x,y,a)    = x*(1-a) + y*a;  clamp / min / max = fmin / fmax
//   pow(x,y)      = exp2(y * log2(max(x, 0)))   (Vulkan: "inherited from exp2(y*log2(x))", undefined
//                   for x < 0 -- the clamp gives 0 there instead of NaN)
//   reflect / refract as in the GLSL specification (refract returns 0 on total internal reflection)
// The two hardware black boxes are callbacks of the pipeline: traceRayEXT (acceleration structure
// traversal + triangle test) and texture() (sampler).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#undef M_PI  // base/Random.glsl defines its own (3.141592)

namespace glsl {

typedef uint32_t uint;

// GLSL bool inside a uniform / push-constant block occupies 4 bytes.
struct bool32 {
  uint32_t v;
  operator bool() const { return v != 0; }
};

struct uvec2 {
  uint x, y;
};
struct uvec3 {
  uint x, y, z;
  uvec2 xy() const { return {x, y}; }
};
struct ivec2 {
  int x, y;
  ivec2() : x(0), y(0) {}
  template <class A, class B> ivec2(A a, B b) : x(int(a)), y(int(b)) {}
  explicit ivec2(uvec2 u) : x(int(u.x)), y(int(u.y)) {}
};
struct ivec3 {
  int x, y, z;
  ivec3() : x(0), y(0), z(0) {}
  template <class A, class B, class C> ivec3(A a, B b, C c) : x(int(a)), y(int(b)), z(int(c)) {}
};

struct vec2 {
  float x, y;
  vec2() : x(0), y(0) {}
  template <class A, class = decltype(float(A()))> explicit vec2(A s) : x(float(s)), y(float(s)) {}
  template <class A, class B> vec2(A a, B b) : x(float(a)), y(float(b)) {}
  explicit vec2(uvec2 u) : x(float(u.x)), y(float(u.y)) {}
};
struct vec4;
struct vec3 {
  union { float x, r; };
  union { float y, g; };
  union { float z, b; };
  vec3() : x(0), y(0), z(0) {}
  template <class A, class = decltype(float(A()))> explicit vec3(A s) : x(float(s)), y(float(s)), z(float(s)) {}
  template <class A, class B, class C> vec3(A a, B b_, C c) : x(float(a)), y(float(b_)), z(float(c)) {}
  vec3(vec2 v, float c) : x(v.x), y(v.y), z(c) {}
  explicit vec3(const vec4& v);
  vec2 xy() const { return vec2(x, y); }
  vec3 xyz() const { return *this; }
  vec3 rgb() const { return *this; }
};
struct vec4 {
  union { float x, r; };
  union { float y, g; };
  union { float z, b; };
  union { float w, a; };
  vec4() : x(0), y(0), z(0), w(0) {}
  vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
  vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  vec4(vec2 v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
  vec2 xy() const { return vec2(x, y); }
  vec3 xyz() const { return vec3(x, y, z); }
  vec3 rgb() const { return vec3(x, y, z); }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

// ---- arithmetic -------------------------------------------------------------------------------
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(vec3 a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline bool operator==(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }  // all()
inline bool operator!=(vec3 a, vec3 b) { return a.x != b.x || a.y != b.y || a.z != b.z; }  // any()

inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator/(vec4 a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline vec4& operator/=(vec4& a, float s) { a = a / s; return a; }

// ---- matrices (column major, like GLSL) ---------------------------------------------------------
struct mat4 {
  vec4 c[4];
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, vec4 v) {
  vec4 r;
  r.x = ((m.c[0].x * v.x + m.c[1].x * v.y) + m.c[2].x * v.z) + m.c[3].x * v.w;
  r.y = ((m.c[0].y * v.x + m.c[1].y * v.y) + m.c[2].y * v.z) + m.c[3].y * v.w;
  r.z = ((m.c[0].z * v.x + m.c[1].z * v.y) + m.c[2].z * v.z) + m.c[3].z * v.w;
  r.w = ((m.c[0].w * v.x + m.c[1].w * v.y) + m.c[2].w * v.z) + m.c[3].w * v.w;
  return r;
}
inline mat4 operator*(const mat4& a, const mat4& b) {
  mat4 r;
  for (int j = 0; j < 4; j++) r.c[j] = a * b.c[j];
  return r;
}
struct mat4x3 {  // 4 columns of 3 rows (gl_WorldToObjectEXT, gl_ObjectToWorldEXT)
  vec3 c[4];
};

// ---- built-in functions -----------------------------------------------------------------------
inline float abs(float x) { return std::fabs(x); }
inline vec3 abs(vec3 v) { return vec3(std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float exp2(float x) { return std::exp2(x); }
inline float log2(float x) { return std::log2(x); }
template <class A, class B> inline float max(A a, B b) { return std::fmax(float(a), float(b)); }
template <class A, class B> inline float min(A a, B b) { return std::fmin(float(a), float(b)); }
template <class A, class B, class C> inline float clamp(A x, B lo, C hi) {
  return std::fmin(std::fmax(float(x), float(lo)), float(hi));
}
inline float pow(float x, float y) { return std::exp2(y * std::log2(std::fmax(x, 0.0f))); }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) {
  return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 reflect(vec3 I, vec3 N) { return I - (2.0f * dot(N, I)) * N; }
inline vec3 refract(vec3 I, vec3 N, float eta) {
  float NdotI = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - NdotI * NdotI);
  if (k < 0.0f) return vec3(0.0f);
  return eta * I - (eta * NdotI + std::sqrt(k)) * N;
}
inline vec4 operator*(vec3 v, const mat4x3& m) {  // row vector times matrix: one dot per column
  return vec4(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2]), dot(v, m.c[3]));
}
template <class T> inline T nonuniformEXT(T x) { return x; }

// ---- resources ----------------------------------------------------------------------------------
struct Pipeline;  // the host side (kf_ref_host.cpp)

struct accelerationStructureEXT {
  int unused;
};
struct image2D {  // rgba32f storage image
  float* texels;
  int w, h;
};
struct sampler2D {
  Pipeline* pl;
  int index;
};
struct samplerCube {
  Pipeline* pl;
};
vec4 texture(const sampler2D& s, vec2 uv);  // black box: the sampler hardware
vec4 texture(const samplerCube& s, vec3 dir);
inline void imageStore(const image2D& im, ivec2 p, vec4 v) {
  float* t = im.texels + 4 * (size_t(p.y) * im.w + p.x);
  t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
}
inline vec4 imageLoad(const image2D& im, ivec2 p) {
  const float* t = im.texels + 4 * (size_t(p.y) * im.w + p.x);
  return vec4(t[0], t[1], t[2], t[3]);
}

// What the ray-tracing pipeline hands to one shader invocation.
struct Invocation {
  uvec3 launchID, launchSize;
  int instanceID = -1, primitiveID = -1;
  uint hitKind = 0;
  float hitT = 0.0f;
  vec3 worldRayOrigin, worldRayDirection;
  mat4x3 worldToObject;
};

struct StageBase {
  Pipeline* pl;
  void* payloadIn;
  // built-in variables (GLSL_EXT_ray_tracing)
  uvec3 gl_LaunchIDEXT, gl_LaunchSizeEXT;
  int gl_InstanceID, gl_PrimitiveID;
  uint gl_HitKindEXT;
  float gl_HitTEXT;
  vec3 gl_WorldRayOriginEXT, gl_WorldRayDirectionEXT;
  mat4x3 gl_WorldToObjectEXT;
  static constexpr uint gl_RayFlagsNoneEXT = 0u, gl_RayFlagsOpaqueEXT = 1u, gl_RayFlagsNoOpaqueEXT = 2u,
                        gl_RayFlagsTerminateOnFirstHitEXT = 4u, gl_RayFlagsSkipClosestHitShaderEXT = 8u;
  static constexpr uint gl_HitKindFrontFacingTriangleEXT = 0xFEu, gl_HitKindBackFacingTriangleEXT = 0xFFu;
  bool ignoreIntersection_ = false;

  StageBase(Pipeline* p, void* payload, const Invocation& iv)
      : pl(p), payloadIn(payload), gl_LaunchIDEXT(iv.launchID), gl_LaunchSizeEXT(iv.launchSize),
        gl_InstanceID(iv.instanceID), gl_PrimitiveID(iv.primitiveID), gl_HitKindEXT(iv.hitKind),
        gl_HitTEXT(iv.hitT), gl_WorldRayOriginEXT(iv.worldRayOrigin),
        gl_WorldRayDirectionEXT(iv.worldRayDirection), gl_WorldToObjectEXT(iv.worldToObject) {}
  virtual ~StageBase() {}
  virtual void* payloadAt(int) { return nullptr; }

  const void* bindingPtr(int set, int binding) const;  // descriptor set lookup
  const void* pushConstantPtr() const;
  uint64_t clockARB();                                 // GL_ARB_shader_clock (deterministic surrogate)
  void traceRayEXT(const accelerationStructureEXT& tlas, uint rayFlags, uint cullMask, uint sbtRecordOffset,
                   uint sbtRecordStride, uint missIndex, vec3 origin, float tMin, vec3 direction, float tMax,
                   int payloadLocation);
};

}  // namespace glsl
