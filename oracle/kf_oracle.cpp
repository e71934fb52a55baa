// kf_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A scalar C++ restatement of the reference's path-tracing shaders, used only as the checker in
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing in
// kuafu_b200/ may include, link or call this file.
//
// PARITY PINNED against the reference's own shader code.  The reference ships no tests, golden
// vectors or known-answer values for this path and its Vulkan-RT pipeline
// cannot run here, but its shading code is plain GLSL: `make -C oracle ref` compiles
// /root/reference/resources/shaders/PathTrace.{rgen,rchit,rahit,rmiss}, PathTraceShadow.rmiss and
// base/*.glsl for the CPU from where they lie (ref_shim/glsl2cpp.py + glsl_shim.hpp +
// kf_ref_host.cpp -> _ref/libkf_ref.so).  tests/test_cpu_ref_pin.py requires this restatement to
// agree with those shaders BIT FOR BIT (image, albedo, normal, primary hit, ray counts) on seeded
// scenes covering every light type, textures, environment map, emissive hits, transmission,
// depth of field, Russian roulette, alpha-0 any-hit and frame accumulation, plus the five BASELINE
// configs; tests/golden/ref_*.npz freezes the shaders' outputs for machines without the reference.
// What stays outside the pin is what the reference itself delegates to driver and hardware --
// acceleration-structure traversal, the triangle test, texture filtering, clockARB() -- which
// shim and oracle share (kfo_trace, kfo_sample_*) and which the deviations below define:
//   D1  clockARB() seeds (PathTrace.rgen:23,32) are replaced by a deterministic surrogate:
//       pixel stream  = tea(pixel, clockBase), sample stream i = tea(pixel, clockBase + 1 + i).
//   D2  triangle facing (gl_HitKindEXT) is "front <=> counter-clockwise seen from the ray origin in
//       object space, right-handed" (det > 0 in the Moller-Trumbore test below).
//   D3  equal-t ties resolve to the lowest (instance, primitive) pair; hardware order is undefined.
//   D4  textures are sampled with float bilinear weights (hardware uses 8-bit fixed point); the
//       cube map is filtered inside one face with clamp-to-edge (hardware filters across seams).
//   D5  the stochastic any-hit test for 0 < alpha < 1 (PathTrace.rahit:30-48) draws its random
//       number from a hash of (ray.seed, instance, primitive) instead of advancing ray.seed in
//       hardware traversal order, which is implementation-defined (the shim runs the real
//       PathTrace.rahit; the two agree statistically, tests/test_cpu_ref_pin.py).
//   D6  (the CUDA path only; mirrored here on request, kfo_set_skip_own_instance, to pin it) a bounce or
//       occlusion ray that leaves a convex geometry to the front side of the triangle it starts on, by more
//       than 1e-3 rad, does not enter the instance it starts on.  In exact arithmetic it cannot hit it;
//       tests/test_cpu_oracle.py renders with and without and compares.
//
// Arithmetic contract (shared with the CUDA kernels so that hit buffers can be bit-exact):
// IEEE-754 binary32, round-to-nearest, NO fused multiply-add in anything transcribed from the shaders
// (compile with -ffp-contract=off), dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z,
// normalize(v) = v * (1 / sqrt(dot(v,v))), mat*vec sums left to right.  The traversal black box
// (world -> object transform, triangle test) uses explicit fused multiply-adds, see fdot().  Transcendentals (sin, cos, pow, exp2, tan) are libm and are the
// reason radiance parity is toleranced rather than bit-exact.
//
// Every function cites the reference file:line it follows (paths relative to the reference root).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace kfo {

// ------------------------------------------------------------------------------------------------
// GLSL-style vector helpers
// ------------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
static inline V3 v3(float a) { return {a, a, a}; }
static inline V3 v3(float a, float b, float c) { return {a, b, c}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
static inline V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
static inline V3& operator*=(V3& a, float s) { a = a * s; return a; }
static inline bool allEq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }  // GLSL ==
static inline bool anyNe(V3 a, V3 b) { return a.x != b.x || a.y != b.y || a.z != b.z; }  // GLSL !=
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline float length(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(V3 a) {
  float inv = 1.0f / std::sqrt(dot(a, a));
  return a * inv;
}
static inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
static inline V3 reflect(V3 I, V3 N) { return I - (2.0f * dot(N, I)) * N; }
// GLSL refract(): returns vec3(0) on total internal reflection.
static inline V3 refract(V3 I, V3 N, float eta) {
  float NdotI = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - NdotI * NdotI);
  if (k < 0.0f) return v3(0.0f);
  return eta * I - (eta * NdotI + std::sqrt(k)) * N;
}

static const float M_PI_F = 3.141592f;  // base/Random.glsl:1  (#define M_PI 3.141592)

// ------------------------------------------------------------------------------------------------
// RNG  -- base/Random.glsl:7-39
// ------------------------------------------------------------------------------------------------
static inline uint32_t tea(uint32_t val0, uint32_t val1) {  // Random.glsl:7-21
  uint32_t v0 = val0, v1 = val1, s0 = 0;
  for (uint32_t n = 0; n < 16; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
static inline uint32_t lcg(uint32_t& prev) {  // Random.glsl:26-32
  prev = 1664525u * prev + 1013904223u;
  return prev & 0x00FFFFFFu;
}
static inline float rnd(uint32_t& prev) {  // Random.glsl:36-39
  return float(lcg(prev)) / float(0x01000000);
}

// ------------------------------------------------------------------------------------------------
// Sampling + microfacet terms -- base/Random.glsl:58-136, base/Sampling.glsl:6-35,72-83,124-134
// ------------------------------------------------------------------------------------------------
static inline V3 getPerpendicularVector(V3 u) {  // Random.glsl:58-65
  V3 a = {std::fabs(u.x), std::fabs(u.y), std::fabs(u.z)};
  uint32_t xm = ((a.x - a.y) < 0 && (a.x - a.z) < 0) ? 1 : 0;
  uint32_t ym = (a.y - a.z) < 0 ? (1u ^ xm) : 0;
  uint32_t zm = 1u ^ (xm | ym);
  return cross(u, v3(float(xm), float(ym), float(zm)));
}
static inline float Schlick(float cosine, float ior) {  // Random.glsl:68-73
  float r0 = (1.0f - ior) / (1.0f + ior);
  r0 *= r0;
  // GLSL pow(x, y) = exp2(y * log2(x)) (Vulkan precision table), undefined for x < 0: clamped to 0
  return r0 + (1.0f - r0) * std::exp2(5.0f * std::log2(std::fmax(1.0f - cosine, 0.0f)));
}
static inline float ggxNormalDistribution(float NdotH, float a2) {  // Random.glsl:75-79
  float d = std::fmax(NdotH * NdotH * (a2 - 1) + 1, 1e-6f);
  return a2 / (d * d * M_PI_F);
}
static inline V3 sampleGGX(uint32_t& seed, float a2, V3 N) {  // Random.glsl:83-101
  float rx = rnd(seed);  // GLSL evaluates constructor arguments left to right
  float ry = rnd(seed);
  V3 B = getPerpendicularVector(N);
  V3 T = cross(B, N);
  float cosThetaH = std::sqrt(clampf((1.0f - rx) / ((a2 - 1.0f) * rx + 1), 0, 1));
  float sinThetaH = std::sqrt(clampf(1.0f - cosThetaH * cosThetaH, 0, 1));
  float phiH = ry * M_PI_F * 2.0f;
  return T * (sinThetaH * std::cos(phiH)) + B * (sinThetaH * std::sin(phiH)) + N * cosThetaH;
}
static inline float G1(float dotValue, float a2) {  // Random.glsl:117-121
  return (2 * dotValue) / (dotValue + std::sqrt(a2 + (1 - a2) * (dotValue * dotValue)));
}
static inline float GeometricShadowing(float NdotV, float NdotL, float a2) {  // Random.glsl:123-127
  return G1(NdotL, a2) * G1(NdotV, a2);
}
static inline V3 schlickFresnel(V3 f0, float lDotH) {  // Random.glsl:133-136
  return f0 + (v3(1.0f) - f0) * std::exp2((-5.55473f * lDotH - 6.98316f) * lDotH);
}
static inline V3 transformLocalToWorld(V3 direction, V3 normal) {  // Sampling.glsl:6-23
  V3 tangent;
  if (std::fabs(normal.x) > std::fabs(normal.y)) {
    tangent = v3(normal.z, 0, -normal.x) / std::sqrt(normal.x * normal.x + normal.z * normal.z);
  } else {
    tangent = v3(0, -normal.z, normal.y) / std::sqrt(normal.y * normal.y + normal.z * normal.z);
  }
  V3 bitangent = cross(normal, tangent);
  return direction.x * tangent + direction.y * bitangent + direction.z * normal;
}
static inline V3 cosineHemisphereSampling(uint32_t& seed, V3 normal) {  // Sampling.glsl:25-35
  float u0 = rnd(seed);
  float u1 = rnd(seed);
  float sq = std::sqrt(1.0f - u1);
  V3 direction = v3(std::cos(2 * M_PI_F * u0) * sq, std::sin(2 * M_PI_F * u0) * sq, std::sqrt(u1));
  return transformLocalToWorld(direction, normal);
}
static inline V3 uniformSphereSampling(uint32_t& seed) {  // Sampling.glsl:72-83
  V3 p;
  do {
    float a = rnd(seed), b = rnd(seed), c = rnd(seed);
    p = v3(a, b, c) * 2.0f - v3(1.0f);
  } while (dot(p, p) >= 1.0f);
  return p;
}
static inline void diskSampling(uint32_t& seed, float& px, float& py) {  // Sampling.glsl:124-134
  do {
    float a = rnd(seed), b = rnd(seed);
    px = 2.0f * a - 1.0f;
    py = 2.0f * b - 1.0f;
  } while (px * px + py * py >= 1.0f);
}

// ------------------------------------------------------------------------------------------------
// Wire structs (mirrors of the reference's UBO/SSBO layouts; sizes asserted)
// ------------------------------------------------------------------------------------------------
struct Vertex {  // include/core/context/vertex.hpp:28-33
  float pos[3], normal[3], color[3], uv[2], pad;
};
struct Material {  // include/core/geometry.hpp:70-87
  float diffuse[4], emission[4];
  float alpha, metallic, specular, roughness, ior, transmission;
  int32_t diffuseTexIdx, metallicTexIdx, roughnessTexIdx, transmissionTexIdx, pad0, pad1;
};
struct Instance {  // include/core/geometry.hpp:91-99
  float transform[16];
  uint32_t geometryIndex, pad[3];
};
struct Camera {  // include/core/camera.hpp:181-192
  float view[16], proj[16], viewInverse[16], projInverse[16], position[4], front[4], pad[8];
};
struct DirLight {  // include/core/light.hpp:41-44
  float direction[4], rgbs[4];
};
struct PointLights {  // include/core/light.hpp:46-49
  float posr[32][4], rgbs[32][4];
};
struct ActiveLights {  // include/core/light.hpp:51-58
  float viewMat[8][16], projMat[8][16], front[8][4], rgbs[8][4], position[8][4], sftp[8][4];
};
struct PushConstants {  // include/core/rt/rt.hpp:31-42
  float clearColor[4];
  int32_t frameCount;
  uint32_t spp, maxPathDepth, useEnvironmentMap, russianRoulette, russianRouletteMinBounces, nee,
      neeMin;
};
static_assert(sizeof(Vertex) == 48 && sizeof(Material) == 80 && sizeof(Instance) == 80, "layout");
static_assert(sizeof(Camera) == 320 && sizeof(DirLight) == 32 && sizeof(PointLights) == 1024, "");
static_assert(sizeof(ActiveLights) == 1536 && sizeof(PushConstants) == 48, "layout");

struct Box {
  float lo[3], hi[3];
  void reset() {
    for (int k = 0; k < 3; k++) { lo[k] = 3.0e38f; hi[k] = -3.0e38f; }
  }
  void grow(const float* p) {
    for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], p[k]); hi[k] = std::fmax(hi[k], p[k]); }
  }
  void grow(const Box& b) { grow(b.lo); grow(b.hi); }
  float area() const {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
  }
};

// A plain binary BVH over boxes (test-infrastructure quality: binned SAH, 16 bins).
struct Bvh {
  struct Node {
    Box box;
    int32_t left, right;  // children, or (-1, -) for leaves
    int32_t first, count; // leaf range into `order`
  };
  std::vector<Node> nodes;
  std::vector<uint32_t> order;

  void build(const std::vector<Box>& boxes, int leafSize) {
    nodes.clear();
    order.resize(boxes.size());
    for (size_t i = 0; i < boxes.size(); i++) order[i] = uint32_t(i);
    if (boxes.empty()) return;
    nodes.reserve(2 * boxes.size());
    std::vector<float> cen(3 * boxes.size());
    for (size_t i = 0; i < boxes.size(); i++)
      for (int k = 0; k < 3; k++) cen[3 * i + k] = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
    subdivide(boxes, cen, 0, int(boxes.size()), leafSize);
  }

 private:
  int subdivide(const std::vector<Box>& boxes, const std::vector<float>& cen, int first, int count,
                int leafSize) {
    int me = int(nodes.size());
    nodes.push_back({});
    Box b, cb;
    b.reset();
    cb.reset();
    for (int i = first; i < first + count; i++) {
      b.grow(boxes[order[i]]);
      cb.grow(&cen[3 * order[i]]);
    }
    nodes[me].box = b;
    nodes[me].left = -1;
    nodes[me].right = -1;
    nodes[me].first = first;
    nodes[me].count = count;
    if (count <= leafSize) return me;
    int axis = 0;
    float ext = -1;
    for (int k = 0; k < 3; k++)
      if (cb.hi[k] - cb.lo[k] > ext) { ext = cb.hi[k] - cb.lo[k]; axis = k; }
    int mid = first + count / 2;
    if (ext > 0) {
      const int NB = 16;
      Box bb[NB];
      int bc[NB];
      for (int i = 0; i < NB; i++) { bb[i].reset(); bc[i] = 0; }
      float scale = NB / ext;
      auto binOf = [&](uint32_t id) {
        int bi = int((cen[3 * id + axis] - cb.lo[axis]) * scale);
        return bi < 0 ? 0 : (bi >= NB ? NB - 1 : bi);
      };
      for (int i = first; i < first + count; i++) {
        int bi = binOf(order[i]);
        bb[bi].grow(boxes[order[i]]);
        bc[bi]++;
      }
      float rightArea[NB];
      Box acc;
      acc.reset();
      int rc[NB];
      int cnt = 0;
      for (int i = NB - 1; i > 0; i--) {
        if (bc[i]) acc.grow(bb[i]);
        cnt += bc[i];
        rightArea[i] = cnt ? acc.area() : 0;
        rc[i] = cnt;
      }
      acc.reset();
      cnt = 0;
      float best = 3.0e38f;
      int bestSplit = -1;
      for (int i = 0; i < NB - 1; i++) {
        if (bc[i]) acc.grow(bb[i]);
        cnt += bc[i];
        if (cnt == 0 || rc[i + 1] == 0) continue;
        float cost = acc.area() * cnt + rightArea[i + 1] * rc[i + 1];
        if (cost < best) { best = cost; bestSplit = i; }
      }
      if (bestSplit >= 0) {
        auto it = std::partition(order.begin() + first, order.begin() + first + count,
                                 [&](uint32_t id) { return binOf(id) <= bestSplit; });
        mid = int(it - order.begin());
      }
    }
    if (mid == first || mid == first + count) mid = first + count / 2;
    int l = subdivide(boxes, cen, first, mid - first, leafSize);
    int r = subdivide(boxes, cen, mid, first + count - mid, leafSize);
    nodes[me].left = l;
    nodes[me].right = r;
    return me;
  }
};

struct Tri {
  V3 v0, e1, e2;
};

struct Geometry {
  std::vector<Vertex> verts;
  std::vector<uint32_t> idx;
  std::vector<uint32_t> matIndex;
  bool opaque = true, hide = false, present = false;
  std::vector<Tri> tris;
  Bvh bvh;
  bool convex = false;  // every vertex on or behind the plane of every triangle (the CUDA path's k_batch_convex)
};

struct Texture {
  uint32_t w = 0, h = 0;
  std::vector<uint8_t> rgba;
};

struct InstanceRt {
  float inv[3][4];  // world -> object, 3 rows x 4 columns (last column = translation)
  Box worldBox;
  uint32_t geom;
};

struct Counters {
  std::atomic<uint64_t> paths{0}, extensionRays{0}, shadowRays{0}, extensionHits{0};
};

struct Scene {
  // Declared deviation D6 of the CUDA path, mirrored here on request (kfo_set_skip_own_instance) so that it can be
  // pinned on the CPU: a bounce or occlusion ray that leaves a CONVEX geometry to the front side of the triangle it
  // starts on, by more than a grazing margin, does not enter the instance it starts on.  Off by default: the
  // oracle then walks every instance, as the reference does.
  bool skipOwnInstance = false;
  // The CUDA path's light-sample culling (kf_wavefront.cuh, nextRelevantLight), mirrored on request
  // (kfo_set_cull_light_samples) so that "cannot matter" can be checked where no kernel is involved: a light
  // sample whose contribution is exactly zero and draws no random number, or -- in a scene with a single light
  // slot -- any finite sample of a path whose BSDF weight is exactly zero, is answered without a ray.
  bool cullLightSamples = false;
  uint64_t lastShadowSkipped = 0;  // light samples the last kfo_render answered without a ray (kfo_last_shadow_skipped)
  std::vector<Geometry> geoms;
  std::vector<Material> mats;
  std::vector<Texture> texs;
  Texture envFaces[6];
  bool hasEnv = false;
  std::vector<Instance> insts;
  std::vector<InstanceRt> instRt;
  Bvh tlas;
  DirLight dl{};
  PointLights pl{};
  ActiveLights al{};
  float srgbToLinear[256];
  float srgbThreshold[255];  // linear value at which the 8-bit sRGB code becomes k+1
  bool accelDirty = true;
  std::vector<const void*> viewVerts, viewIdx, viewMat;  // kfo_scene_view()
  Counters counters;
  std::string err;

  Scene() {
    // sRGB EOTF (Vulkan R8G8B8A8Srgb decode) and its inverse thresholds (B8G8R8A8Srgb encode).
    for (int i = 0; i < 256; i++) {
      double c = i / 255.0;
      srgbToLinear[i] = float(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
    }
    for (int k = 0; k < 255; k++) {
      double c = (k + 0.5) / 255.0;
      srgbThreshold[k] = float(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
    }
  }
};

// world->object of an instance.  The reference hands glm::transpose(transform) (3 rows) to the
// driver (src/core/rt/rt.cpp:127-138); the driver's inverse is unspecified, this float cofactor
// inverse is the contract the CUDA refit kernel follows too.
static void affineInverse(const float* m /*col-major 4x4*/, float inv[3][4]) {
  float a00 = m[0], a10 = m[1], a20 = m[2];
  float a01 = m[4], a11 = m[5], a21 = m[6];
  float a02 = m[8], a12 = m[9], a22 = m[10];
  float t0 = m[12], t1 = m[13], t2 = m[14];
  float c00 = a11 * a22 - a12 * a21;
  float c01 = a12 * a20 - a10 * a22;
  float c02 = a10 * a21 - a11 * a20;
  float det = (a00 * c00 + a01 * c01) + a02 * c02;
  float id = 1.0f / det;
  inv[0][0] = c00 * id;
  inv[1][0] = c01 * id;
  inv[2][0] = c02 * id;
  inv[0][1] = (a02 * a21 - a01 * a22) * id;
  inv[1][1] = (a00 * a22 - a02 * a20) * id;
  inv[2][1] = (a01 * a20 - a00 * a21) * id;
  inv[0][2] = (a01 * a12 - a02 * a11) * id;
  inv[1][2] = (a02 * a10 - a00 * a12) * id;
  inv[2][2] = (a00 * a11 - a01 * a10) * id;
  for (int r = 0; r < 3; r++) inv[r][3] = -((inv[r][0] * t0 + inv[r][1] * t1) + inv[r][2] * t2);
}

static inline void padBox(Box& b) {
  // Conservative margin so that box culling can never reject a triangle the float
  // Moller-Trumbore test would accept (see DESIGN.md "bit-exact hits").
  float m = 0;
  for (int k = 0; k < 3; k++) m = std::fmax(m, std::fmax(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
  float pad = m * (1.0f / 16384.0f) + 1e-30f;
  for (int k = 0; k < 3; k++) { b.lo[k] -= pad; b.hi[k] += pad; }
}

static void buildAccel(Scene& s) {
  for (auto& g : s.geoms) {
    if (!g.present || !g.tris.empty() || g.idx.empty()) continue;
    size_t nt = g.idx.size() / 3;
    g.tris.resize(nt);
    std::vector<Box> boxes(nt);
    for (size_t t = 0; t < nt; t++) {
      const float* p0 = g.verts[g.idx[3 * t + 0]].pos;
      const float* p1 = g.verts[g.idx[3 * t + 1]].pos;
      const float* p2 = g.verts[g.idx[3 * t + 2]].pos;
      V3 a = v3(p0[0], p0[1], p0[2]), b = v3(p1[0], p1[1], p1[2]), c = v3(p2[0], p2[1], p2[2]);
      g.tris[t] = {a, b - a, c - a};
      boxes[t].reset();
      boxes[t].grow(p0);
      boxes[t].grow(p1);
      boxes[t].grow(p2);
      padBox(boxes[t]);
    }
    g.bvh.build(boxes, 4);
    // convexity, the criterion of kf_blas_batch.cuh (k_batch_convex): tolerance 1e-5 of the extent
    g.convex = false;
    if (uint64_t(nt) * g.verts.size() <= (uint64_t(1) << 28) && !g.bvh.nodes.empty()) {
      const Box& gb = g.bvh.nodes[0].box;
      const float ext = std::fmax(gb.hi[0] - gb.lo[0], std::fmax(gb.hi[1] - gb.lo[1], gb.hi[2] - gb.lo[2]));
      bool ok = true;
      for (size_t t = 0; t < nt && ok; t++) {
        const Tri& tr = g.tris[t];
        const float nx = tr.e1.y * tr.e2.z - tr.e1.z * tr.e2.y, ny = tr.e1.z * tr.e2.x - tr.e1.x * tr.e2.z,
                    nz = tr.e1.x * tr.e2.y - tr.e1.y * tr.e2.x;
        const float tol = 1e-5f * ext * std::sqrt(nx * nx + ny * ny + nz * nz);
        for (size_t v = 0; v < g.verts.size() && ok; v++) {
          const float* q = g.verts[v].pos;
          ok = nx * (q[0] - tr.v0.x) + ny * (q[1] - tr.v0.y) + nz * (q[2] - tr.v0.z) <= tol;
        }
      }
      g.convex = ok;
    }
  }
  s.instRt.resize(s.insts.size());
  std::vector<Box> ib(s.insts.size());
  for (size_t i = 0; i < s.insts.size(); i++) {
    InstanceRt& r = s.instRt[i];
    r.geom = s.insts[i].geometryIndex;
    affineInverse(s.insts[i].transform, r.inv);
    Box wb;
    wb.reset();
    const Geometry* g = r.geom < s.geoms.size() ? &s.geoms[r.geom] : nullptr;
    if (g && g->present && !g->hide && !g->bvh.nodes.empty()) {
      const Box& ob = g->bvh.nodes[0].box;
      const float* m = s.insts[i].transform;
      for (int c = 0; c < 8; c++) {
        float x = (c & 1) ? ob.hi[0] : ob.lo[0];
        float y = (c & 2) ? ob.hi[1] : ob.lo[1];
        float z = (c & 4) ? ob.hi[2] : ob.lo[2];
        float p[3];
        for (int k = 0; k < 3; k++) p[k] = ((m[k] * x + m[4 + k] * y) + m[8 + k] * z) + m[12 + k];
        wb.grow(p);
      }
      padBox(wb);
    } else {
      for (int k = 0; k < 3; k++) { wb.lo[k] = 3e38f; wb.hi[k] = 3e38f; }  // unreachable
    }
    r.worldBox = wb;
    ib[i] = wb;
  }
  s.tlas.build(ib, 1);
  s.accelDirty = false;
}

// ------------------------------------------------------------------------------------------------
// Ray / scene intersection (what the reference delegates to traceRayEXT + the driver's BVH)
// ------------------------------------------------------------------------------------------------
struct Hit {
  float t;
  float u, v;
  int32_t inst, prim;
  bool front;
};

// Fused arithmetic of the traversal black box.  What traceRayEXT does inside the driver -- the
// world -> object ray transform and the ray / triangle test -- is not shader code, so its arithmetic is
// ours to define; it is defined with fused multiply-adds (one rounding per fma, std::fma here,
// __fmaf_rn on the device), which is what a GPU executes natively.  Everything transcribed from the
// GLSL keeps the unfused contract of the file header.
static inline float fdot(V3 a, V3 b) { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
static inline V3 fcross(V3 a, V3 b) {
  return {std::fma(a.y, b.z, -(a.z * b.y)), std::fma(a.z, b.x, -(a.x * b.z)), std::fma(a.x, b.y, -(a.y * b.x))};
}

// Moller-Trumbore in object space, no culling (instances use TriangleFacingCullDisable,
// src/core/rt/rt.cpp:134).  The operation order below is the bit-exactness contract.
static inline bool intersectTri(const Tri& tr, V3 o, V3 d, float& t, float& u, float& v,
                                float& det) {
  V3 p = fcross(d, tr.e2);
  det = fdot(tr.e1, p);
  if (det == 0.0f) return false;
  float inv = 1.0f / det;
  V3 tv = o - tr.v0;
  u = fdot(tv, p) * inv;
  if (!(u >= 0.0f && u <= 1.0f)) return false;
  V3 q = fcross(tv, tr.e1);
  v = fdot(d, q) * inv;
  if (!(v >= 0.0f && u + v <= 1.0f)) return false;
  t = fdot(tr.e2, q) * inv;
  return true;
}

static inline bool slab(const Box& b, V3 o, V3 id, float tmin, float tmax) {
  float t0 = tmin, t1 = tmax;
  const float oo[3] = {o.x, o.y, o.z}, ii[3] = {id.x, id.y, id.z};
  for (int k = 0; k < 3; k++) {
    float a = (b.lo[k] - oo[k]) * ii[k];
    float c = (b.hi[k] - oo[k]) * ii[k];
    float lo = std::fmin(a, c), hi = std::fmax(a, c);  // fmin/fmax drop NaN (0 * inf)
    t0 = std::fmax(t0, lo);
    t1 = std::fmin(t1, hi);
  }
  return t0 <= t1 * 1.0000005f + 1e-30f;
}

static inline float hashRnd(uint32_t seed, uint32_t inst, uint32_t prim) {  // deviation D5
  uint32_t h = tea(seed ^ (prim * 0x9e3779b9u), inst);
  return float(h & 0x00FFFFFFu) / float(0x01000000);
}

typedef int (*AnyHitFn)(void* user, int32_t inst, int32_t prim);  // 1 = accept the candidate

struct Tracer {
  const Scene& s;
  bool brute;
  // When set, candidates of non-opaque geometry are put to this callback instead of the D5 hash
  // draw: oracle/ref_shim runs the reference's own PathTrace.rahit through it.
  AnyHitFn anyHitFn = nullptr;
  void* anyHitUser = nullptr;
  explicit Tracer(const Scene& sc, bool b) : s(sc), brute(b) {}

  // PathTrace.rahit:30-48 on a candidate of a non-opaque geometry.
  bool anyHitAccepts(uint32_t seed, uint32_t inst, uint32_t prim, const Geometry& g) const {
    uint32_t mi = g.matIndex[prim];
    float alpha = s.mats[mi].alpha;
    if (alpha == 0.0f) return false;
    if (hashRnd(seed, inst, prim) > alpha) return false;
    return true;
  }

  inline void testTri(const Geometry& g, uint32_t inst, uint32_t prim, V3 o, V3 d, float tmin,
                      bool anyHit, uint32_t seed, Hit& best, bool& found) const {
    float t, u, v, det;
    if (!intersectTri(g.tris[prim], o, d, t, u, v, det)) return;
    if (!(t > tmin)) return;
    bool closer = t < best.t ||
                  (t == best.t && found &&
                   (int32_t(inst) < best.inst || (int32_t(inst) == best.inst && int32_t(prim) < best.prim)));
    if (!closer) return;
    if (anyHit && !g.opaque) {
      if (anyHitFn ? !anyHitFn(anyHitUser, int32_t(inst), int32_t(prim)) : !anyHitAccepts(seed, inst, prim, g)) return;
    }
    best = {t, u, v, int32_t(inst), int32_t(prim), det > 0.0f};
    found = true;
  }

  void traceInstance(uint32_t i, V3 o, V3 d, float tmin, bool anyHit, uint32_t seed, Hit& best,
                     bool& found, bool terminateOnFirst) const {
    const InstanceRt& r = s.instRt[i];
    if (r.geom >= s.geoms.size()) return;
    const Geometry& g = s.geoms[r.geom];
    if (!g.present || g.hide || g.tris.empty()) return;
    V3 oo, od;
    // world -> object, fused (see fdot): row . o + translation, row . d
    oo.x = std::fma(r.inv[0][2], o.z, std::fma(r.inv[0][1], o.y, std::fma(r.inv[0][0], o.x, r.inv[0][3])));
    oo.y = std::fma(r.inv[1][2], o.z, std::fma(r.inv[1][1], o.y, std::fma(r.inv[1][0], o.x, r.inv[1][3])));
    oo.z = std::fma(r.inv[2][2], o.z, std::fma(r.inv[2][1], o.y, std::fma(r.inv[2][0], o.x, r.inv[2][3])));
    od.x = std::fma(r.inv[0][2], d.z, std::fma(r.inv[0][1], d.y, r.inv[0][0] * d.x));
    od.y = std::fma(r.inv[1][2], d.z, std::fma(r.inv[1][1], d.y, r.inv[1][0] * d.x));
    od.z = std::fma(r.inv[2][2], d.z, std::fma(r.inv[2][1], d.y, r.inv[2][0] * d.x));
    if (brute) {
      for (uint32_t p = 0; p < g.tris.size(); p++) {
        testTri(g, i, p, oo, od, tmin, anyHit, seed, best, found);
        if (terminateOnFirst && found) return;
      }
      return;
    }
    V3 id = {1.0f / od.x, 1.0f / od.y, 1.0f / od.z};
    int stack[96];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
      const Bvh::Node& n = g.bvh.nodes[stack[--sp]];
      if (!slab(n.box, oo, id, tmin, best.t)) continue;
      if (n.left < 0) {
        for (int k = n.first; k < n.first + n.count; k++) {
          testTri(g, i, g.bvh.order[k], oo, od, tmin, anyHit, seed, best, found);
          if (terminateOnFirst && found) return;
        }
      } else {
        stack[sp++] = n.right;
        stack[sp++] = n.left;
      }
    }
  }

  // traceRayEXT equivalent.  tmax is exclusive (a hit needs tmin < t < tmax).
  bool trace(V3 o, V3 d, float tmin, float tmax, bool anyHit, uint32_t seed, bool terminateOnFirst,
             Hit& out, int32_t skipInst = -1) const {
    Hit best{};
    best.t = tmax;
    best.inst = -1;
    best.prim = -1;
    bool found = false;
    // "found" with t == tmax must not count: hits need t < tmax.  testTri accepts t < best.t only
    // while !found, so the initial best.t = tmax enforces the exclusive bound.
    if (brute || s.tlas.nodes.empty()) {
      for (uint32_t i = 0; i < s.instRt.size(); i++) {
        if (int32_t(i) == skipInst) continue;
        traceInstance(i, o, d, tmin, anyHit, seed, best, found, terminateOnFirst);
        if (terminateOnFirst && found) break;
      }
    } else {
      V3 id = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
      int stack[96];
      int sp = 0;
      stack[sp++] = 0;
      while (sp) {
        const Bvh::Node& n = s.tlas.nodes[stack[--sp]];
        if (!slab(n.box, o, id, tmin, best.t)) continue;
        if (n.left < 0) {
          for (int k = n.first; k < n.first + n.count; k++) {
            if (int32_t(s.tlas.order[k]) == skipInst) continue;
            traceInstance(s.tlas.order[k], o, d, tmin, anyHit, seed, best, found, terminateOnFirst);
            if (terminateOnFirst && found) break;
          }
          if (terminateOnFirst && found) break;
        } else {
          stack[sp++] = n.right;
          stack[sp++] = n.left;
        }
      }
    }
    out = best;
    return found;
  }
};

// ------------------------------------------------------------------------------------------------
// Textures  (vkCore.hpp:580-637: R8G8B8A8Srgb, linear filter, repeat, LOD 0)
// ------------------------------------------------------------------------------------------------
static inline float wrap01(float u) {
  float f = u - std::floor(u);
  if (!(f >= 0.0f && f <= 1.0f)) f = 0.0f;
  return f;
}
static inline void texel(const Scene& s, const Texture& t, int x, int y, float out[3]) {
  const uint8_t* p = &t.rgba[(size_t(y) * t.w + x) * 4];
  out[0] = s.srgbToLinear[p[0]];
  out[1] = s.srgbToLinear[p[1]];
  out[2] = s.srgbToLinear[p[2]];
}
static inline V3 bilinear(const Scene& s, const Texture& t, float x, float y, bool repeat) {
  float x0f = std::floor(x), y0f = std::floor(y);
  float fx = x - x0f, fy = y - y0f;
  int x0 = int(x0f), y0 = int(y0f), x1 = x0 + 1, y1 = y0 + 1;
  int W = int(t.w), H = int(t.h);
  if (repeat) {
    if (x0 < 0) x0 += W;
    if (y0 < 0) y0 += H;
    if (x1 >= W) x1 -= W;
    if (y1 >= H) y1 -= H;
  } else {
    x0 = std::max(0, std::min(W - 1, x0));
    x1 = std::max(0, std::min(W - 1, x1));
    y0 = std::max(0, std::min(H - 1, y0));
    y1 = std::max(0, std::min(H - 1, y1));
  }
  float a[3], b[3], c[3], d[3];
  texel(s, t, x0, y0, a);
  texel(s, t, x1, y0, b);
  texel(s, t, x0, y1, c);
  texel(s, t, x1, y1, d);
  float r[3];
  for (int k = 0; k < 3; k++)
    r[k] = (a[k] * (1.0f - fx) + b[k] * fx) * (1.0f - fy) + (c[k] * (1.0f - fx) + d[k] * fx) * fy;
  return v3(r[0], r[1], r[2]);
}
static V3 sampleTexture(const Scene& s, int idx, float u, float v) {  // texture(textures[i], uv)
  if (idx < 0 || size_t(idx) >= s.texs.size() || s.texs[idx].w == 0) return v3(0.0f);
  const Texture& t = s.texs[idx];
  float x = wrap01(u) * float(t.w) - 0.5f;
  float y = wrap01(v) * float(t.h) - 0.5f;
  return bilinear(s, t, x, y, true);
}
static V3 sampleCube(const Scene& s, V3 r) {  // texture(samplerCube, dir), Vulkan face selection
  float ax = std::fabs(r.x), ay = std::fabs(r.y), az = std::fabs(r.z);
  int face;
  float sc, tc, ma;
  if (az >= ax && az >= ay) {
    face = r.z < 0 ? 5 : 4;
    sc = r.z < 0 ? -r.x : r.x;
    tc = -r.y;
    ma = az;
  } else if (ay >= ax) {
    face = r.y < 0 ? 3 : 2;
    sc = r.x;
    tc = r.y < 0 ? -r.z : r.z;
    ma = ay;
  } else {
    face = r.x < 0 ? 1 : 0;
    sc = r.x < 0 ? r.z : -r.z;
    tc = -r.y;
    ma = ax;
  }
  const Texture& t = s.envFaces[face];
  if (t.w == 0) return v3(0.0f);
  float u = 0.5f * (sc / ma + 1.0f), v = 0.5f * (tc / ma + 1.0f);
  if (!(u >= 0.0f && u <= 1.0f)) u = 0.0f;
  if (!(v >= 0.0f && v <= 1.0f)) v = 0.0f;
  return bilinear(s, t, u * float(t.w) - 0.5f, v * float(t.h) - 0.5f, false);
}

static inline void mulMat4(const float* m, float x, float y, float z, float w, float out[4]) {
  for (int r = 0; r < 4; r++) out[r] = ((m[r] * x + m[4 + r] * y) + m[8 + r] * z) + m[12 + r] * w;
}

// ------------------------------------------------------------------------------------------------
// Shader transcription
// ------------------------------------------------------------------------------------------------
struct RayPayLoad {  // base/Ray.glsl:1-14
  V3 direction, albedo, normal, emission, origin, weight;
  uint32_t seed, depth, type;
  V3 shadow_color;
  bool refractive;
  int32_t skipInst = -1;  // D6 (Scene::skipOwnInstance): the instance the NEXT extension ray does not enter
};

struct Shader {
  const Scene& s;
  const PushConstants& pc;
  Tracer tr;
  uint64_t extRays = 0, shRays = 0, extHits = 0, shSkipped = 0;
  bool deadPath = false;  // cullLightSamples: the BSDF sample of the hit being shaded has weight exactly zero
  Shader(const Scene& sc, const PushConstants& p, bool brute) : s(sc), pc(p), tr(sc, brute) {}
  // the number of light slots that are switched on, by the CUDA path's conditions (kfrtSetLights)
  uint32_t lightSlots() const {
    uint32_t n = 0;
    if (s.dl.rgbs[0] * s.dl.rgbs[3] != 0 || s.dl.rgbs[1] * s.dl.rgbs[3] != 0 || s.dl.rgbs[2] * s.dl.rgbs[3] != 0) n++;
    for (int i = 0; i < 32; i++) n += s.pl.rgbs[i][3] > 0 ? 1u : 0u;
    for (int i = 0; i < 8; i++) n += s.al.front[i][3] > 0 ? 1u : 0u;
    return n;
  }

  // PathTrace.rmiss:12-31
  void miss(RayPayLoad& ray) {
    V3 dir = ray.direction;
    dir = v3(-dir.y, dir.z, -dir.x);
    if (pc.useEnvironmentMap)
      ray.emission = sampleCube(s, dir);
    else
      ray.emission = v3(pc.clearColor[0], pc.clearColor[1], pc.clearColor[2]) * pc.clearColor[3];
    ray.depth = pc.maxPathDepth + 1;
  }

  struct HitCtx {
    bool backFacing;
    V3 rayDirection;  // ray.direction at hit time (not yet overwritten)
    // D6 (Scene::skipOwnInstance): instance of the hit and the geometric normal e1 x e2 of its triangle taken to
    // world space like a shading normal; inst < 0 when the option is off or the geometry is not convex
    int32_t inst = -1;
    V3 Ng = {0.0f, 0.0f, 0.0f};
  };
  // the CUDA path's leavesSurface(): to the front side of the triangle, by more than a grazing margin
  static bool leavesSurface(V3 Ng, V3 d) {
    const float sd = dot(Ng, d);
    return sd > 0.0f && sd * sd > 1e-6f * dot(Ng, Ng) * dot(d, d);
  }

  // PathTrace.rchit:103-171
  V3 calcDirectContribution(RayPayLoad& ray, const HitCtx& hc, V3 L, V3 V, V3 N, V3 lightEmission,
                            float f, float a2, V3 diffuseColor, V3 specularColor,
                            V3 transmissionColor) {
    V3 weight = v3(0.0f);
    if (anyNe(transmissionColor, v3(0.0f))) {
      bool isInside = hc.backFacing;
      if (isInside) {
        weight = v3(0.0f);
      } else {
        float NdotV = dot(N, V);
        V3 refractedL = refract(-V, N, 1 / f);
        float reflectProb = anyNe(refractedL, v3(0.0f)) ? Schlick(NdotV, f) : 1.0f;
        if (rnd(ray.seed) <= reflectProb) {
          V3 H = normalize(L + V);
          float NdotL = dot(N, L);
          float NdotH = dot(N, H);
          float HdotV = dot(H, V);
          float NdotV2 = std::fmax(dot(N, V), 1e-6f);
          float LdotH = dot(L, H);
          float D = ggxNormalDistribution(NdotH, a2);
          float G = GeometricShadowing(NdotL, NdotV2, a2);
          V3 F = schlickFresnel(transmissionColor, LdotH);
          weight = D * F * G * HdotV / NdotH * NdotV2;  // left to right: ((((D*F)*G)*HdotV)/NdotH)*NdotV
        } else {
          weight = v3(0.0f);
        }
      }
    } else {
      V3 H = normalize(L + V);
      float NdotL = dot(N, L);
      float NdotH = dot(N, H);
      float HdotV = dot(H, V);
      float NdotV = std::fmax(dot(N, V), 1e-6f);
      float LdotH = dot(L, H);
      float D = ggxNormalDistribution(NdotH, a2);
      float G = GeometricShadowing(NdotL, NdotV, a2);
      V3 F = schlickFresnel(specularColor, LdotH);
      float diffuseLum = length(diffuseColor);
      float specularLum = length(specularColor);
      float probDiffuse = diffuseLum / (diffuseLum + specularLum);
      if (diffuseLum == 0 && specularLum == 0) probDiffuse = 0.5f;
      V3 diffuseWeight = diffuseColor * v3(NdotL);
      V3 specularWeight = D * F * G * HdotV / NdotH * NdotV;
      weight = rnd(ray.seed) < probDiffuse ? diffuseWeight * probDiffuse
                                           : specularWeight * (1 - probDiffuse);
    }
    return lightEmission * weight;
  }

  // PathTrace.rchit:175-200
  V3 traceShadowRay(RayPayLoad& ray, const HitCtx& hc, V3 worldPos, V3 L, V3 V, V3 N, float maxDist,
                    V3 lightEmission, float f, float a2, V3 diffuseColor, V3 specularColor,
                    V3 transmissionColor) {
    bool isShadowed = true;
    float NdotL = dot(N, L);
    if (NdotL > 0.0f) {
      // TerminateOnFirstHit | Opaque | SkipClosestHitShader, miss index 1 (PathTraceShadow.rmiss)
      if (s.cullLightSamples) {
        RayPayLoad probe = ray;
        const V3 c = calcDirectContribution(probe, hc, L, V, N, lightEmission, f, a2, diffuseColor, specularColor,
                                            transmissionColor);
        const bool zeroNoDraw = allEq(c, v3(0.0f)) && probe.seed == ray.seed;
        const bool finite = std::isfinite(c.x) && std::isfinite(c.y) && std::isfinite(c.z);
        if (zeroNoDraw || (deadPath && finite && lightSlots() <= 1)) {
          shSkipped++;
          return v3(0.0f);
        }
      }
      Hit h;
      shRays++;
      const int32_t skip = (hc.inst >= 0 && leavesSurface(hc.Ng, L)) ? hc.inst : -1;
      isShadowed = tr.trace(worldPos, L, 0.001f, maxDist, /*anyHit=*/false, 0, true, h, skip);
    }
    return isShadowed ? v3(0.0f)
                      : calcDirectContribution(ray, hc, L, V, N, lightEmission, f, a2, diffuseColor,
                                               specularColor, transmissionColor);
  }

  // PathTrace.rchit:204-223
  V3 traceDirectionalLight(RayPayLoad& ray, const HitCtx& hc, V3 worldPos, V3 N, float f, float a2,
                           V3 diffuseColor, V3 specularColor, V3 transmissionColor) {
    V3 lightEmission = v3(s.dl.rgbs[0], s.dl.rgbs[1], s.dl.rgbs[2]) * s.dl.rgbs[3];
    if (allEq(lightEmission, v3(0.0f))) return v3(0.0f);
    V3 L = -v3(s.dl.direction[0], s.dl.direction[1], s.dl.direction[2]);
    V3 V = normalize(-hc.rayDirection);
    if (s.dl.direction[3] != 0) {
      float a = rnd(ray.seed), b = rnd(ray.seed), c = rnd(ray.seed);
      V3 perturb = v3(a, b, c);
      L = normalize(L + s.dl.direction[3] * perturb);
    }
    float maxDist = 1e6f;
    return traceShadowRay(ray, hc, worldPos, L, V, N, maxDist, lightEmission, f, a2, diffuseColor,
                          specularColor, transmissionColor);
  }

  // PathTrace.rchit:227-258
  V3 tracePointLights(RayPayLoad& ray, const HitCtx& hc, V3 worldPos, V3 N, float f, float a2,
                      V3 diffuseColor, V3 specularColor, V3 transmissionColor) {
    V3 ret = v3(0.0f);
    for (uint32_t i = 0; i < 32; ++i)
      if (s.pl.rgbs[i][3] > 0) {
        V3 rgb = v3(s.pl.rgbs[i][0], s.pl.rgbs[i][1], s.pl.rgbs[i][2]);
        float lum = length(rgb * s.pl.rgbs[i][3]);
        if (lum == 0) continue;
        V3 V = normalize(-hc.rayDirection);
        V3 lpos = v3(s.pl.posr[i][0], s.pl.posr[i][1], s.pl.posr[i][2]);
        if (s.pl.posr[i][3] != 0) {
          V3 perturb = uniformSphereSampling(ray.seed);
          lpos += s.pl.posr[i][3] * normalize(perturb);
        }
        V3 L = lpos - worldPos;
        float d = length(L);
        L = normalize(L);
        V3 lightEmission = rgb * s.pl.rgbs[i][3] / d / d;
        ret += traceShadowRay(ray, hc, worldPos, L, V, N, d, lightEmission, f, a2, diffuseColor,
                              specularColor, transmissionColor);
      }
    return ret;
  }

  // PathTrace.rchit:262-316
  V3 traceActiveLights(RayPayLoad& ray, const HitCtx& hc, V3 worldPos, V3 N, float f, float a2,
                       V3 diffuseColor, V3 specularColor, V3 transmissionColor) {
    V3 ret = v3(0.0f);
    for (uint32_t i = 0; i < 8; ++i)
      if (s.al.front[i][3] > 0) {
        V3 rgb = v3(s.al.rgbs[i][0], s.al.rgbs[i][1], s.al.rgbs[i][2]);
        float lum = length(rgb * s.al.rgbs[i][3]);
        if (lum == 0) continue;
        V3 V = normalize(-hc.rayDirection);
        float softness = s.al.sftp[i][0];
        V3 lpos = v3(s.al.position[i][0], s.al.position[i][1], s.al.position[i][2]);
        if (softness != 0) {
          V3 perturb = uniformSphereSampling(ray.seed);
          lpos += softness * normalize(perturb);
        }
        V3 L = lpos - worldPos;
        float d = length(L);
        L = normalize(L);
        float fov = s.al.sftp[i][1];
        V3 alightDir = normalize(v3(s.al.front[i][0], s.al.front[i][1], s.al.front[i][2]));
        float halfAngle = clampf(fov, 0, M_PI_F) / 2;
        float cos_ = dot(alightDir, -L);
        if (cos_ > std::cos(halfAngle)) {
          int texID = int(s.al.sftp[i][2]);
          V3 color = rgb;
          if (texID >= 0) {
            // `proj * view * vec4(worldPos, 1)` (rchit:300) is left-associative: (proj * view) * p
            float pvm[16], cp[4];
            for (int j = 0; j < 4; j++)
              mulMat4(s.al.projMat[i], s.al.viewMat[i][4 * j], s.al.viewMat[i][4 * j + 1],
                      s.al.viewMat[i][4 * j + 2], s.al.viewMat[i][4 * j + 3], pvm + 4 * j);
            mulMat4(pvm, worldPos.x, worldPos.y, worldPos.z, 1.0f, cp);
            float tu = cp[0] / cp[3], tv = cp[1] / cp[3];
            float u = tu * 0.5f + 0.5f, v = tv * 0.5f + 0.5f;
            color *= sampleTexture(s, texID, u, v);
          }
          V3 lightEmission = color * s.al.rgbs[i][3] / d / d;
          ret += traceShadowRay(ray, hc, worldPos, L, V, N, d, lightEmission, f, a2, diffuseColor,
                                specularColor, transmissionColor);
        }
      }
    return ret;
  }

  // PathTrace.rchit:66-98 (getShadingData) + :322-469 (main)
  void closestHit(RayPayLoad& ray, const Hit& h, V3 rayOrigin, V3 rayDirection) {
    const InstanceRt& ir = s.instRt[h.inst];
    uint32_t geometryIndex = s.insts[h.inst].geometryIndex;
    const Geometry& g = s.geoms[geometryIndex];
    uint32_t i0 = g.idx[3 * h.prim + 0], i1 = g.idx[3 * h.prim + 1], i2 = g.idx[3 * h.prim + 2];
    const Vertex &v0 = g.verts[i0], &v1 = g.verts[i1], &v2 = g.verts[i2];
    float bx = 1.0f - h.u - h.v, by = h.u, bz = h.v;
    V3 n0 = v3(v0.normal[0], v0.normal[1], v0.normal[2]);
    V3 n1 = v3(v1.normal[0], v1.normal[1], v1.normal[2]);
    V3 n2 = v3(v2.normal[0], v2.normal[1], v2.normal[2]);
    V3 localNormal = n0 * bx + n1 * by + n2 * bz;
    // normalize(vec3(localNormal * gl_WorldToObjectEXT)): row-vector times the 3x4 world->object
    V3 wn;
    wn.x = (localNormal.x * ir.inv[0][0] + localNormal.y * ir.inv[1][0]) + localNormal.z * ir.inv[2][0];
    wn.y = (localNormal.x * ir.inv[0][1] + localNormal.y * ir.inv[1][1]) + localNormal.z * ir.inv[2][1];
    wn.z = (localNormal.x * ir.inv[0][2] + localNormal.y * ir.inv[1][2]) + localNormal.z * ir.inv[2][2];
    V3 N = normalize(wn);
    V3 worldPos = rayOrigin + rayDirection * h.t;
    float uvx = (v0.uv[0] * bx + v1.uv[0] * by) + v2.uv[0] * bz;
    float uvy = (v0.uv[1] * bx + v1.uv[1] * by) + v2.uv[1] * bz;
    uint32_t matIndex = g.matIndex[h.prim];
    const Material& mat = s.mats[matIndex];

    V3 emission = v3(mat.emission[0], mat.emission[1], mat.emission[2]) * mat.emission[3];
    if (anyNe(emission, v3(0.0f))) {
      ray.depth = pc.maxPathDepth + 1;
      ray.emission = emission;
      return;
    }
    V3 baseColor = v3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]);
    if (mat.diffuseTexIdx >= 0) baseColor = sampleTexture(s, mat.diffuseTexIdx, uvx, uvy);
    baseColor = baseColor / M_PI_F;
    float metallic = mat.metallicTexIdx >= 0 ? sampleTexture(s, mat.metallicTexIdx, uvx, uvy).x
                                             : mat.metallic;
    float roughness = mat.roughnessTexIdx >= 0 ? sampleTexture(s, mat.roughnessTexIdx, uvx, uvy).x
                                               : mat.roughness;
    float a2 = roughness * roughness;
    float transmission = mat.transmissionTexIdx >= 0
                             ? sampleTexture(s, mat.transmissionTexIdx, uvx, uvy).x
                             : mat.transmission;
    float f = std::fmax(mat.ior, 1e-5f);
    float diffuse_weight = (1.0f - clampf(metallic, 0.0f, 1.0f)) * (1.0f - clampf(transmission, 0.0f, 1.0f));
    float final_transmission = clampf(transmission, 0.0f, 1.0f) * (1.0f - clampf(metallic, 0.0f, 1.0f));
    float specular_weight = (1.0f - final_transmission);

    V3 weight = v3(0.0f);
    V3 V = normalize(-rayDirection);
    bool isInside = !h.front;
    N = isInside ? -N : N;
    V3 L = v3(0.0f);
    float NdotV = dot(N, V);

    V3 diffuseColor = diffuse_weight * baseColor;
    V3 specularColor = specular_weight * (baseColor * metallic + (mat.specular * 0.08f * v3(1.0f)) * (1.0f - metallic));
    V3 transmissionColor = transmission * baseColor;

    float diffuseLum = length(diffuseColor);
    float specularLum = length(specularColor);
    float probDiffuse = diffuseLum / (diffuseLum + specularLum);
    if (diffuseLum == 0 && specularLum == 0) {
      if (allEq(baseColor, v3(0.0f))) {
        if (diffuse_weight == 1) probDiffuse = 1.0f;
        else if (diffuse_weight == 0) probDiffuse = 0.0f;
        else probDiffuse = 0.5f;
      } else
        probDiffuse = 0.0f;
    } else {
      probDiffuse *= specular_weight;
    }
    bool chooseDiffuse = rnd(ray.seed) < probDiffuse;

    if (chooseDiffuse) {
      L = cosineHemisphereSampling(ray.seed, N);
      float NdotL = clampf(dot(N, L), 0, 1);
      weight = M_PI_F * diffuseColor * NdotL / probDiffuse;
    }
    if (!chooseDiffuse) {
      if (final_transmission == 0) {
        V3 H = sampleGGX(ray.seed, a2, N);
        float HdotV = dot(H, V);
        L = 2 * HdotV * H - V;
        float NoV = std::fmax(dot(N, V), 1e-7f);
        float NoL = dot(N, L);
        float NoH = std::fmax(dot(N, H), 1e-7f);
        float VoH = std::fmax(dot(V, H), 1e-7f);
        if (NoL >= 0) {
          float G = GeometricShadowing(NoV, NoL, a2);
          V3 F = schlickFresnel(specularColor, VoH);
          weight = M_PI_F * F * G * VoH / (NoH * NoV * (1 - probDiffuse));
        } else
          weight = v3(0.0f);
      } else {
        float ior = isInside ? 1 / f : f;
        float _dot = isInside ? NdotV * ior : NdotV;
        V3 refractedL = refract(-V, N, 1 / f);
        float reflectProb = anyNe(refractedL, v3(0.0f)) ? Schlick(_dot, f) : 1.0f;
        if (rnd(ray.seed) >= reflectProb) {
          ray.refractive = true;
          L = refractedL;
          weight = M_PI_F * transmissionColor / (1 - probDiffuse);
        } else {
          L = reflect(-V, N);
          weight = M_PI_F * transmissionColor / (1 - probDiffuse);
        }
      }
    }

    HitCtx hc{isInside, rayDirection};
    deadPath = allEq(weight, v3(0.0f));
    ray.skipInst = -1;
    if (s.skipOwnInstance && g.convex) {
      const Tri& tr0 = g.tris[h.prim];
      const V3 gn = {tr0.e1.y * tr0.e2.z - tr0.e1.z * tr0.e2.y, tr0.e1.z * tr0.e2.x - tr0.e1.x * tr0.e2.z,
                     tr0.e1.x * tr0.e2.y - tr0.e1.y * tr0.e2.x};
      V3 Ng;
      Ng.x = (gn.x * ir.inv[0][0] + gn.y * ir.inv[1][0]) + gn.z * ir.inv[2][0];
      Ng.y = (gn.x * ir.inv[0][1] + gn.y * ir.inv[1][1]) + gn.z * ir.inv[2][1];
      Ng.z = (gn.x * ir.inv[0][2] + gn.y * ir.inv[1][2]) + gn.z * ir.inv[2][2];
      if (Ng.x != 0.0f || Ng.y != 0.0f || Ng.z != 0.0f) {
        hc.inst = h.inst;
        hc.Ng = Ng;
        if (leavesSurface(Ng, L)) ray.skipInst = h.inst;
      }
    }
    V3 sc = traceDirectionalLight(ray, hc, worldPos, N, f, a2, diffuseColor, specularColor, transmissionColor);
    sc = sc + tracePointLights(ray, hc, worldPos, N, f, a2, diffuseColor, specularColor, transmissionColor);
    sc = sc + traceActiveLights(ray, hc, worldPos, N, f, a2, diffuseColor, specularColor, transmissionColor);
    ray.shadow_color = sc;

    ray.origin = worldPos;
    ray.direction = L;
    ray.emission = v3(0.0f);
    ray.weight = weight;
    ray.albedo = baseColor;
    ray.normal = N;
  }
};

struct FrameOut {
  float* sum;       // w*h*4, un-normalised sample sum (alpha unused)
  float* albedo;    // w*h*4
  float* normal;    // w*h*4
  int32_t* hitIds;  // w*h*2
  float* hitT;      // w*h
  float* depth;     // w*h
};

// PathTrace.rgen:19-141 for one pixel, samples [s0, s1)
static void raygenPixel(Shader& sh, const Camera& cam, uint32_t w, uint32_t h, uint32_t x, uint32_t y,
                        uint32_t s0, uint32_t s1, uint32_t clockBase, const FrameOut& out) {
  const PushConstants& pc = sh.pc;
  uint32_t mapping = y * w + x;
  uint32_t seed = tea(mapping, clockBase);
  for (uint32_t k = 0; k < 2 * s0; k++) lcg(seed);  // samples before s0 consumed two draws each
  V3 colors = v3(0.0f), albedo = v3(0.0f), normal = v3(0.0f);
  int32_t hitInst = -1, hitPrim = -1;
  float hitT = 0.0f, hitDepth = 0.0f;

  for (uint32_t i = s0; i < s1; ++i) {
    RayPayLoad ray{};
    ray.seed = tea(mapping, clockBase + 1u + i);
    float jx = rnd(seed);
    float jy = rnd(seed);
    float px = float(x) + jx, py = float(y) + jy;
    float nx = px / float(w), ny = py / float(h);
    float dx = nx * 2.0f - 1.0f, dy = ny * 2.0f - 1.0f;

    float aperture = cam.position[3];
    float focusDistance = cam.front[3];
    float ox, oy;
    diskSampling(ray.seed, ox, oy);
    ox = aperture / 2.0f * ox;
    oy = aperture / 2.0f * oy;

    float target[4], origin[4], direction[4];
    mulMat4(cam.projInverse, dx, dy, 1.0f, 1.0f, target);
    if (aperture > 0.0f) {
      mulMat4(cam.viewInverse, ox, oy, 0.0f, 1.0f, origin);
      V3 dd = normalize(v3(target[0], target[1], target[2]) * focusDistance - v3(ox, oy, 0.0f));
      mulMat4(cam.viewInverse, dd.x, dd.y, dd.z, 0.0f, direction);
    } else {
      mulMat4(cam.viewInverse, 0.0f, 0.0f, 0.0f, 1.0f, origin);
      V3 dd = normalize(v3(target[0], target[1], target[2]));
      mulMat4(cam.viewInverse, dd.x, dd.y, dd.z, 0.0f, direction);
    }
    ray.direction = v3(direction[0], direction[1], direction[2]);
    ray.origin = v3(origin[0], origin[1], origin[2]);
    ray.weight = v3(0.0f);
    ray.emission = v3(1.0f);
    ray.albedo = v3(0.0f);
    ray.normal = v3(0.0f);
    ray.refractive = false;
    ray.type = 0;
    ray.shadow_color = v3(0.0f);
    ray.skipInst = -1;

    V3 weight = v3(1.0f);
    V3 color = v3(0.0f);
    const float tMin = 0.001f, tMax = 10000.0f;

    for (ray.depth = 0; ray.depth <= pc.maxPathDepth; ++ray.depth) {
      if (ray.depth > 0) ray.type = 2;
      Hit hit;
      sh.extRays++;
      V3 ro = ray.origin, rd = ray.direction;
      bool found = sh.tr.trace(ro, rd, tMin, tMax, /*anyHit=*/true, ray.seed, false, hit, ray.skipInst);
      if (i == 0 && ray.depth == 0) {
        if (found) {
          hitInst = hit.inst;
          hitPrim = hit.prim;
          hitT = hit.t;
          V3 P = ro + rd * hit.t;
          const float* vm = cam.view;
          hitDepth = -(((vm[2] * P.x + vm[6] * P.y) + vm[10] * P.z) + vm[14]);
        }
      }
      uint32_t depthBefore = ray.depth;
      if (found) {
        sh.extHits++;
        sh.closestHit(ray, hit, ro, rd);
      } else {
        sh.miss(ray);
      }
      color += ray.emission * weight;
      weight *= ray.weight;
      color += ray.shadow_color * weight;
      ray.shadow_color = v3(0.0f);
      if (i == 0 && ray.depth == 0) {  // NB: miss / emissive hits set ray.depth first (rgen:114)
        albedo = ray.albedo;
        normal = ray.normal;
      }
      (void)depthBefore;
      if (allEq(weight, v3(0.0f))) break;
      if (pc.russianRoulette && ray.depth >= pc.russianRouletteMinBounces) {
        float p = std::fmax(weight.x, std::fmax(weight.y, weight.z));
        float r = rnd(ray.seed);
        if (r > p) break;
        weight *= 1.0f / p;
      }
    }
    colors += color;
  }
  size_t pi = size_t(y) * w + x;
  if (out.sum) {
    out.sum[4 * pi + 0] = colors.x;
    out.sum[4 * pi + 1] = colors.y;
    out.sum[4 * pi + 2] = colors.z;
    out.sum[4 * pi + 3] = float(s1 - s0);
  }
  if (s0 == 0) {
    if (out.albedo) {
      out.albedo[4 * pi + 0] = albedo.x; out.albedo[4 * pi + 1] = albedo.y;
      out.albedo[4 * pi + 2] = albedo.z; out.albedo[4 * pi + 3] = 1.0f;
    }
    if (out.normal) {
      out.normal[4 * pi + 0] = normal.x; out.normal[4 * pi + 1] = normal.y;
      out.normal[4 * pi + 2] = normal.z; out.normal[4 * pi + 3] = 1.0f;
    }
    if (out.hitIds) { out.hitIds[2 * pi] = hitInst; out.hitIds[2 * pi + 1] = hitPrim; }
    if (out.hitT) out.hitT[pi] = hitT;
    if (out.depth) out.depth[pi] = hitDepth;
  }
}

}  // namespace kfo

// ------------------------------------------------------------------------------------------------
// C entry points (ctypes-friendly)
// ------------------------------------------------------------------------------------------------
using namespace kfo;

extern "C" {

void* kfo_create() { return new Scene(); }
void kfo_destroy(void* h) { delete static_cast<Scene*>(h); }

uint32_t kfo_tea(uint32_t a, uint32_t b) { return tea(a, b); }
uint32_t kfo_lcg(uint32_t* state) { return lcg(*state); }
float kfo_rnd(uint32_t* state) { return rnd(*state); }

int kfo_set_geometry(void* h, uint32_t index, const void* verts, uint32_t nVerts, const uint32_t* idx,
                     uint32_t nIdx, const uint32_t* matIndex, uint32_t nMat, int opaque, int hide) {
  Scene& s = *static_cast<Scene*>(h);
  if (nIdx % 3 != 0 || nMat < nIdx / 3) return 1;
  for (uint32_t i = 0; i < nIdx; i++)
    if (idx[i] >= nVerts) return 1;
  if (s.geoms.size() <= index) s.geoms.resize(index + 1);
  Geometry& g = s.geoms[index];
  g = Geometry();
  g.verts.assign(static_cast<const Vertex*>(verts), static_cast<const Vertex*>(verts) + nVerts);
  g.idx.assign(idx, idx + nIdx);
  g.matIndex.assign(matIndex, matIndex + nMat);
  g.opaque = opaque != 0;
  g.hide = hide != 0;
  g.present = true;
  s.accelDirty = true;
  return 0;
}
int kfo_set_materials(void* h, const void* mats, uint32_t n) {
  Scene& s = *static_cast<Scene*>(h);
  s.mats.assign(static_cast<const Material*>(mats), static_cast<const Material*>(mats) + n);
  return 0;
}
int kfo_set_texture(void* h, uint32_t index, const uint8_t* rgba, uint32_t w, uint32_t ht) {
  Scene& s = *static_cast<Scene*>(h);
  if (s.texs.size() <= index) s.texs.resize(index + 1);
  s.texs[index].w = w;
  s.texs[index].h = ht;
  s.texs[index].rgba.assign(rgba, rgba + size_t(w) * ht * 4);
  return 0;
}
int kfo_set_env_cube(void* h, const uint8_t* const* faces, uint32_t size) {
  Scene& s = *static_cast<Scene*>(h);
  for (int f = 0; f < 6; f++) {
    s.envFaces[f].w = s.envFaces[f].h = size;
    s.envFaces[f].rgba.assign(faces[f], faces[f] + size_t(size) * size * 4);
  }
  s.hasEnv = true;
  return 0;
}
int kfo_set_instances(void* h, const void* insts, uint32_t n) {
  Scene& s = *static_cast<Scene*>(h);
  s.insts.assign(static_cast<const Instance*>(insts), static_cast<const Instance*>(insts) + n);
  s.accelDirty = true;
  return 0;
}
int kfo_set_transforms(void* h, const float* transforms, uint32_t n) {
  Scene& s = *static_cast<Scene*>(h);
  if (n != s.insts.size()) return 1;
  for (uint32_t i = 0; i < n; i++) std::memcpy(s.insts[i].transform, transforms + 16 * i, 64);
  s.accelDirty = true;
  return 0;
}
// D6 mirror (see Scene::skipOwnInstance); off by default.
int kfo_set_skip_own_instance(void* h, int on) {
  static_cast<Scene*>(h)->skipOwnInstance = on != 0;
  return 0;
}
// Mirror of the CUDA path's light-sample culling (see Scene::cullLightSamples); off by default.
int kfo_set_cull_light_samples(void* h, int on) {
  static_cast<Scene*>(h)->cullLightSamples = on != 0;
  return 0;
}
unsigned long long kfo_last_shadow_skipped(void* h) { return static_cast<Scene*>(h)->lastShadowSkipped; }
int kfo_set_lights(void* h, const void* dl, const void* pl, const void* al) {
  Scene& s = *static_cast<Scene*>(h);
  if (dl) std::memcpy(&s.dl, dl, sizeof(DirLight)); else std::memset(&s.dl, 0, sizeof(DirLight));
  if (pl) std::memcpy(&s.pl, pl, sizeof(PointLights)); else std::memset(&s.pl, 0, sizeof(PointLights));
  if (al) std::memcpy(&s.al, al, sizeof(ActiveLights)); else std::memset(&s.al, 0, sizeof(ActiveLights));
  return 0;
}

// Renders samples [s0, s1) of every pixel of nCams cameras.  Outputs are camera-major; any may be
// NULL.  mode: 0 = BVH traversal, 1 = brute force over all triangles.  threads <= 0: all cores.
// counters (may be NULL): [paths, extensionRays, shadowRays, extensionHits].
int kfo_render(void* h, const void* cams, uint32_t nCams, uint32_t w, uint32_t ht, const void* pcIn,
               uint32_t s0, uint32_t s1, uint32_t clockBase, int mode, int threads, float* sum,
               float* albedo, float* normal, int32_t* hitIds, float* hitT, float* depth,
               uint64_t* counters) {
  Scene& s = *static_cast<Scene*>(h);
  if (s.accelDirty) buildAccel(s);
  PushConstants pc;
  std::memcpy(&pc, pcIn, sizeof(pc));
  for (const Instance& in : s.insts) {
    if (in.geometryIndex >= s.geoms.size() || !s.geoms[in.geometryIndex].present) return 1;
  }
  for (const Geometry& g : s.geoms)
    for (size_t p = 0; p < g.idx.size() / 3; p++)
      if (g.matIndex[p] >= s.mats.size()) return 2;
  int nt = threads > 0 ? threads : int(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  std::atomic<uint64_t> nextRow{0};
  std::atomic<uint64_t> cExt{0}, cSh{0}, cHit{0}, cSkip{0};
  uint64_t totalRows = uint64_t(nCams) * ht;
  auto worker = [&]() {
    Shader sh(s, pc, mode == 1);
    for (;;) {
      uint64_t row = nextRow.fetch_add(1);
      if (row >= totalRows) break;
      uint32_t c = uint32_t(row / ht), y = uint32_t(row % ht);
      size_t off = size_t(c) * w * ht;
      FrameOut o{sum ? sum + 4 * off : nullptr,       albedo ? albedo + 4 * off : nullptr,
                 normal ? normal + 4 * off : nullptr, hitIds ? hitIds + 2 * off : nullptr,
                 hitT ? hitT + off : nullptr,         depth ? depth + off : nullptr};
      const Camera& cam = static_cast<const Camera*>(cams)[c];
      for (uint32_t x = 0; x < w; x++) raygenPixel(sh, cam, w, ht, x, y, s0, s1, clockBase, o);
    }
    cExt += sh.extRays;
    cSh += sh.shRays;
    cHit += sh.extHits;
    cSkip += sh.shSkipped;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  s.lastShadowSkipped = cSkip;
  if (counters) {
    counters[0] = uint64_t(nCams) * w * ht * (s1 - s0);
    counters[1] = cExt;
    counters[2] = cSh;
    counters[3] = cHit;
  }
  return 0;
}

// PathTrace.rgen:143-163 (accumulate) followed by PostProcessing.frag:11-18 into a B8G8R8A8Srgb
// attachment (include/core/config.hpp:144): rgba <- running mean, bgra8 <- encode(rgba).
int kfo_resolve(void* h, const float* sum, float* rgba, uint8_t* bgra8, uint64_t nPixels,
                uint32_t spp, int32_t frameCount) {
  Scene& s = *static_cast<Scene*>(h);
  auto encode = [&](float c) -> uint8_t {
    if (!(c > 0.0f)) return 0;  // clamp (NaN -> 0 like a UNORM conversion)
    int lo = 0, hi = 255;       // number of thresholds <= c
    while (lo < hi) {
      int mid = (lo + hi) / 2;
      if (c >= s.srgbThreshold[mid]) lo = mid + 1; else hi = mid;
    }
    return uint8_t(lo);
  };
  for (uint64_t i = 0; i < nPixels; i++) {
    float fc[3];
    for (int k = 0; k < 3; k++) fc[k] = sum[4 * i + k] / float(spp);
    if (frameCount > 0) {
      float a = 1.0f / float(frameCount + 1);
      for (int k = 0; k < 3; k++) fc[k] = rgba[4 * i + k] * (1.0f - a) + fc[k] * a;  // mix()
    }
    for (int k = 0; k < 3; k++) rgba[4 * i + k] = fc[k];
    rgba[4 * i + 3] = 1.0f;
    if (bgra8) {
      bgra8[4 * i + 0] = encode(fc[2]);
      bgra8[4 * i + 1] = encode(fc[1]);
      bgra8[4 * i + 2] = encode(fc[0]);
      bgra8[4 * i + 3] = 255;
    }
  }
  return 0;
}

// ---- black boxes handed to oracle/ref_shim (the reference's shaders compiled for the CPU) ----------
// The shim runs the reference's GLSL; what the GLSL delegates to the driver and the hardware --
// traceRayEXT's traversal + triangle test, texture()'s sampler -- it takes from here, so that oracle
// and shim share exactly these definitions and differ only in who wrote the shading code.
struct KfoHitOut {
  float t, u, v;
  int32_t inst, prim, front;
  float worldToObject[12];  // 3 rows x 4 columns
};
struct KfoSceneView {
  uint32_t nGeoms, nMats, nInsts, nTex;
  const void* const* verts;     // per geometry: 48-byte vertices
  const void* const* idx;       // per geometry: uint32 indices
  const void* const* matIndex;  // per geometry: uint32 material index per primitive
  const void* mats;
  const void* insts;
  const void* dl;
  const void* pl;
  const void* al;
  int32_t hasEnv;
};

int kfo_prepare(void* h) {
  Scene& s = *static_cast<Scene*>(h);
  if (s.accelDirty) buildAccel(s);
  return 0;
}

int kfo_scene_view(void* h, KfoSceneView* v) {
  Scene& s = *static_cast<Scene*>(h);
  s.viewVerts.clear(); s.viewIdx.clear(); s.viewMat.clear();
  for (const Geometry& g : s.geoms) {
    s.viewVerts.push_back(g.verts.data());
    s.viewIdx.push_back(g.idx.data());
    s.viewMat.push_back(g.matIndex.data());
  }
  v->nGeoms = uint32_t(s.geoms.size());
  v->nMats = uint32_t(s.mats.size());
  v->nInsts = uint32_t(s.insts.size());
  v->nTex = uint32_t(s.texs.size());
  v->verts = s.viewVerts.data();
  v->idx = s.viewIdx.data();
  v->matIndex = s.viewMat.data();
  v->mats = s.mats.data();
  v->insts = s.insts.data();
  v->dl = &s.dl;
  v->pl = &s.pl;
  v->al = &s.al;
  v->hasEnv = s.hasEnv ? 1 : 0;
  return 0;
}

// traceRayEXT's traversal: closest (or first, when terminateOnFirst) accepted hit with tmin < t < tmax.
// anyHit != 0: candidates of non-opaque geometry go to `fn` (gl_RayFlagsOpaqueEXT clear).
int kfo_trace(void* h, const float* o, const float* d, float tmin, float tmax, int anyHit, int terminateOnFirst,
              int brute, AnyHitFn fn, void* user, KfoHitOut* out) {
  const Scene& s = *static_cast<const Scene*>(h);
  Tracer tr(s, brute != 0);
  tr.anyHitFn = fn;
  tr.anyHitUser = user;
  Hit hit;
  bool found = tr.trace(v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), tmin, tmax, anyHit != 0, 0u,
                        terminateOnFirst != 0, hit);
  if (!found) return 0;
  out->t = hit.t; out->u = hit.u; out->v = hit.v;
  out->inst = hit.inst; out->prim = hit.prim; out->front = hit.front ? 1 : 0;
  std::memcpy(out->worldToObject, s.instRt[hit.inst].inv, sizeof(out->worldToObject));
  return 1;
}

void kfo_sample_texture(void* h, int32_t index, float u, float v, float* rgb) {
  V3 c = sampleTexture(*static_cast<const Scene*>(h), index, u, v);
  rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}

void kfo_sample_cube(void* h, const float* dir, float* rgb) {
  V3 c = sampleCube(*static_cast<const Scene*>(h), v3(dir[0], dir[1], dir[2]));
  rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}

int kfo_hardware_threads() { return int(std::thread::hardware_concurrency()); }

}  // extern "C"
