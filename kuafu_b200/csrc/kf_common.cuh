// kf_common.cuh -- shared device/host definitions of the kfrt CUDA core (sm_100a).
//
// Arithmetic contract (DESIGN.md "bit-exact hits"): everything that decides WHICH triangle a ray
// hits and at what t -- camera ray generation, the world->object ray transform, the
// Moller-Trumbore test, the instance inverse -- is written with the c*() helpers below, which map to
// __fmul_rn/__fadd_rn/__fsub_rn and therefore are never contracted into FMA, whatever -fmad says.
// Box tests and shading are free to use FMA: boxes are padded conservatively and radiance parity
// is toleranced.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kf_rt.h"

#define KF_HD __host__ __device__ __forceinline__
#define KF_D __device__ __forceinline__

namespace kf {

// ---------------------------------------------------------------------------------------------
// contract arithmetic (no FMA contraction, IEEE round-to-nearest)
// ---------------------------------------------------------------------------------------------
KF_D float cmul(float a, float b) { return __fmul_rn(a, b); }
KF_D float cadd(float a, float b) { return __fadd_rn(a, b); }
KF_D float csub(float a, float b) { return __fsub_rn(a, b); }
KF_D float cdiv(float a, float b) { return __fdiv_rn(a, b); }
KF_D float csqrt(float a) { return __fsqrt_rn(a); }
KF_D float cdot3(float ax, float ay, float az, float bx, float by, float bz) {
  return cadd(cadd(cmul(ax, bx), cmul(ay, by)), cmul(az, bz));
}

struct V3 {
  float x, y, z;
};
KF_HD V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
KF_HD V3 mk3(float a) { return V3{a, a, a}; }
KF_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
KF_HD V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
KF_HD V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
KF_HD V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
KF_HD V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
KF_HD V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
// Shading math only (radiance is compared with a tolerance, never bit for bit): one correctly rounded
// reciprocal and three multiplies instead of three IEEE divisions (each a ~12-instruction sequence;
// ncu showed this operator as 8.6 % of the shade kernel's instructions).  Within one ulp of a / s:
// the parity margins of tests/parity_margin.py do not move (0.02 - 0.04 % of pixels beyond 1e-3).
// Going further -- rsqrt in normalize(), reciprocal quotients in the microfacet terms, SFU sine /
// cosine -- buys another 7 % of the shade stage but puts 0.7 % of the pixels of the mirror-heavy
// scenes beyond 1e-3 (limit 2 %), so it is not done.
KF_HD V3 operator/(V3 a, float s) {
#ifdef __CUDA_ARCH__
  const float r = __frcp_rn(s);
#else
  const float r = 1.0f / s;
#endif
  return {a.x * r, a.y * r, a.z * r};
}
KF_HD V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
KF_HD V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
KF_HD V3& operator*=(V3& a, float s) { a = a * s; return a; }
KF_HD bool allEq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
KF_HD bool anyNe(V3 a, V3 b) { return a.x != b.x || a.y != b.y || a.z != b.z; }
KF_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
KF_HD V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
KF_D float length(V3 a) { return sqrtf(dot(a, a)); }
KF_D V3 normalize(V3 a) {
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
// contract versions (used where bit-exactness with the oracle is required)
KF_D V3 cnormalize(V3 a) {
  float inv = cdiv(1.0f, csqrt(cdot3(a.x, a.y, a.z, a.x, a.y, a.z)));
  return {cmul(a.x, inv), cmul(a.y, inv), cmul(a.z, inv)};
}
KF_D V3 ccross(V3 a, V3 b) {
  return {csub(cmul(a.y, b.z), cmul(a.z, b.y)), csub(cmul(a.z, b.x), cmul(a.x, b.z)),
          csub(cmul(a.x, b.y), cmul(a.y, b.x))};
}
KF_D float cdot(V3 a, V3 b) { return cdot3(a.x, a.y, a.z, b.x, b.y, b.z); }
KF_D V3 csub3(V3 a, V3 b) { return {csub(a.x, b.x), csub(a.y, b.y), csub(a.z, b.z)}; }
// Fused arithmetic of the traversal black box (world -> object transform, triangle test): explicit
// single-rounding FMAs, the same expressions as fdot() / fcross() in oracle/kf_oracle.cpp.
KF_D float cfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
KF_D float fdot(V3 a, V3 b) { return cfma(a.z, b.z, cfma(a.y, b.y, cmul(a.x, b.x))); }
KF_D V3 fcross(V3 a, V3 b) {
  return {cfma(a.y, b.z, -cmul(a.z, b.y)), cfma(a.z, b.x, -cmul(a.x, b.z)), cfma(a.x, b.y, -cmul(a.y, b.x))};
}
// column-major 4x4 times (x,y,z,w), summed left to right (oracle mulMat4)
KF_D void cmulMat4(const float* m, float x, float y, float z, float w, float out[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++)
    out[r] = cadd(cadd(cadd(cmul(m[r], x), cmul(m[4 + r], y)), cmul(m[8 + r], z)), cmul(m[12 + r], w));
}

KF_D float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

#define KF_PI 3.141592f /* reference base/Random.glsl:1 */

// ---------------------------------------------------------------------------------------------
// RNG (reference base/Random.glsl:7-39)
// ---------------------------------------------------------------------------------------------
KF_HD uint32_t tea(uint32_t val0, uint32_t val1) {
  uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
  for (int n = 0; n < 16; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
KF_HD uint32_t lcg(uint32_t& prev) {
  prev = 1664525u * prev + 1013904223u;
  return prev & 0x00FFFFFFu;
}
// The LCG state after n more steps, in O(log n): x -> a^n x + c (a^n - 1) / (a - 1) mod 2^32, by squaring
// the affine map (the state after n calls of lcg(), bit for bit).
KF_HD uint32_t lcgSkip(uint32_t state, uint32_t n) {
  uint32_t mul = 1664525u, add = 1013904223u;  // the map of 2^k steps
  uint32_t accMul = 1u, accAdd = 0u;           // the map of the steps taken so far
  while (n) {
    if (n & 1u) {
      accMul *= mul;
      accAdd = accAdd * mul + add;
    }
    add = (mul + 1u) * add;
    mul *= mul;
    n >>= 1;
  }
  return accMul * state + accAdd;
}
// float(lcg)/float(2^24): the 24-bit integer converts exactly and the division by a power of two is
// exact, so a multiply by 2^-24 is bit-identical to the reference's division.
KF_HD float rnd(uint32_t& prev) { return float(lcg(prev)) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------------------------------------
// Pieces shared by the stage kernels of the wavefront scheduler
// ---------------------------------------------------------------------------------------------
// Warp-aggregated append: one atomic per warp, lanes get consecutive positions.
KF_D void queueAppend(uint32_t* __restrict__ queue, uint32_t* __restrict__ count, bool pred, uint32_t value) {
  const uint32_t mask = __ballot_sync(0xffffffffu, pred);
  if (mask == 0u) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(count, uint32_t(__popc(mask)));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

// Block-aggregated append (blocks of up to 8 warps, every thread of the block calls it the same number of
// times): one atomic per block, and the block's entries stay together in thread order -- runs four times
// as long as the per-warp append leaves, for the stage that reads the queue next.  `round` alternates the
// staging slots so that one barrier pair per call is enough.
KF_D void queueAppendBlock(uint32_t* __restrict__ queue, uint32_t* __restrict__ count, bool pred, uint32_t value,
                           uint32_t (*sCnt)[8], uint32_t* sBase, uint32_t round) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nWarps = (blockDim.x + 31u) >> 5;
  const uint32_t mask = __ballot_sync(0xffffffffu, pred);
  if (lane == 0) sCnt[round][warp] = __popc(mask);
  __syncthreads();
  uint32_t before = 0, total = 0;
  for (uint32_t w = 0; w < nWarps; w++) {
    const uint32_t c = sCnt[round][w];
    if (w < warp) before += c;
    total += c;
  }
  if (threadIdx.x == 0 && total) sBase[round] = atomicAdd(count, total);
  __syncthreads();
  if (pred) queue[sBase[round] + before + __popc(mask & ((1u << lane) - 1u))] = value;
}

// Russian roulette and hand-over to the next bounce (reference PathTrace.rgen:119-138).
// Returns true when the path continues.
KF_D bool advancePath(const KfrtPushConstants& pc, uint32_t depth, V3& weight, uint32_t& seed) {
  if (allEq(weight, mk3(0.0f))) return false;
  if (pc.russianRoulette && depth >= pc.russianRouletteMinBounces) {
    const float p = fmaxf(weight.x, fmaxf(weight.y, weight.z));
    const float r = rnd(seed);
    if (r > p) return false;
    weight *= 1.0f / p;
  }
  return depth < pc.maxPathDepth;
}

// ---------------------------------------------------------------------------------------------
// Acceleration-structure records
// ---------------------------------------------------------------------------------------------
// 8-wide compressed node, 80 bytes = 5 x 16 B loads.  Child boxes are 8-bit offsets from `p` on a
// per-axis power-of-two grid 2^(e-127) (stored: e + 15); children sit in octant-ordered slots so that traversal
// order is a bit trick instead of a sort (after Ylitie, Karras, Laine 2017).
struct __align__(16) Node8 {
  float px, py, pz;
  uint8_t ex, ey, ez, imask;  // biased grid exponents + 15 (see intersectNodeWords); imask bit i: child slot i is an internal node
  uint32_t childBase;         // index of first internal child (children are consecutive by slot)
  uint32_t primBase;          // index of first leaf primitive (triangle / instance-list entry)
  // Leaf children: slot s owns the two-bit field 2s..2s+1 of a 16-bit triangle mask (unary count of
  // its triangles, at most 2 per leaf); the triangles of a node lie compactly from primBase in slot
  // order, so the triangle of mask bit b is primBase + popc(mask below b).  The mask is stored in
  // both halves of triMask (low half: ANDed with the expanded hit mask, high half: kept as is).
  uint32_t triMask;
  uint32_t reserved;
  uint8_t qlox[8], qloy[8], qloz[8];
  uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

// Stored triangle, 48 bytes = 3 x 16 B loads: v0 + primitive id, e1 = v1 - v0, e2 = v2 - v0.
struct __align__(16) Tri48 {
  float v0x, v0y, v0z;
  uint32_t prim;
  float e1x, e1y, e1z, pad1;
  float e2x, e2y, e2z, pad2;
};
static_assert(sizeof(Tri48) == 48, "Tri48 must be 48 bytes");

// Per-instance shading record, indexed by instance: world->object (for the normal transform) and the
// geometry's attribute tables, so that the attribute gather of a hit is instance -> indices ->
// vertices with no detour over the instance SSBO and the geometry table.  80 bytes = 5 x 16 B loads.
// What shading a hit needs of its triangle, gathered once per BLAS build in primitive order (64 bytes
// = 4 x 16 B loads): the three vertex normals, the three texture coordinates and the material index.
// It replaces the reference's chain indices -> three 48-byte vertices -> matIndices[prim]
// (PathTrace.rchit:66-98: 12 + 144 + 4 scattered bytes behind one more dependent load) with the same
// bits read from one place.
struct __align__(16) ShadeTri {
  float n0[3], n1[3], n2[3];
  float uv0[2], uv1[2], uv2[2];
  uint32_t matIndex;
};
static_assert(sizeof(ShadeTri) == 64, "ShadeTri must be 64 bytes");

struct __align__(16) InstRec {
  float inv[12];            // world->object, 3 rows x 4 columns
  const KfrtVertex* verts;  // geometry tables (reference PathTrace.rchit:66-98)
  const uint32_t* idx;
  const float4* faceNormal;  // convex geometries only, else NULL: object-space geometric normal e1 x e2 of every primitive (see kf_blas_batch.cuh, k_batch_convex)
  const ShadeTri* shade;    // per-primitive shading records of the geometry
};
static_assert(sizeof(InstRec) == 80, "InstRec must be 80 bytes");

// Instance record as it sits in the top-level node array (same 80-byte stride as a Node8, addressed
// like an internal child): what a ray needs to enter a bottom-level structure, 5 x 16 B loads.
struct __align__(16) InstNode {
  float inv[12];       // world->object, 3 rows x 4 columns
  const Node8* nodes;  // BLAS nodes; NULL when the geometry is hidden/empty
  const Tri48* tris;   // BLAS triangles in leaf order; bit 0 set == non-opaque geometry
  uint16_t box[6];     // world box as fp16 bits, rounded outward: lo.xyz, hi.xyz (re-test against the
                       // current closest hit before the ray is transformed)
  uint32_t instIndex;  // instance index (gl_InstanceID)
};
static_assert(sizeof(InstNode) == 80, "InstNode must be 80 bytes");

// Per-geometry shading tables (fetched only at shading time).
struct GeomRec {
  const KfrtVertex* verts;
  const uint32_t* idx;
  const uint32_t* matIndex;
  uint32_t nTris;
  uint32_t flags;  // bit0 opaque, bit1 hideRender
};

struct TexRec {
  const uchar4* texels;
  uint32_t w, h;
};

struct SceneDev {
  const Node8* tlasNodes;  // top level: Node8 and InstNode records in one array (kf_bvh_build.cuh)
  const Node8* wsiNodes;   // top level + world-space instance subtrees in one array (kf_wsi.cuh); the traversal
  const Tri48* wsiTris;    // stages walk it instead of the two-level structure while it is valid
  const InstRec* inst;
  const KfrtInstance* instSsbo;
  const GeomRec* geoms;
  const KfrtMaterial* mats;
  const TexRec* texs;
  uint32_t nTex;
  uint32_t nInst;
  const uchar4* envFaces;  // 6 faces, size x size each
  uint32_t envSize;
  const float* srgbToLinear;  // 256 entries
  const KfrtDirectionalLight* dl;
  const KfrtPointLights* pl;
  const KfrtActiveLights* al;
  const float* alProjView;  // 16 floats per projector slot: proj * view, column major
  unsigned long long lightMask;  // bit k: light slot k (0 directional, 1..32 point, 33..40 active) is on
};

struct Hit {
  float t, u, v;
  int32_t inst, prim;  // -1 on miss
  uint32_t front;      // 1: front face (det > 0)
};

}  // namespace kf
