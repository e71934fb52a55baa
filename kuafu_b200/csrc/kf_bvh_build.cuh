// kf_bvh_build.cuh -- device construction of the acceleration structures that replace
// VkAccelerationStructureKHR BLAS/TLAS (reference src/core/rt/rt.cpp:142-370 build + compaction,
// :372-495 per-frame TLAS update).
//
// Here: the pieces both levels share (Morton keys, LSD radix sort, Karras hierarchy, bottom-up boxes,
// collapse of a binary hierarchy into quantised 8-wide nodes) and everything of the top level (exact
// instance boxes, binned-SAH hierarchy in one launch, collapse in one launch, instance records, refit).
// The bottom level is built for all geometries at once by kf_blas_batch.cuh on top of these pieces.
// Refit keeps the binary topology and the slot assignment, recomputes boxes bottom-up and
// re-quantises every wide node.
#pragma once

#include <cuda_fp16.h>

#include "kf_common.cuh"

namespace kf {

// -------------------------------------------------------------------------------------------------
// helpers
// -------------------------------------------------------------------------------------------------
struct Box6 {
  float lo[3], hi[3];
};
KF_D void boxReset(Box6& b) {
  b.lo[0] = b.lo[1] = b.lo[2] = 3.0e38f;
  b.hi[0] = b.hi[1] = b.hi[2] = -3.0e38f;
}
KF_D void boxGrow(Box6& b, const Box6& o) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    b.lo[k] = fminf(b.lo[k], o.lo[k]);
    b.hi[k] = fmaxf(b.hi[k], o.hi[k]);
  }
}
KF_D float boxArea(const Box6& b) {
  float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return dx * dy + dy * dz + dz * dx;
}
// Conservative margin: culling must never reject a primitive the float triangle test accepts.
KF_D void boxPad(Box6& b) {
  float m = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; k++) m = fmaxf(m, fmaxf(fabsf(b.lo[k]), fabsf(b.hi[k])));
  float pad = m * (1.0f / 16384.0f) + 1e-30f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    b.lo[k] -= pad;
    b.hi[k] += pad;
  }
}
KF_D Box6 loadBox(const float* p) {
  Box6 b;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    b.lo[k] = p[k];
    b.hi[k] = p[3 + k];
  }
  return b;
}
KF_D Box6 loadBoxCG(const float* p) {
  Box6 b;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    b.lo[k] = __ldcg(p + k);
    b.hi[k] = __ldcg(p + 3 + k);
  }
  return b;
}
KF_D void storeBox(float* p, const Box6& b) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    p[k] = b.lo[k];
    p[3 + k] = b.hi[k];
  }
}
// float <-> order-preserving int for atomicMin/atomicMax
KF_D int floatToOrdered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
KF_D float orderedToFloat(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// -------------------------------------------------------------------------------------------------
// primitive setup
// -------------------------------------------------------------------------------------------------
__global__ void k_init_scene_box(int* sceneBox) {
  if (threadIdx.x < 3) sceneBox[threadIdx.x] = floatToOrdered(3.0e38f);
  else if (threadIdx.x < 6) sceneBox[threadIdx.x] = floatToOrdered(-3.0e38f);
}

// World->object matrix and InstRec of every instance (k_instance_setup: one thread per instance) and
// its world box (k_instance_box: one block per instance, exact over the transformed vertices -- the
// box of the 8 transformed corners of the BLAS root box is up to 5x larger in volume for a rotated
// round object, and every false box hit costs a ray transform plus a BLAS root visit).
// The inverse follows the float cofactor contract of oracle affineInverse().
struct BlasInfo {
  const Node8* nodes;
  const Tri48* tris;
  const KfrtVertex* verts;
  const uint32_t* idx;
  const uint32_t* matIndex;
  const ShadeTri* shade;
  const float4* faceNormal;  // object-space e1 x e2 per primitive
  float box[6];
  uint32_t flags;  // bit0: usable (has triangles, not hidden); bit1: non-opaque; bit2: convex (k_batch_convex)
  uint32_t nVerts;
};
__global__ void k_instance_setup(const KfrtInstance* __restrict__ insts, uint32_t n,
                                 const BlasInfo* __restrict__ blas, uint32_t nBlas,
                                 InstRec* __restrict__ recs, const float* __restrict__ primBox,
                                 const uint32_t* __restrict__ slotOfInst, Node8* __restrict__ tlasNodes) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* m = insts[i].transform;
  const float a00 = m[0], a10 = m[1], a20 = m[2];
  const float a01 = m[4], a11 = m[5], a21 = m[6];
  const float a02 = m[8], a12 = m[9], a22 = m[10];
  const float t0 = m[12], t1 = m[13], t2 = m[14];
  const float c00 = csub(cmul(a11, a22), cmul(a12, a21));
  const float c01 = csub(cmul(a12, a20), cmul(a10, a22));
  const float c02 = csub(cmul(a10, a21), cmul(a11, a20));
  const float det = cadd(cadd(cmul(a00, c00), cmul(a01, c01)), cmul(a02, c02));
  const float id = cdiv(1.0f, det);
  float inv[3][4];
  inv[0][0] = cmul(c00, id);
  inv[1][0] = cmul(c01, id);
  inv[2][0] = cmul(c02, id);
  inv[0][1] = cmul(csub(cmul(a02, a21), cmul(a01, a22)), id);
  inv[1][1] = cmul(csub(cmul(a00, a22), cmul(a02, a20)), id);
  inv[2][1] = cmul(csub(cmul(a01, a20), cmul(a00, a21)), id);
  inv[0][2] = cmul(csub(cmul(a01, a12), cmul(a02, a11)), id);
  inv[1][2] = cmul(csub(cmul(a02, a10), cmul(a00, a12)), id);
  inv[2][2] = cmul(csub(cmul(a00, a11), cmul(a01, a10)), id);
#pragma unroll
  for (int r = 0; r < 3; r++)
    inv[r][3] = -cadd(cadd(cmul(inv[r][0], t0), cmul(inv[r][1], t1)), cmul(inv[r][2], t2));
  InstRec rec;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) rec.inv[4 * r + c] = inv[r][c];
  const uint32_t g = insts[i].geometryIndex;
  const bool usable = g < nBlas && (blas[g].flags & 1u);
  rec.verts = g < nBlas ? blas[g].verts : nullptr;
  rec.idx = g < nBlas ? blas[g].idx : nullptr;
  rec.faceNormal = (g < nBlas && (blas[g].flags & 4u)) ? blas[g].faceNormal : nullptr;
  rec.shade = g < nBlas ? blas[g].shade : nullptr;
  recs[i] = rec;
  // the record the traversal reads, in the top-level node array
  InstNode in;
#pragma unroll
  for (int k = 0; k < 12; k++) in.inv[k] = rec.inv[k];
  in.nodes = usable ? blas[g].nodes : nullptr;
  in.tris = usable ? reinterpret_cast<const Tri48*>(reinterpret_cast<uintptr_t>(blas[g].tris) |
                                                    ((blas[g].flags & 2u) ? 1ull : 0ull))
                   : nullptr;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    in.box[k] = __half_as_ushort(__float2half_rd(primBox[6 * i + k]));
    in.box[3 + k] = __half_as_ushort(__float2half_ru(primBox[6 * i + 3 + k]));
  }
  in.instIndex = i;
  reinterpret_cast<InstNode*>(tlasNodes)[slotOfInst[i]] = in;
}

// k_instance_box: block (i, c) reduces vertices [c, c + 1) * KF_BOX_CHUNK of instance i and merges its
// partial box into ibox (ordered ints, reset by k_instance_box_init); k_instance_box_finish pads the
// boxes and merges them into the scene box.  min / max are exact, so the result does not depend on
// how the vertices are split.  (The first version gave every instance a single block: 0.5 ms for a
// 250 k-vertex mesh, most of a top-level refit.)
#define KF_BOX_CHUNK 4096
__global__ void k_instance_box_init(int* __restrict__ ibox, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * n) ibox[i] = floatToOrdered((i % 6) < 3 ? 3.0e38f : -3.0e38f);
}
__global__ void __launch_bounds__(128) k_instance_box(const KfrtInstance* __restrict__ insts, uint32_t n,
                                                      const BlasInfo* __restrict__ blas, uint32_t nBlas,
                                                      int* __restrict__ ibox) {
  const uint32_t i = blockIdx.x;
  if (i >= n) return;
  const float* m = insts[i].transform;
  const uint32_t g = insts[i].geometryIndex;
  if (!(g < nBlas && (blas[g].flags & 1u))) return;
  const uint32_t nv = blas[g].nVerts;
  const uint32_t v0 = blockIdx.y * KF_BOX_CHUNK, v1 = min(nv, v0 + KF_BOX_CHUNK);
  if (v0 >= nv) return;
  Box6 wb;
  boxReset(wb);
  {
    const float m00 = m[0], m10 = m[1], m20 = m[2], m01 = m[4], m11 = m[5], m21 = m[6];
    const float m02 = m[8], m12 = m[9], m22 = m[10], t0 = m[12], t1 = m[13], t2 = m[14];
    const KfrtVertex* __restrict__ verts = blas[g].verts;
    for (uint32_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
      const float x = verts[v].pos[0], y = verts[v].pos[1], z = verts[v].pos[2];
      const float px = ((m00 * x + m01 * y) + m02 * z) + t0;
      const float py = ((m10 * x + m11 * y) + m12 * z) + t1;
      const float pz = ((m20 * x + m21 * y) + m22 * z) + t2;
      wb.lo[0] = fminf(wb.lo[0], px); wb.hi[0] = fmaxf(wb.hi[0], px);
      wb.lo[1] = fminf(wb.lo[1], py); wb.hi[1] = fmaxf(wb.hi[1], py);
      wb.lo[2] = fminf(wb.lo[2], pz); wb.hi[2] = fmaxf(wb.hi[2], pz);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wb.lo[k] = fminf(wb.lo[k], __shfl_xor_sync(0xffffffffu, wb.lo[k], o));
      wb.hi[k] = fmaxf(wb.hi[k], __shfl_xor_sync(0xffffffffu, wb.hi[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0 && wb.lo[0] <= wb.hi[0]) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(ibox + 6 * i + k, floatToOrdered(wb.lo[k]));
      atomicMax(ibox + 6 * i + 3 + k, floatToOrdered(wb.hi[k]));
    }
  }
}
__global__ void k_instance_box_finish(const int* __restrict__ ibox, uint32_t n, float* __restrict__ primBox,
                                      int* __restrict__ sceneBox) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Box6 wb;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    wb.lo[k] = orderedToFloat(ibox[6 * i + k]);
    wb.hi[k] = orderedToFloat(ibox[6 * i + 3 + k]);
  }
  if (wb.lo[0] <= wb.hi[0]) {
    boxPad(wb);
  } else {
    // hidden / empty geometry: a harmless degenerate box at the origin (its InstRec has no nodes)
#pragma unroll
    for (int k = 0; k < 3; k++) wb.lo[k] = wb.hi[k] = 0.0f;
  }
  storeBox(primBox + 6 * i, wb);
  if (sceneBox) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(sceneBox + k, floatToOrdered(wb.lo[k]));
      atomicMax(sceneBox + 3 + k, floatToOrdered(wb.hi[k]));
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Morton keys
// -------------------------------------------------------------------------------------------------
KF_D uint64_t expandBits21(uint64_t v) {
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
__global__ void k_morton(const float* __restrict__ primBox, uint32_t n, const int* __restrict__ sceneBox,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t code = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float lo = orderedToFloat(sceneBox[k]), hi = orderedToFloat(sceneBox[3 + k]);
    const float c = 0.5f * (primBox[6 * i + k] + primBox[6 * i + 3 + k]);
    const float ext = hi - lo;
    float f = ext > 0.0f ? (c - lo) / ext : 0.0f;
    f = fminf(fmaxf(f * 2097152.0f, 0.0f), 2097151.0f);
    code |= expandBits21(uint64_t(f)) << (2 - k);
  }
  keys[i] = code;
  vals[i] = i;
}

// -------------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, stable.  Three kernels per pass: per-block digit histograms,
// a single-block exclusive scan in (digit, block) order, and a stable scatter that ranks keys
// warp by warp with __match_any_sync.
// -------------------------------------------------------------------------------------------------
#define KF_SORT_THREADS 256
#define KF_SORT_ITEMS 8
#define KF_SORT_TILE (KF_SORT_THREADS * KF_SORT_ITEMS)

__global__ void k_sort_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                            uint32_t* __restrict__ hist /* [256][numBlocks] */, uint32_t numBlocks) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * KF_SORT_TILE;
#pragma unroll
  for (int it = 0; it < KF_SORT_ITEMS; it++) {
    uint32_t i = base + it * KF_SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * numBlocks + blockIdx.x] = sh[threadIdx.x];
}

// Exclusive scan over hist[256 * numBlocks] by one block (sizes here are small: n / 2048 blocks).
__global__ void k_sort_scan(uint32_t* __restrict__ hist, uint32_t total) {
  __shared__ uint32_t warpSums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < total; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < total ? hist[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warpSums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < (blockDim.x >> 5) ? warpSums[lane] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warpSums[lane] = w;
    }
    __syncthreads();
    uint32_t prefix = carry + (warp > 0 ? warpSums[warp - 1] : 0u) + x - v;
    if (i < total) hist[i] = prefix;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
    __syncthreads();
  }
}

__global__ void k_sort_scatter(const uint64_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                               uint64_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t n,
                               int shift, const uint32_t* __restrict__ hist, uint32_t numBlocks) {
  __shared__ uint32_t running[256];                       // block offset per digit so far
  __shared__ uint32_t warpCnt[KF_SORT_THREADS / 32][256]; // per-warp digit counts of this round
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  running[threadIdx.x] = hist[threadIdx.x * numBlocks + blockIdx.x];
  const uint32_t base = blockIdx.x * KF_SORT_TILE;
  for (int it = 0; it < KF_SORT_ITEMS; it++) {
    for (int w = 0; w < KF_SORT_THREADS / 32; w++) warpCnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = base + it * KF_SORT_THREADS + threadIdx.x;
    const bool valid = i < n;
    uint64_t key = valid ? keysIn[i] : 0ull;
    uint32_t digit = valid ? uint32_t((key >> shift) & 0xffu) : 0x100u + lane;  // invalid: unique
    const uint32_t peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rankInWarp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rankInWarp == 0) warpCnt[warp][digit] = __popc(peers);
    __syncthreads();
    uint32_t pos = 0;
    if (valid) {
      pos = running[digit] + rankInWarp;
      for (int w = 0; w < warp; w++) pos += warpCnt[w][digit];
    }
    __syncthreads();
    {
      uint32_t tot = 0;
      for (int w = 0; w < KF_SORT_THREADS / 32; w++) tot += warpCnt[w][threadIdx.x];
      running[threadIdx.x] += tot;
    }
    if (valid) {
      keysOut[pos] = key;
      valsOut[pos] = valsIn[i];
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// LBVH (Karras 2012).  Internal nodes 0..n-2 (root 0); child code >= 0 internal, < 0 leaf ~j.
// parent[] is indexed [0,n-1) for internal nodes and [n-1, 2n-1) for leaves.
// -------------------------------------------------------------------------------------------------
KF_D int lbvhDelta(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(uint32_t(i) ^ uint32_t(j));
  return __clzll((long long)(a ^ b));
}
__global__ void k_lbvh_hierarchy(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children,
                                 int2* __restrict__ range, int* __restrict__ parent) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (lbvhDelta(keys, n, i, i + 1) - lbvhDelta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lbvhDelta(keys, n, i, i - d);
  int lmax = 2;
  while (lbvhDelta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (lbvhDelta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = lbvhDelta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) >> 1;
    if (lbvhDelta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int left = (lo == gamma) ? ~gamma : gamma;
  const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  children[i] = make_int2(left, right);
  range[i] = make_int2(lo, hi);
  if (left >= 0) parent[left] = i; else parent[n - 1 + gamma] = i;
  if (right >= 0) parent[right] = i; else parent[n - 1 + gamma + 1] = i;
  if (i == 0) parent[0] = -1;
}

KF_D Box6 memberBox(int code, const float* __restrict__ nodeBox, const float* __restrict__ primBox,
                    const uint32_t* __restrict__ vals) {
  return code >= 0 ? loadBox(nodeBox + 6 * code) : loadBox(primBox + 6 * vals[~code]);
}

__global__ void k_lbvh_bounds(int n, const int2* __restrict__ children, const int* __restrict__ parent,
                              const float* __restrict__ primBox, const uint32_t* __restrict__ vals,
                              float* __restrict__ nodeBox, uint32_t* __restrict__ flags) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int node = parent[n - 1 + j];
  while (node >= 0) {
    __threadfence();
    if (atomicAdd(flags + node, 1u) == 0u) return;  // first arrival: the sibling finishes the node
    const int2 c = children[node];
    // child boxes were written by other SMs before their fence: read them from L2 (__ldcg),
    // never through this SM's L1
    Box6 b = c.x >= 0 ? loadBoxCG(nodeBox + 6 * c.x) : loadBox(primBox + 6 * vals[~c.x]);
    boxGrow(b, c.y >= 0 ? loadBoxCG(nodeBox + 6 * c.y) : loadBox(primBox + 6 * vals[~c.y]));
    storeBox(nodeBox + 6 * node, b);
    node = parent[node];
  }
}

// -------------------------------------------------------------------------------------------------
// Top level: binary hierarchy over the instance boxes by top-down binned SAH, on the device.
//
// The top level has at most a few thousand primitives and its quality shows in every ray (on the
// articulated scene, 2 049 overlapping link boxes, a Morton-order hierarchy costs 9.1 top-level node
// visits per ray against 6.3 for this one), so it gets a real SAH sweep: 16 centroid bins per axis,
// cost = area x count of both sides, the cheapest of the 45 candidate planes, stable partition; two
// primitives, or centroids that all coincide, split in the middle.
//
// One block builds the whole tree level by level (no host round trip, no synchronisation of the
// stream): per level every position finds its segment's centroid bounds and bins with atomics in
// global scratch, one thread per segment sweeps the bins, a block-wide prefix sum over the "goes left"
// flags gives every position its place in the stable partition, and a prefix sum over the segments
// numbers the nodes and segments of the next level.  Output in the layout of k_lbvh_hierarchy: internal
// nodes 0..n-2 (root 0), child code >= 0 internal, < 0 leaf ~position, range = positions covered,
// parent[] for internal nodes then leaves, and the position -> primitive permutation in `vals`.
// -------------------------------------------------------------------------------------------------
#define KF_TLAS_BINS 16
#define KF_TLAS_SAH_THREADS 1024

struct TlasSahArgs {
  uint32_t n;
  const float* primBox;
  uint32_t* vals;       // position -> primitive (result)
  uint32_t* valsTmp;
  int2* children;
  int2* range;
  int* parent;
  uint32_t* segOf;      // [2][n]: segment of every position, this level / next level
  int4* segs;           // [2][n / 2 + 1]: node, lo, hi, bin slot (segments of more than two primitives, else -1)
  int* cbounds;         // [bin slots][6] centroid bounds, ordered ints
  uint32_t* binCount;   // [bin slots][3][BINS]
  int* binBox;          // [bin slots][3][BINS][6] ordered ints
  unsigned long long* bestKey;  // [bin slots] cheapest plane so far: cost bits << 32 | axis << 8 | bin
  int4* decision;       // [segments]: axis (-1: middle), split bin, nLeft, index of the first new segment
  float2* decisionF;    // [segments]: centroid lower bound on the axis, bins / extent
  uint32_t* pre;        // [n + 1] prefix sums over positions
  uint32_t* segPre;     // [n / 2 + 2] prefix sums over segments: new segments | new bin slots << 16
  int inShared;         // the working arrays above fit the block's shared memory (tlasSahSharedBytes)
};

#define KF_TLAS_SHARED_MAX 2560u
// Shared memory k_tlas_sah needs to keep its working arrays on chip for n instances.
inline size_t tlasSahSharedBytes(uint32_t n) {
  const size_t maxSeg = n / 2 + 1;
  auto up = [](size_t b) { return (b + 15) & ~size_t(15); };
  return up(sizeof(int4) * 2 * maxSeg) + up(sizeof(int4) * maxSeg) + up(sizeof(float2) * maxSeg) + up(sizeof(float) * 6 * n) +
         up(sizeof(uint32_t) * 2 * n) + 2 * up(sizeof(uint32_t) * n) + up(sizeof(uint32_t) * (n + 1)) +
         up(sizeof(uint32_t) * (maxSeg + 2));
}

// Exclusive prefix sum of value(i), i in [0, total), into out[0..total] (out[total] = sum), by the
// whole block: every thread sums a contiguous run of ceil(total / threads) elements, one scan of the
// thread totals (two barriers) gives each run its base, and the run is walked again to write the
// prefixes.  (The first version scanned blockDim.x elements per round with four barriers a round; with
// three scans per level the barriers were most of the top-level build.)  `sh` holds 33 words.
template <class F>
KF_D void blockExclusiveScan(uint32_t total, uint32_t* __restrict__ out, uint32_t* sh, F value) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  const uint32_t per = (total + blockDim.x - 1) / blockDim.x;
  const uint32_t begin = min(total, threadIdx.x * per), end = min(total, begin + per);
  uint32_t mine = 0;
  for (uint32_t i = begin; i < end; i++) mine += value(i);
  uint32_t x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nWarps ? sh[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    sh[lane] = w;
  }
  __syncthreads();
  uint32_t running = (warp > 0 ? sh[warp - 1] : 0u) + x - mine;
  for (uint32_t i = begin; i < end; i++) {
    out[i] = running;
    running += value(i);
  }
  if (threadIdx.x == blockDim.x - 1) out[total] = sh[nWarps - 1];
  __syncthreads();
}

KF_D float sahHalfArea(const float* lo, const float* hi) {
  const float dx = csub(hi[0], lo[0]), dy = csub(hi[1], lo[1]), dz = csub(hi[2], lo[2]);
  return cadd(cadd(cmul(dx, dy), cmul(dy, dz)), cmul(dz, dx));
}
KF_D int sahBin(float c, float clo, float scale) {
  return min(KF_TLAS_BINS - 1, max(0, int(cmul(csub(c, clo), scale))));
}

__global__ void __launch_bounds__(KF_TLAS_SAH_THREADS, 1) k_tlas_sah(TlasSahArgs a) {
  __shared__ uint32_t sScan[33];
  __shared__ uint32_t sCount[3];  // segments of this level, nodes allocated so far, bin slots of this level
  const uint32_t n = a.n, T = blockDim.x, tid = threadIdx.x;
  const uint32_t maxSeg = n / 2 + 1;
  // Working arrays: in shared memory when they fit (a.inShared: up to KF_TLAS_SHARED_MAX instances), else in
  // the global scratch.  One block on one SM pays a full memory round trip for every dependent access, and
  // a level is a dozen phases of two or three of them.  (Measured, 2 049 instances: 0.58 -> 0.54 ms for the
  // hierarchy -- the bins, which do not fit on chip, and the barriers behind their atomics are what is left.)
  extern __shared__ __align__(16) unsigned char sahShared[];
  uint32_t* vals = a.vals;
  uint32_t* valsOther = a.valsTmp;
  uint32_t* segOf = a.segOf;
  int4* segs = a.segs;
  uint32_t* pre = a.pre;
  uint32_t* segPre = a.segPre;
  int4* decision = a.decision;
  float2* decisionF = a.decisionF;
  const float* primBox = a.primBox;
  if (a.inShared) {
    unsigned char* q = sahShared;
    auto take = [&](size_t bytes) { unsigned char* r = q; q += (bytes + 15) & ~size_t(15); return r; };
    segs = reinterpret_cast<int4*>(take(sizeof(int4) * 2 * maxSeg));
    decision = reinterpret_cast<int4*>(take(sizeof(int4) * maxSeg));
    decisionF = reinterpret_cast<float2*>(take(sizeof(float2) * maxSeg));
    float* box = reinterpret_cast<float*>(take(sizeof(float) * 6 * n));
    segOf = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * 2 * n));
    vals = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * n));
    valsOther = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * n));
    pre = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * (n + 1)));
    segPre = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * (maxSeg + 2)));
    for (uint32_t i = tid; i < 6 * n; i += T) box[i] = a.primBox[i];
    primBox = box;
  }
  uint32_t* segOfNext = segOf + n;
  int4* segsNext = segs + maxSeg;
  for (uint32_t p = tid; p < n; p += T) {
    vals[p] = p;
    segOf[p] = 0;
  }
  if (tid == 0) {
    segs[0] = make_int4(0, 0, int(n) - 1, n > 2 ? 0 : -1);
    a.parent[0] = -1;
    sCount[0] = 1;
    sCount[1] = 1;
    sCount[2] = n > 2 ? 1 : 0;
  }
  __syncthreads();
  for (;;) {
    const uint32_t nSeg = sCount[0], nBig = sCount[2];
    if (nSeg == 0) break;
    // ---- centroid bounds and bins of every segment with more than two primitives --------------
    for (uint32_t i = tid; i < nBig * 6; i += T) a.cbounds[i] = floatToOrdered((i % 6) < 3 ? 3.0e38f : -3.0e38f);
    for (uint32_t i = tid; i < nBig * 3 * KF_TLAS_BINS; i += T) a.binCount[i] = 0;
    for (uint32_t i = tid; i < nBig; i += T) a.bestKey[i] = ~0ull;
    for (uint32_t i = tid; i < nBig * 3 * KF_TLAS_BINS * 6; i += T)
      a.binBox[i] = floatToOrdered((i % 6) < 3 ? 3.0e38f : -3.0e38f);
    __syncthreads();
    for (uint32_t p = tid; p < n; p += T) {
      const uint32_t s = segOf[p];
      if (s == 0xffffffffu) continue;  // a finished leaf
      const int slot = segs[s].w;
      if (slot < 0) continue;
      const float* b = primBox + 6 * size_t(vals[p]);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float c = cmul(0.5f, cadd(b[k], b[3 + k]));
        atomicMin(a.cbounds + 6 * slot + k, floatToOrdered(c));
        atomicMax(a.cbounds + 6 * slot + 3 + k, floatToOrdered(c));
      }
    }
    __syncthreads();
    for (uint32_t p = tid; p < n; p += T) {
      const uint32_t s = segOf[p];
      if (s == 0xffffffffu) continue;
      const int slot = segs[s].w;
      if (slot < 0) continue;
      const float* b = primBox + 6 * size_t(vals[p]);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float clo = orderedToFloat(a.cbounds[6 * slot + k]), chi = orderedToFloat(a.cbounds[6 * slot + 3 + k]);
        const float ext = csub(chi, clo);
        if (!(ext > 0.0f)) continue;
        const int bin = sahBin(cmul(0.5f, cadd(b[k], b[3 + k])), clo, cdiv(float(KF_TLAS_BINS), ext));
        const size_t bi = (size_t(slot) * 3 + k) * KF_TLAS_BINS + bin;
        atomicAdd(a.binCount + bi, 1u);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          atomicMin(a.binBox + 6 * bi + c, floatToOrdered(b[c]));
          atomicMax(a.binBox + 6 * bi + 3 + c, floatToOrdered(b[3 + c]));
        }
      }
    }
    __syncthreads();
    // ---- sweep: sixteen lanes per (segment, axis), one bin each.  Suffix and prefix unions of the bin
    // boxes by shuffles, the cost of the plane after every bin, the cheapest of the sixteen by a shuffle
    // reduction, and the cheapest of the three axes by an atomic minimum on (cost, axis, bin) -- a lower
    // axis or bin wins a tie, as in a serial sweep in that order.  (One thread per segment walking
    // 3 x 2 x 16 bins of 7 words was the slowest phase of the build: 600 dependent loads per level.)
    for (uint32_t item = tid >> 4; item < ((nBig * 3u + (T >> 4) - 1u) / (T >> 4)) * (T >> 4); item += T >> 4) {
      const uint32_t lane16 = tid & 15u;
      const bool live = item < nBig * 3u;
      const uint32_t slot = live ? item / 3u : 0u, k = live ? item % 3u : 0u;
      const float clo = orderedToFloat(a.cbounds[6 * slot + k]), chi = orderedToFloat(a.cbounds[6 * slot + 3 + k]);
      const bool axisOk = live && csub(chi, clo) > 0.0f;
      const size_t bi = (size_t(slot) * 3 + k) * KF_TLAS_BINS + lane16;
      uint32_t cnt = axisOk ? a.binCount[bi] : 0u;
      float lo[3], hi[3];
#pragma unroll
      for (int q = 0; q < 3; q++) {
        lo[q] = axisOk ? orderedToFloat(a.binBox[6 * bi + q]) : 3.0e38f;
        hi[q] = axisOk ? orderedToFloat(a.binBox[6 * bi + 3 + q]) : -3.0e38f;
      }
      // inclusive prefix (bins 0 .. b) and suffix (bins b .. 15) of count and box
      uint32_t pc = cnt, sc = cnt;
      float plo[3] = {lo[0], lo[1], lo[2]}, phi[3] = {hi[0], hi[1], hi[2]};
      float slo[3] = {lo[0], lo[1], lo[2]}, shi[3] = {hi[0], hi[1], hi[2]};
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const uint32_t pcn = __shfl_up_sync(0xffffffffu, pc, o, 16), scn = __shfl_down_sync(0xffffffffu, sc, o, 16);
        float pl[3], ph[3], sl[3], sh2[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
          pl[q] = __shfl_up_sync(0xffffffffu, plo[q], o, 16);
          ph[q] = __shfl_up_sync(0xffffffffu, phi[q], o, 16);
          sl[q] = __shfl_down_sync(0xffffffffu, slo[q], o, 16);
          sh2[q] = __shfl_down_sync(0xffffffffu, shi[q], o, 16);
        }
        if (int(lane16) >= o) {
          pc += pcn;
#pragma unroll
          for (int q = 0; q < 3; q++) { plo[q] = fminf(plo[q], pl[q]); phi[q] = fmaxf(phi[q], ph[q]); }
        }
        if (int(lane16) + o < 16) {
          sc += scn;
#pragma unroll
          for (int q = 0; q < 3; q++) { slo[q] = fminf(slo[q], sl[q]); shi[q] = fmaxf(shi[q], sh2[q]); }
        }
      }
      // the plane after bin b: left = prefix(b), right = suffix(b + 1)
      const uint32_t rc = __shfl_down_sync(0xffffffffu, sc, 1, 16);
      float rlo[3], rhi[3];
#pragma unroll
      for (int q = 0; q < 3; q++) {
        rlo[q] = __shfl_down_sync(0xffffffffu, slo[q], 1, 16);
        rhi[q] = __shfl_down_sync(0xffffffffu, shi[q], 1, 16);
      }
      unsigned long long key = ~0ull;
      if (axisOk && lane16 < 15u && pc > 0u && rc > 0u) {
        const float cost = cadd(cmul(sahHalfArea(plo, phi), float(pc)), cmul(sahHalfArea(rlo, rhi), float(rc)));
        key = ((unsigned long long)__float_as_uint(cost) << 32) | (unsigned long long)(k << 8 | lane16);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o, 16);
        key = other < key ? other : key;
      }
      if (live && lane16 == 0u && key != ~0ull) atomicMin(a.bestKey + slot, key);
    }
    __syncthreads();
    for (uint32_t s = tid; s < nSeg; s += T) {
      const int slot = segs[s].w;
      int bestAxis = -1, bestSplit = 0;
      float bestClo = 0.0f, bestScale = 0.0f;
      const unsigned long long key = slot >= 0 ? __ldcg(a.bestKey + slot) : ~0ull;
      if (key != ~0ull) {
        bestAxis = int((key >> 8) & 3ull);
        bestSplit = int(key & 15ull);
        bestClo = orderedToFloat(a.cbounds[6 * slot + bestAxis]);
        bestScale = cdiv(float(KF_TLAS_BINS), csub(orderedToFloat(a.cbounds[6 * slot + 3 + bestAxis]), bestClo));
      }
      decision[s] = make_int4(bestAxis, bestSplit, 0, 0);
      decisionF[s] = make_float2(bestClo, bestScale);
    }
    __syncthreads();
    // ---- stable partition: a prefix sum over the "goes left" flags of all positions ----------------
    auto goesLeft = [&](uint32_t p) -> uint32_t {
      const uint32_t s = segOf[p];
      if (s == 0xffffffffu) return 0u;
      const int4 sg = segs[s];
      const int4 d = decision[s];
      if (d.x < 0) return int(p) <= sg.y + (sg.z - sg.y) / 2 ? 1u : 0u;  // mid = lo + (cnt - 1) / 2
      const float* b = primBox + 6 * size_t(vals[p]);
      const float2 df = decisionF[s];
      return sahBin(cmul(0.5f, cadd(b[d.x], b[3 + d.x])), df.x, df.y) <= d.y ? 1u : 0u;
    };
    blockExclusiveScan(n, pre, sScan, goesLeft);
    // ---- children, nodes and segments of the next level ---------------------------------------------
    // new segments (children of more than one primitive) in the low half, new bin slots (more than two)
    // in the high half: at most n / 2 and n / 3 of them, n <= 65 536, so one scan carries both
    auto newSegments = [&](uint32_t s) -> uint32_t {
      const int4 sg = segs[s];
      const uint32_t nLeft = pre[sg.z + 1] - pre[sg.y], nRight = uint32_t(sg.z - sg.y + 1) - nLeft;
      return (nLeft > 1 ? 1u : 0u) + (nRight > 1 ? 1u : 0u) + (((nLeft > 2 ? 1u : 0u) + (nRight > 2 ? 1u : 0u)) << 16);
    };
    blockExclusiveScan(nSeg, segPre, sScan, newSegments);
    const uint32_t nodeBase = sCount[1];
    for (uint32_t s = tid; s < nSeg; s += T) {
      const int4 sg = segs[s];
      const int lo = sg.y, hi = sg.z;
      const int nLeft = int(pre[hi + 1] - pre[lo]);
      const int mid = lo + nLeft - 1;
      uint32_t next = segPre[s] & 0xffffu, nextBig = segPre[s] >> 16;
      int4 d = decision[s];
      d.z = nLeft;
      d.w = int(next);
      decision[s] = d;
      int code[2];
      const int clo2[2] = {lo, mid + 1}, chi2[2] = {mid, hi};
      for (int side = 0; side < 2; side++) {
        if (clo2[side] == chi2[side]) {
          code[side] = ~clo2[side];
          a.parent[size_t(n) - 1 + clo2[side]] = sg.x;
        } else {
          const int node = int(nodeBase + next);
          code[side] = node;
          a.parent[node] = sg.x;
          segsNext[next] = make_int4(node, clo2[side], chi2[side], chi2[side] - clo2[side] + 1 > 2 ? int(nextBig++) : -1);
          next++;
        }
      }
      a.children[sg.x] = make_int2(code[0], code[1]);
      a.range[sg.x] = make_int2(lo, hi);
    }
    __syncthreads();
    for (uint32_t p = tid; p < n; p += T) {
      const uint32_t s = segOf[p];
      if (s == 0xffffffffu) {
        valsOther[p] = vals[p];
        segOfNext[p] = 0xffffffffu;
        continue;
      }
      const int4 sg = segs[s];
      const int4 d = decision[s];
      const uint32_t rankLeft = pre[p] - pre[sg.y];
      const bool left = pre[p + 1] != pre[p];
      const uint32_t cnt = uint32_t(sg.z - sg.y + 1), nLeft = uint32_t(d.z);
      const uint32_t dst = left ? uint32_t(sg.y) + rankLeft : uint32_t(sg.y) + nLeft + (p - uint32_t(sg.y)) - rankLeft;
      valsOther[dst] = vals[p];
      uint32_t ns = 0xffffffffu;
      if (left) {
        if (nLeft > 1) ns = uint32_t(d.w);
      } else if (cnt - nLeft > 1) {
        ns = uint32_t(d.w) + (nLeft > 1 ? 1u : 0u);
      }
      segOfNext[dst] = ns;
    }
    __syncthreads();
    if (tid == 0) {
      sCount[1] = nodeBase + (segPre[nSeg] & 0xffffu);
      sCount[0] = segPre[nSeg] & 0xffffu;
      sCount[2] = segPre[nSeg] >> 16;
    }
    { uint32_t* t = vals; vals = valsOther; valsOther = t; }
    { uint32_t* t = segOf; segOf = segOfNext; segOfNext = t; }
    { int4* t = segs; segs = segsNext; segsNext = t; }
    __syncthreads();
  }
  if (vals != a.vals)
    for (uint32_t p = tid; p < n; p += T) a.vals[p] = vals[p];  // (from shared memory, or from the other global copy)
}

// Sum of the surface areas of the binary internal nodes: the SAH cost of the hierarchy up to a
// constant.  A refit keeps the topology, so this number says how far the moving instances have
// stretched it since it was built.
__global__ void k_area_sum(const float* __restrict__ nodeBox, uint32_t nInternal, float* __restrict__ out) {
  float s = 0.0f;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nInternal; i += gridDim.x * blockDim.x)
    s += boxArea(loadBox(nodeBox + 6 * i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// -------------------------------------------------------------------------------------------------
// Wide-node emission
// -------------------------------------------------------------------------------------------------
KF_D uint32_t exponentFor(float extent) {
  // smallest e with extent <= 255 * 2^(e-127); biased exponent byte in [1, 254]
  const float v = extent * (1.0f / 255.0f);
  int e = int((__float_as_uint(v) >> 23) & 0xffu) + 1;  // 2^(e-127) > v for normal v
  if (!(v > 0.0f)) e = 1;
  return uint32_t(min(max(e, 1), 239));  // + 15 still fits the byte (KF_EXP_BIASED)
}

// Quantises the member boxes of one wide node.  slotBox[s] is ignored where slotUsed bit s is 0.
// The grid origin sits a little outside the node box and planes are rounded outward with a guard of
// 1/128 step, so every stored plane is at least ~0.008 step outside the true one: that margin covers
// the 2^-9-step rounding of the biased dequantisation in intersectNode() (kf_traverse.cuh).
KF_D void quantiseNode(Node8& nd, const Box6& nbIn, const Box6* slotBox, uint32_t slotUsed) {
  Box6 nb = nbIn;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float eps = (nb.hi[k] - nb.lo[k]) * (1.0f / 8192.0f);
    nb.lo[k] -= eps;
    nb.hi[k] += eps;
  }
  nd.px = nb.lo[0];
  nd.py = nb.lo[1];
  nd.pz = nb.lo[2];
  const uint32_t ex = exponentFor(nb.hi[0] - nb.lo[0]);
  const uint32_t ey = exponentFor(nb.hi[1] - nb.lo[1]);
  const uint32_t ez = exponentFor(nb.hi[2] - nb.lo[2]);
  nd.ex = uint8_t(ex + 15u);  // stored with the 2^15 of planeFloat() folded in (kf_traverse.cuh)
  nd.ey = uint8_t(ey + 15u);
  nd.ez = uint8_t(ez + 15u);
  const float isx = 1.0f / __uint_as_float(ex << 23);
  const float isy = 1.0f / __uint_as_float(ey << 23);
  const float isz = 1.0f / __uint_as_float(ez << 23);
  const float g = 1.0f / 128.0f;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    if (!((slotUsed >> s) & 1u)) {
      nd.qlox[s] = nd.qloy[s] = nd.qloz[s] = 255;
      nd.qhix[s] = nd.qhiy[s] = nd.qhiz[s] = 0;
      continue;
    }
    const Box6& b = slotBox[s];
    nd.qlox[s] = uint8_t(fminf(fmaxf(floorf((b.lo[0] - nb.lo[0]) * isx - g), 0.0f), 255.0f));
    nd.qloy[s] = uint8_t(fminf(fmaxf(floorf((b.lo[1] - nb.lo[1]) * isy - g), 0.0f), 255.0f));
    nd.qloz[s] = uint8_t(fminf(fmaxf(floorf((b.lo[2] - nb.lo[2]) * isz - g), 0.0f), 255.0f));
    nd.qhix[s] = uint8_t(fminf(fmaxf(ceilf((b.hi[0] - nb.lo[0]) * isx + g), 0.0f), 255.0f));
    nd.qhiy[s] = uint8_t(fminf(fmaxf(ceilf((b.hi[1] - nb.lo[1]) * isy + g), 0.0f), 255.0f));
    nd.qhiz[s] = uint8_t(fminf(fmaxf(ceilf((b.hi[2] - nb.lo[2]) * isz + g), 0.0f), 255.0f));
  }
}

#define KF_MEMBER_EMPTY 0x7fffffff
#ifndef KF_LEAF_MAX
#define KF_LEAF_MAX 2
#endif
static_assert(KF_LEAF_MAX <= 2, "Node8::triMask holds two bits per leaf slot");

struct CollapseArgs {
  int n;                      // primitives
  const int2* children;       // binary
  const int2* range;          // binary
  const float* nodeBox;       // binary internal boxes
  const float* primBox;       // padded primitive boxes (original order)
  const uint32_t* vals;       // sorted position -> primitive
  Node8* outNodes;            // wide nodes
  uint32_t* outPrim;          // leaf order -> primitive
  int* wideBinary;            // wide node -> binary internal node it collapses
  int* wideMembers;           // 8 member codes per wide node (slot order), for refit
  uint32_t* counters;         // [0] wide nodes allocated, [1] leaf primitives allocated,
                              // [2], [3] = [lo, hi) of the level being collapsed,
                              // [4] = levels collapsed so far = depth of the wide tree
  uint32_t* slotOfInst;       // top level only: instance -> index of its InstNode in outNodes
};

KF_D int memberCount(int code, const int2* __restrict__ range) {
  if (code < 0) return 1;
  const int2 r = range[code];
  return r.y - r.x + 1;
}
KF_D int memberFirst(int code, const int2* __restrict__ range) { return code < 0 ? ~code : range[code].x; }

// One thread per wide node of the current level [lo, hi).
//
// TLAS == false (bottom level): leaves hold up to KF_LEAF_MAX triangles; triMask/primBase as
// documented at Node8.
// TLAS == true (top level): every child is addressed like an internal child (imask = present mask,
// no triangle fields, children consecutive from childBase in slot order) and is either a real
// node or an InstNode record; primBase holds the mask of the slots that are real nodes.  Instances
// thereby take part in the octant-ordered front-to-back traversal instead of being entered in
// storage order.
template <bool TLAS>
KF_D void collapseNode(const CollapseArgs& a, uint32_t w) {
  constexpr int LEAF_MAX = TLAS ? 1 : KF_LEAF_MAX;
  const int b = a.wideBinary[w];
  if (TLAS && b < 0) {  // an InstNode slot: filled by k_instance_setup, nothing to collapse
    for (int s = 0; s < 8; s++) a.wideMembers[8 * w + s] = KF_MEMBER_EMPTY;
    return;
  }
  // (wideMembers is what a refit re-quantises from: the top level keeps it, bottom levels are never refitted)
  int mem[8];
  int m = 0;
  {
    const int2 c = a.children[b];
    mem[m++] = c.x;
    mem[m++] = c.y;
  }
  // expand the largest-area expandable member until 8 children
  while (m < 8) {
    int best = -1;
    float bestArea = -1.0f;
    for (int k = 0; k < m; k++) {
      if (mem[k] >= 0 && memberCount(mem[k], a.range) > LEAF_MAX) {
        const float ar = boxArea(loadBox(a.nodeBox + 6 * mem[k]));
        if (ar > bestArea) { bestArea = ar; best = k; }
      }
    }
    if (best < 0) break;
    const int2 c = a.children[mem[best]];
    mem[best] = c.x;
    mem[m++] = c.y;
  }
  const Box6 nb = loadBox(a.nodeBox + 6 * b);
  float cx = 0.5f * (nb.lo[0] + nb.hi[0]), cy = 0.5f * (nb.lo[1] + nb.hi[1]), cz = 0.5f * (nb.lo[2] + nb.hi[2]);
  Box6 mb[8];
  float dx[8], dy[8], dz[8];
  for (int k = 0; k < m; k++) {
    mb[k] = memberBox(mem[k], a.nodeBox, a.primBox, a.vals);
    dx[k] = 0.5f * (mb[k].lo[0] + mb[k].hi[0]) - cx;
    dy[k] = 0.5f * (mb[k].lo[1] + mb[k].hi[1]) - cy;
    dz[k] = 0.5f * (mb[k].lo[2] + mb[k].hi[2]) - cz;
  }
  // greedy octant slot assignment: slot s prefers the child farthest along (+-1,+-1,+-1)_s
  int slotOf[8];
  for (int k = 0; k < 8; k++) slotOf[k] = -1;
  uint32_t used = 0;
  for (int round = 0; round < m; round++) {
    float bestCost = -3.0e38f;
    int bk = -1, bs = -1;
    for (int k = 0; k < m; k++) {
      if (slotOf[k] >= 0) continue;
      for (int s = 0; s < 8; s++) {
        if ((used >> s) & 1u) continue;
        const float cost = ((s & 1) ? dx[k] : -dx[k]) + ((s & 2) ? dy[k] : -dy[k]) + ((s & 4) ? dz[k] : -dz[k]);
        if (cost > bestCost) { bestCost = cost; bk = k; bs = s; }
      }
    }
    slotOf[bk] = bs;
    used |= 1u << bs;
  }
  int slotMem[8];
  Box6 slotBox[8];
  for (int s = 0; s < 8; s++) slotMem[s] = KF_MEMBER_EMPTY;
  for (int k = 0; k < m; k++) {
    slotMem[slotOf[k]] = mem[k];
    slotBox[slotOf[k]] = mb[k];
  }
  // count internal children / leaf primitives, allocate
  uint32_t nInternal = 0, nPrims = 0;
  for (int s = 0; s < 8; s++) {
    if (slotMem[s] == KF_MEMBER_EMPTY) continue;
    const int cnt = memberCount(slotMem[s], a.range);
    if (TLAS || cnt > LEAF_MAX) nInternal++; else nPrims += cnt;
  }
  const uint32_t childBase = nInternal ? atomicAdd(a.counters + 0, nInternal) : 0u;
  const uint32_t primBase = nPrims ? atomicAdd(a.counters + 1, nPrims) : 0u;
  Node8 nd;
  nd.childBase = childBase;
  nd.primBase = primBase;
  uint32_t imask = 0, ci = 0, po = 0, realNodes = 0, triMask = 0;
  for (int s = 0; s < 8; s++) {
    const int code = slotMem[s];
    if (TLAS) a.wideMembers[8 * w + s] = code;
    if (code == KF_MEMBER_EMPTY) continue;
    const int cnt = memberCount(code, a.range);
    if (cnt > LEAF_MAX) {
      imask |= 1u << s;
      realNodes |= 1u << s;
      a.wideBinary[childBase + ci] = code;
      ci++;
    } else if (TLAS) {
      const uint32_t prim = a.vals[memberFirst(code, a.range)];
      imask |= 1u << s;
      a.wideBinary[childBase + ci] = ~int(prim);
      a.slotOfInst[prim] = childBase + ci;
      ci++;
    } else {
      const int first = memberFirst(code, a.range);
      triMask |= ((1u << cnt) - 1u) << (2 * s);
      for (int q = 0; q < cnt; q++) a.outPrim[primBase + po + q] = a.vals[first + q];
      po += cnt;
    }
  }
  if (TLAS) nd.primBase = realNodes;
  nd.imask = uint8_t(imask);
  nd.triMask = triMask * 0x00010001u;
  nd.reserved = 0;
  quantiseNode(nd, nb, slotBox, used);
  a.outNodes[w] = nd;
}

// Top level: the whole collapse in one launch of one block (a top level has at most a few thousand
// nodes per level), so that a build needs no host round trip at all: the block walks the levels itself.
// counters are read and written around L1 (the allocations of collapseNode are L2 atomics).
#define KF_COLLAPSE_ALL_THREADS 512
template <bool TLAS>
__global__ void __launch_bounds__(KF_COLLAPSE_ALL_THREADS) k_collapse_all(CollapseArgs a) {
  for (;;) {
    const uint32_t lo = __ldcg(a.counters + 2), hi = __ldcg(a.counters + 3);
    if (lo >= hi) break;
    for (uint32_t w = lo + threadIdx.x; w < hi; w += blockDim.x) collapseNode<TLAS>(a, w);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      __stcg(a.counters + 4, __ldcg(a.counters + 4) + 1u);
      __stcg(a.counters + 2, hi);
      __stcg(a.counters + 3, __ldcg(a.counters + 0));
    }
    __syncthreads();
  }
}

// Top-level root for a single instance: node 0 with one child, the InstNode at index 1.
__global__ void k_single_instance_root(const float* __restrict__ primBox, Node8* outNodes, int* wideBinary,
                                       int* wideMembers, uint32_t* slotOfInst, float* rootBox) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Box6 nb = loadBox(primBox);
  Node8 nd;
  nd.childBase = 1;
  nd.primBase = 0;  // no real-node children
  nd.imask = 1;
  Box6 slotBox[8];
  for (int s = 0; s < 8; s++) { wideMembers[s] = KF_MEMBER_EMPTY; wideMembers[8 + s] = KF_MEMBER_EMPTY; }
  nd.triMask = 0;
  nd.reserved = 0;
  slotBox[0] = nb;
  quantiseNode(nd, nb, slotBox, 1u);
  outNodes[0] = nd;
  wideBinary[0] = 0;
  wideBinary[1] = ~0;
  slotOfInst[0] = 1;
  storeBox(rootBox, nb);
}

// Refit: re-quantise every wide node from its members' freshly recomputed boxes.  wideBinary < 0
// marks the InstNode slots of a top-level array, which k_instance_setup rewrites instead.
// nWide: the number of wide nodes, on the device (the host only knows its bound).
__global__ void k_requantise(const uint32_t* __restrict__ nWide, const int* __restrict__ wideMembers, const int* __restrict__ wideBinary,
                             const float* __restrict__ nodeBox, const float* __restrict__ primBox,
                             const uint32_t* __restrict__ vals, Node8* nodes, int singleLeafN) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= *nWide) return;
  if (wideBinary && wideBinary[w] < 0) return;
  Box6 nb;
  boxReset(nb);
  Box6 slotBox[8];
  uint32_t used = 0;
  if (singleLeafN > 0) {
    for (int i = 0; i < singleLeafN; i++) boxGrow(nb, loadBox(primBox + 6 * i));
    slotBox[0] = nb;
    used = 1u;
  } else {
    for (int s = 0; s < 8; s++) {
      const int code = wideMembers[8 * w + s];
      if (code == KF_MEMBER_EMPTY) continue;
      slotBox[s] = memberBox(code, nodeBox, primBox, vals);
      boxGrow(nb, slotBox[s]);
      used |= 1u << s;
    }
  }
  Node8 nd = nodes[w];
  quantiseNode(nd, nb, slotBox, used);
  nodes[w] = nd;
}

}  // namespace kf
