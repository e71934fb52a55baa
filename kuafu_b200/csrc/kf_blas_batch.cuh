// kf_blas_batch.cuh -- one build for all dirty geometries (reference src/core/rt/rt.cpp:142-370: the
// reference rebuilds every bottom-level structure whenever the instance set changes, "very heavy").
//
// A per-geometry build is bound by its launches below ~20 k triangles (about 70 of them, 0.45 ms): config 1
// spent 4.3 ms building ten geometries one after the other.  Here the triangles of all geometries of a
// batch are concatenated, and every stage runs once over the lot:
//   boxes + per-geometry scene boxes -> keys = geometry << 54 | 54-bit Morton code inside the geometry's box
//   -> one radix sort -> one Karras hierarchy over all keys: the keys of a geometry share a prefix no
//   other key has, so its triangles form one subtree, whose root is found from the ranges (k_batch_roots)
//   -> one bottom-up box pass -> the level-by-level collapse of all geometries in lock step, every geometry
//   allocating in its own region of the node array (so that child indices stay relative to its root)
//   -> nodes copied to their final, exactly sized place; bounding spheres, leaf-ordered triangles and
//   shading records written by kernels that run over all geometries at once.
// The host looks at the device once per batch (node counts, depths, boxes), not once per geometry.
#pragma once

#include "kf_bvh_build.cuh"
#include "kf_common.cuh"

namespace kf {

#define KF_BATCH_GEOM_BITS 10
#define KF_BATCH_MAX_GEOMS (1u << KF_BATCH_GEOM_BITS)
#define KF_BATCH_MORTON_BITS 18  // per axis; 3 x 18 + 10 = 64 key bits

struct BatchGeom {
  const KfrtVertex* verts;
  const uint32_t* idx;
  const uint32_t* matIndex;
  uint32_t nVerts, nTris;
  uint32_t triOffset;    // first triangle of the geometry in the concatenated arrays
  uint32_t nodeOffset;   // first wide node of its region in the scratch node array (nTris + 1 nodes)
  Node8* nodesAlloc;     // final storage: header record, then the nodes (set before the finishing kernels)
  Tri48* tris;
  ShadeTri* shade;
  float4* faceNormal;    // object-space e1 x e2 per primitive
};

// Per-geometry state of the collapse, 8 words: [0] wide nodes allocated, [1] leaf primitives allocated,
// [2], [3] = [lo, hi) of the level being collapsed, [4] levels so far, [5] binary root (or ~position of a
// single triangle), [6], [7] unused.
#define KF_BATCH_COUNTERS 8

KF_D uint32_t batchGeomOf(const BatchGeom* __restrict__ geoms, uint32_t nGeoms, uint32_t tri) {
  uint32_t lo = 0, hi = nGeoms;  // last geometry whose triOffset <= tri
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (geoms[mid].triOffset <= tri) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_batch_init(int* __restrict__ sceneBoxes, uint32_t nGeoms) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * nGeoms) sceneBoxes[i] = floatToOrdered((i % 6) < 3 ? 3.0e38f : -3.0e38f);
}

// One thread per triangle of the batch: padded box, contribution to its geometry's box.
__global__ void k_batch_tri_boxes(const BatchGeom* __restrict__ geoms, uint32_t nGeoms, uint32_t nTris,
                                  float* __restrict__ primBox, uint32_t* __restrict__ primGeom,
                                  int* __restrict__ sceneBoxes) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nTris) return;
  const uint32_t g = batchGeomOf(geoms, nGeoms, t);
  const BatchGeom& G = geoms[g];
  const uint32_t local = t - G.triOffset;
  Box6 b;
  boxReset(b);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float* p = G.verts[G.idx[3 * local + c]].pos;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      b.lo[k] = fminf(b.lo[k], p[k]);
      b.hi[k] = fmaxf(b.hi[k], p[k]);
    }
  }
  boxPad(b);
  storeBox(primBox + 6 * size_t(t), b);
  primGeom[t] = g;
  // one atomic per warp and bound where the whole warp belongs to one geometry (the common case)
  const uint32_t active = __activemask();
  const uint32_t same = __match_any_sync(active, g);
  if (same == active) {
    int wlo[3], whi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      wlo[k] = __reduce_min_sync(active, floatToOrdered(b.lo[k]));
      whi[k] = __reduce_max_sync(active, floatToOrdered(b.hi[k]));
    }
    if ((threadIdx.x & 31u) == uint32_t(__ffs(active) - 1)) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        atomicMin(sceneBoxes + 6 * g + k, wlo[k]);
        atomicMax(sceneBoxes + 6 * g + 3 + k, whi[k]);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(sceneBoxes + 6 * g + k, floatToOrdered(b.lo[k]));
      atomicMax(sceneBoxes + 6 * g + 3 + k, floatToOrdered(b.hi[k]));
    }
  }
}

__global__ void k_batch_morton(const float* __restrict__ primBox, const uint32_t* __restrict__ primGeom, uint32_t n,
                               const int* __restrict__ sceneBoxes, uint64_t* __restrict__ keys,
                               uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t g = primGeom[i];
  uint64_t code = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float lo = orderedToFloat(sceneBoxes[6 * g + k]), hi = orderedToFloat(sceneBoxes[6 * g + 3 + k]);
    const float c = 0.5f * (primBox[6 * size_t(i) + k] + primBox[6 * size_t(i) + 3 + k]);
    const float ext = hi - lo;
    float f = ext > 0.0f ? (c - lo) / ext : 0.0f;
    const float cells = float(1u << KF_BATCH_MORTON_BITS);
    f = fminf(fmaxf(f * cells, 0.0f), cells - 1.0f);
    code |= expandBits21(uint64_t(f)) << (2 - k);  // f < 2^18: bits 3i, i < 18
  }
  keys[i] = (uint64_t(g) << (3 * KF_BATCH_MORTON_BITS)) | code;
  vals[i] = i;
}

// One thread per geometry: the binary node that covers exactly its triangles (sorted positions
// [triOffset, triOffset + nTris)) is the Karras node at one of the two ends of that range; the collapse
// starts from it.
__global__ void k_batch_roots(const BatchGeom* __restrict__ geoms, uint32_t nGeoms, const int2* __restrict__ range,
                              uint32_t totalTris, uint32_t* __restrict__ counters, int* __restrict__ wideBinary) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nGeoms) return;
  const BatchGeom& G = geoms[g];
  const int lo = int(G.triOffset), hi = int(G.triOffset + G.nTris) - 1;
  int root;
  if (G.nTris == 1) {
    root = ~lo;
  } else if (lo == 0 && hi == int(totalTris) - 1) {
    root = 0;
  } else {
    const int2 a = lo < int(totalTris) - 1 ? range[lo] : make_int2(-1, -1);
    root = (a.x == lo && a.y == hi) ? lo : hi;
  }
  uint32_t* c = counters + KF_BATCH_COUNTERS * g;
  c[0] = 1u; c[1] = 0u; c[2] = 0u; c[3] = 1u; c[4] = 0u; c[5] = uint32_t(root); c[6] = c[7] = 0u;
  wideBinary[G.nodeOffset] = root;
}

struct BatchCollapseArgs {
  const BatchGeom* geoms;
  uint32_t nGeoms;
  CollapseArgs base;   // global arrays; outNodes / wideBinary / outPrim / counters are offset per geometry
};

// Root of a geometry with a single triangle (no binary node to collapse): one leaf slot.
KF_D void batchSingleTriangleRoot(const CollapseArgs& a, int pos) {
  const uint32_t prim = a.vals[pos];
  const Box6 nb = loadBox(a.primBox + 6 * size_t(prim));
  Node8 nd;
  nd.childBase = 0;
  nd.primBase = 0;
  nd.imask = 0;
  nd.triMask = 1u * 0x00010001u;
  nd.reserved = 0;
  Box6 slotBox[8];
  slotBox[0] = nb;
  quantiseNode(nd, nb, slotBox, 1u);
  a.outNodes[0] = nd;
  a.outPrim[0] = prim;
  atomicAdd(a.counters + 1, 1u);
}

// grid.y = geometry; the blocks of a geometry stride over the wide nodes of its current level.
__global__ void k_batch_collapse_level(BatchCollapseArgs b) {
  const uint32_t g = blockIdx.y;
  const BatchGeom& G = b.geoms[g];
  CollapseArgs a = b.base;
  a.outNodes += G.nodeOffset;
  a.wideBinary += G.nodeOffset;
  a.outPrim += G.triOffset;
  a.counters += KF_BATCH_COUNTERS * g;
  a.wideMembers = nullptr;
  const uint32_t lo = a.counters[2], hi = a.counters[3];
  for (uint32_t w = lo + blockIdx.x * blockDim.x + threadIdx.x; w < hi; w += gridDim.x * blockDim.x) {
    if (w == 0 && a.wideBinary[0] < 0) batchSingleTriangleRoot(a, ~a.wideBinary[0]);
    else collapseNode<false>(a, w);
  }
}

__global__ void k_batch_next_level(uint32_t* __restrict__ counters, uint32_t nGeoms) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nGeoms) return;
  uint32_t* c = counters + KF_BATCH_COUNTERS * g;
  if (c[3] > c[2]) c[4]++;
  c[2] = c[3];
  c[3] = c[0];
}

// grid.y = geometry: wide nodes from the geometry's scratch region to their final place behind the header.
__global__ void k_batch_copy_nodes(const BatchGeom* __restrict__ geoms, const Node8* __restrict__ scratch,
                                   const uint32_t* __restrict__ counters) {
  const BatchGeom& G = geoms[blockIdx.y];
  const uint32_t nWide = counters[KF_BATCH_COUNTERS * blockIdx.y];
  const uint4* src = reinterpret_cast<const uint4*>(scratch + G.nodeOffset);
  uint4* dst = reinterpret_cast<uint4*>(G.nodesAlloc + 1);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nWide * 5u; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// grid.y = geometry: bounding sphere in the record in front of the root (see k_blas_sphere).  The header
// record was zeroed by the host.
__global__ void k_batch_spheres(const BatchGeom* __restrict__ geoms, const int* __restrict__ sceneBoxes) {
  const uint32_t g = blockIdx.y;
  const BatchGeom& G = geoms[g];
  const int* sb = sceneBoxes + 6 * g;
  const float cx = 0.5f * (orderedToFloat(sb[0]) + orderedToFloat(sb[3]));
  const float cy = 0.5f * (orderedToFloat(sb[1]) + orderedToFloat(sb[4]));
  const float cz = 0.5f * (orderedToFloat(sb[2]) + orderedToFloat(sb[5]));
  float r2 = 0.0f;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < G.nVerts; v += gridDim.x * blockDim.x) {
    const float dx = G.verts[v].pos[0] - cx, dy = G.verts[v].pos[1] - cy, dz = G.verts[v].pos[2] - cz;
    r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
  float* header = reinterpret_cast<float*>(G.nodesAlloc);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(header) + 3, __float_as_uint(r2));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    header[0] = cx;
    header[1] = cy;
    header[2] = cz;
  }
}

// One thread per triangle of the batch: the leaf-ordered traversal record (k_write_tris) and the shading
// record in primitive order (k_write_shade_tris) of its geometry.
__global__ void k_batch_write_tris(const BatchGeom* __restrict__ geoms, const uint32_t* __restrict__ primGeom,
                                   const uint32_t* __restrict__ order, uint32_t nTris) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nTris) return;
  const BatchGeom& G = geoms[primGeom[k]];  // leaf position k lies in the region of the geometry of triangle k
  const uint32_t prim = order[k] - G.triOffset;
  const float* p0 = G.verts[G.idx[3 * prim + 0]].pos;
  const float* p1 = G.verts[G.idx[3 * prim + 1]].pos;
  const float* p2 = G.verts[G.idx[3 * prim + 2]].pos;
  Tri48 t;
  t.v0x = p0[0]; t.v0y = p0[1]; t.v0z = p0[2];
  t.prim = prim;
  t.e1x = csub(p1[0], p0[0]); t.e1y = csub(p1[1], p0[1]); t.e1z = csub(p1[2], p0[2]); t.pad1 = 0.0f;
  t.e2x = csub(p2[0], p0[0]); t.e2y = csub(p2[1], p0[1]); t.e2z = csub(p2[2], p0[2]); t.pad2 = 0.0f;
  G.tris[k - G.triOffset] = t;
}

__global__ void k_batch_write_shade_tris(const BatchGeom* __restrict__ geoms, const uint32_t* __restrict__ primGeom,
                                         uint32_t nTris) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nTris) return;
  const BatchGeom& G = geoms[primGeom[i]];
  const uint32_t prim = i - G.triOffset;
  const KfrtVertex& a = G.verts[G.idx[3 * prim + 0]];
  const KfrtVertex& b = G.verts[G.idx[3 * prim + 1]];
  const KfrtVertex& c = G.verts[G.idx[3 * prim + 2]];
  ShadeTri t;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    t.n0[k] = a.normal[k];
    t.n1[k] = b.normal[k];
    t.n2[k] = c.normal[k];
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    t.uv0[k] = a.texCoord[k];
    t.uv1[k] = b.texCoord[k];
    t.uv2[k] = c.texCoord[k];
  }
  t.matIndex = G.matIndex[prim];
  G.shade[prim] = t;
  if (G.faceNormal) {
    const float e1[3] = {b.pos[0] - a.pos[0], b.pos[1] - a.pos[1], b.pos[2] - a.pos[2]};
    const float e2[3] = {c.pos[0] - a.pos[0], c.pos[1] - a.pos[1], c.pos[2] - a.pos[2]};
    G.faceNormal[prim] = make_float4(e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                                     e1[0] * e2[1] - e1[1] * e2[0], 0.0f);
  }
}

// Is the geometry convex, i.e. does every vertex lie on or behind the plane of every triangle (normal
// e1 x e2)?  Then every triangle lies on the boundary of the convex hull of the mesh, and a ray that leaves a
// point of the surface to the front side of its triangle cannot meet the mesh again -- what lets the
// traversal stages skip the instance a bounce ray starts on (kf_trace.cuh, "skipInst").  Closed convex
// meshes with outward winding qualify, and so do planar ones and convex caps.  grid.y = geometry, the
// blocks of a geometry stride over its triangles; every thread walks all vertices.  Geometries whose
// triangles x vertices exceed KF_CONVEX_MAX_WORK, and those that come after a batch has spent
// KF_CONVEX_BATCH_WORK on the ones before them, are not examined and count as not convex: the host enters
// them with convex[g] = 0.  The tolerance,
// 1e-5 of the geometry's extent, lets the two coplanar triangles of a quad pass.
#define KF_CONVEX_MAX_WORK (uint64_t(1) << 28)
#define KF_CONVEX_BATCH_WORK (uint64_t(1) << 33)  // ~40 ms of a B200 in the worst case (no early exit: all of them convex)
__global__ void k_batch_convex(const BatchGeom* __restrict__ geoms, const int* __restrict__ sceneBoxes,
                               uint32_t* __restrict__ convex) {
  const uint32_t g = blockIdx.y;
  const BatchGeom& G = geoms[g];
  if (convex[g] == 0u) return;  // not examined: too large, or the batch's budget is spent (the host says so)
  const int* sb = sceneBoxes + 6 * g;
  const float ext = fmaxf(orderedToFloat(sb[3]) - orderedToFloat(sb[0]),
                          fmaxf(orderedToFloat(sb[4]) - orderedToFloat(sb[1]), orderedToFloat(sb[5]) - orderedToFloat(sb[2])));
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < G.nTris; t += gridDim.x * blockDim.x) {
    const float* p0 = G.verts[G.idx[3 * t + 0]].pos;
    const float* p1 = G.verts[G.idx[3 * t + 1]].pos;
    const float* p2 = G.verts[G.idx[3 * t + 2]].pos;
    const float e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
    const float e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
    const float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    const float tol = 1e-5f * ext * sqrtf(nx * nx + ny * ny + nz * nz);
    bool ok = true;
    for (uint32_t v = 0; v < G.nVerts && ok; v++) {
      const float* q = G.verts[v].pos;
      ok = nx * (q[0] - p0[0]) + ny * (q[1] - p0[1]) + nz * (q[2] - p0[2]) <= tol;
      if ((v & 255u) == 255u && convex[g] == 0u) break;  // another triangle has already decided
    }
    if (!ok) {
      convex[g] = 0u;
      return;
    }
  }
}

}  // namespace kf
