// kf_wsi.cuh -- world-space instance subtrees: the two-level structure with its bottom levels moved
// into world space, and the single-space traversal stage that walks it.
//
// In the two-level structure of kf_trace.cuh every instance visit is a phase of its own (instance entry:
// re-test, world -> object transform, bounding-sphere test, ray set-up, stack bookkeeping -- 154
// instructions for ~10 of 32 lanes), the world-space ray is parked in shared memory and restored on the
// way back, and a bottom-level structure always starts at its root.  All of that is the price of sharing
// ONE bottom-level structure between the instances of a geometry.  A B200 has 180 GB: when the instanced
// triangles of a scene fit a budget (kfrtSetInstanceSubtrees, default 4 Mi triangles = 0.6 GB of nodes and
// triangles), every instance gets its OWN subtree, built over its triangles in world space by the batched
// builder of kf_blas_batch.cuh (one pseudo-geometry per instance), and the subtree's root is written into
// the slot the instance occupies in the top-level node array.  What the traversal sees is one hierarchy of
// ordinary 8-wide nodes under the SAH-built top level: no instance phase, no ray transform on the way
// down, no parked ray, nothing to undo on the way up.
//
// Hit buffers stay bit-exact with the oracle because only the BOXES live in world space.  A leaf triangle
// is the object-space record of its geometry (the same v0 / e1 / e2 bits as the shared bottom-level
// structure holds); every node of a subtree carries its instance index, and the lane keeps the
// object-space ray of the instance it tested last (world -> object with the oracle's fused expressions,
// traceInstance()), re-deriving it when a triangle of another instance comes up.  World boxes are padded
// by 2^-14 of the largest coordinate (boxPad) -- some 80x the worst disagreement between the world-space ray and the
// object-space one -- so that no triangle the object-space test accepts is culled by a world-space box.
//
// Scenes over the budget (config 5: 10 M instanced triangles) and scenes with more than 1 024 visible
// instances keep the two-level walk.  After a transform change the subtrees are rebuilt by the next
// kfrtRender (the batched builder runs at 500 - 650 Mtris/s).
//
// (An earlier variant, one LBVH over ALL instanced triangles without a top level, lost: 19.8 node steps
// per ray against 13.5, profiles/r2_session2_experiments.txt.)
#pragma once

#include "kf_blas_batch.cuh"
#include "kf_bvh_build.cuh"
#include "kf_common.cuh"
#include "kf_trace.cuh"
#include "kf_traverse.cuh"

namespace kf {

// What the build kernels need to know of one visible instance (one pseudo-geometry of the batch).
struct WsiInst {
  uint32_t instance;    // index into the instance arrays
  uint32_t vertOffset;  // first world-space vertex
  uint32_t nodeStart;   // where nodes 1.. of its subtree go in the final node array (node 0 goes to its top-level slot)
  uint32_t flags;       // bit 1: non-opaque geometry
};

// World-space positions of every visible instance's vertices (grid.y = visible instance).
__global__ void k_wsi_world_verts(const WsiInst* __restrict__ vis, const KfrtInstance* __restrict__ insts,
                                  const BlasInfo* __restrict__ blas, KfrtVertex* __restrict__ out) {
  const WsiInst V = vis[blockIdx.y];
  const KfrtInstance& I = insts[V.instance];
  const BlasInfo& G = blas[I.geometryIndex];
  const float* m = I.transform;
  const float m00 = m[0], m10 = m[1], m20 = m[2], m01 = m[4], m11 = m[5], m21 = m[6];
  const float m02 = m[8], m12 = m[9], m22 = m[10], t0 = m[12], t1 = m[13], t2 = m[14];
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < G.nVerts; v += gridDim.x * blockDim.x) {
    const float x = G.verts[v].pos[0], y = G.verts[v].pos[1], z = G.verts[v].pos[2];
    float* p = out[V.vertOffset + v].pos;
    p[0] = ((m00 * x + m01 * y) + m02 * z) + t0;
    p[1] = ((m10 * x + m11 * y) + m12 * z) + t1;
    p[2] = ((m20 * x + m21 * y) + m22 * z) + t2;
  }
}

// The padding of k_batch_tri_boxes (boxPad: 2^-14 of the largest coordinate) already covers the
// disagreement between the two spaces; nothing to add here.

// grid.y = visible instance: its subtree from the build scratch to its final place.  Node 0 goes into the
// instance's slot of the top-level array, nodes 1.. to [nodeStart, ...); child and triangle indices become
// absolute, and every node is stamped with (instance + 1) | non-opaque << 31.
__global__ void k_wsi_place_nodes(const WsiInst* __restrict__ vis, const BatchGeom* __restrict__ geoms,
                                  const uint32_t* __restrict__ counters, const Node8* __restrict__ scratch,
                                  const uint32_t* __restrict__ slotOfInst, Node8* __restrict__ nodes) {
  const WsiInst V = vis[blockIdx.y];
  const BatchGeom& G = geoms[blockIdx.y];
  const uint32_t nWide = counters[KF_BATCH_COUNTERS * blockIdx.y];
  const uint32_t stamp = (V.instance + 1u) | ((V.flags & 2u) ? 0x80000000u : 0u);  // (top-level nodes carry 0)
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nWide; j += gridDim.x * blockDim.x) {
    Node8 nd = scratch[G.nodeOffset + j];
    nd.childBase = V.nodeStart + nd.childBase - 1u;  // relative index c >= 1 -> nodeStart + c - 1
    nd.primBase += G.triOffset;
    nd.reserved = stamp;
    nodes[j == 0 ? slotOfInst[V.instance] : V.nodeStart + j - 1u] = nd;
  }
}

// Slots of instances without triangles (hidden or empty geometries): a node without children.
__global__ void k_wsi_empty_slots(const uint32_t* __restrict__ visibleRank, uint32_t nInst,
                                  const uint32_t* __restrict__ slotOfInst, Node8* __restrict__ nodes) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nInst || visibleRank[i] != 0xffffffffu) return;
  Node8 nd;
  memset(&nd, 0, sizeof(nd));
  nodes[slotOfInst[i]] = nd;
}

// One thread per leaf position of the batch: the OBJECT-space record of the triangle that landed there
// (the bits k_batch_write_tris writes into the geometry's own structure).
__global__ void k_wsi_write_tris(const WsiInst* __restrict__ vis, const BatchGeom* __restrict__ geoms,
                                 const KfrtInstance* __restrict__ insts, const BlasInfo* __restrict__ blas,
                                 const uint32_t* __restrict__ primGeom, const uint32_t* __restrict__ order,
                                 uint32_t nTris, Tri48* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nTris) return;
  const uint32_t g = primGeom[k];  // leaf position k lies in the region of the pseudo-geometry of triangle k
  const uint32_t prim = order[k] - geoms[g].triOffset;
  const BlasInfo& G = blas[insts[vis[g].instance].geometryIndex];
  const float* p0 = G.verts[G.idx[3 * prim + 0]].pos;
  const float* p1 = G.verts[G.idx[3 * prim + 1]].pos;
  const float* p2 = G.verts[G.idx[3 * prim + 2]].pos;
  Tri48 t;
  t.v0x = p0[0]; t.v0y = p0[1]; t.v0z = p0[2];
  t.prim = prim;
  t.e1x = csub(p1[0], p0[0]); t.e1y = csub(p1[1], p0[1]); t.e1z = csub(p1[2], p0[2]); t.pad1 = 0.0f;
  t.e2x = csub(p2[0], p0[0]); t.e2y = csub(p2[1], p0[1]); t.e2z = csub(p2[2], p0[2]); t.pad2 = 0.0f;
  out[k] = t;
}

// ---------------------------------------------------------------------------------------------
// Traversal: the persistent-lane loop of k_wf_trace (kf_trace.cuh) without its instance phase --
// [node] [triangle] [pop] per iteration, same refill, same node step, same triangle arithmetic, same
// tie-break, so the results are those of the two-level walk bit for bit.
template <bool ANY, bool DETAIL>
__global__ void __launch_bounds__(128, ANY ? 8 : 7) k_wf_trace_wsi(TraceArgs a) {
  const uint32_t count = *a.count;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (a.clear0) *a.clear0 = 0;
    if (a.clear1) *a.clear1 = 0;
    if (a.clear2) *a.clear2 = 0;
    atomicAdd(a.counters + a.rayCounter, (unsigned long long)count);
  }
  const SceneDev& sc = a.sc;
  const Node8* __restrict__ nodes = sc.wsiNodes;
  const Tri48* __restrict__ tris = sc.wsiTris;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t laneLt = (1u << lane) - 1u;
  const uint32_t tid = threadIdx.x;
  const float tmin = 0.001f;

  TravCounters tc{0, 0, 0, 0, 0};
  uint2 stack[KF_STACK];
  int sp = 0;
  bool active = false, exhausted = false;
  uint32_t qpos = 0;
  RaySetup r = setupRay(mk3(0.0f), mk3(1.0f));  // the world-space ray, for the whole walk
  V3 oo = mk3(0.0f), od = mk3(1.0f);            // the ray in the object space of instance `curInst`
  uint32_t curInst = 0xffffffffu;               // (instance + 1) | non-opaque << 31, as stamped on the nodes
  uint32_t skipStamp = 0xffffffffu;
  __shared__ MaskTables sMask;
  fillMaskTables(sMask);
  __syncthreads();
  Hit hit;
  hit.t = 0.0f; hit.u = hit.v = 0.0f; hit.inst = hit.prim = -1; hit.front = 0;
  uint2 ng = make_uint2(0u, 0u), tg = make_uint2(0u, 0u);
  uint32_t tgInst = 0u;  // stamp of the node the pending triangles belong to
  bool finished = false;
  __shared__ uint2 sStack[KF_STACK_SHARED][128];
  uint2 top = make_uint2(0u, 0u);
  auto push = [&](uint2 e) {
    if (sp < KF_STACK_SHARED) {
      sStack[sp++][tid] = e;
      top = e;
    } else if (sp < KF_STACK_SHARED + KF_STACK) {
      stack[sp++ - KF_STACK_SHARED] = e;
      top = e;
    }
  };

  for (;;) {
    // ---- refill: lanes without a ray take consecutive queue positions -------------------------
    const uint32_t idle = __ballot_sync(0xffffffffu, !active);
    if (idle) {
      if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= KF_REFILL_IDLE)) {
        const uint32_t want = uint32_t(__popc(idle));
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.fetch, want);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!active) {
          const uint32_t qi = base + uint32_t(__popc(idle & laneLt));
          if (qi < count) {
            qpos = qi;
            const uint32_t slot = a.queue[qi];
            const float4 o4 = a.rayO[slot], d4 = a.rayD[slot];
            hit.t = ANY ? d4.w : 10000.0f;
            hit.u = hit.v = 0.0f;
            hit.inst = -1;
            hit.prim = -1;
            hit.front = 0;
            r = setupRay(mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z));
            {
              const uint32_t ob = __float_as_uint(o4.w);  // the instance the ray must not enter (kf_trace.cuh, sSkip), as a stamp
              skipStamp = ((ob >> (ANY ? 31 : 30)) & 1u) ? (ob & 0x1fffffffu) + 1u : 0xffffffffu;
            }
            curInst = 0xffffffffu;
            sp = 0;
            ng = make_uint2(0u, nodes ? 0x80000000u : 0u);  // the root: "child 7 ^ octinv of a virtual parent"
            tg = make_uint2(0u, 0u);
            active = true;
          }
        }
        if (base + want >= count) exhausted = true;
      }
      if (exhausted && __ballot_sync(0xffffffffu, active) == 0u) break;
    }
    finished = false;

    // ---- node phase: lanes without pending triangles take one node step -------------------------
    if (active && !(tg.y & 0xffffu) && (ng.y & 0xff000000u)) {
      const uint32_t hits = ng.y;
      const int p = 31 - __clz(hits);
      const uint32_t cslot = uint32_t(p - 24) ^ r.octinv;
      ng.y &= ~(1u << p);
      if (ng.y & 0xff000000u) push(ng);
      const uint32_t rel = __popc(hits & 0xffu & ((1u << cslot) - 1u));
      const uint4* q = reinterpret_cast<const uint4*>(nodes + (ng.x + rel));
      const uint4 n0 = __ldg(q + 0), n1 = __ldg(q + 1), n2 = __ldg(q + 2), n3 = __ldg(q + 3), n4 = __ldg(q + 4);
      uint32_t childBase, primBase, imask, triMask;
      const uint32_t miss0 = intersectNodeWords(n0, n1, n2, n3, n4, r, tmin, hit.t, childBase, primBase, imask, triMask);
      if (DETAIL) tc.nodes++;
      uint32_t miss = miss0;
      if ((n1.w & 0x7fffffffu) == skipStamp) miss = 0xffu;  // a node of the subtree the ray starts on and leaves
      const uint32_t inner = sMask.perm[r.octinv][imask & ~miss];
      ng = make_uint2(childBase, (inner << 24) | imask);
      tg = make_uint2(primBase, (uint32_t(sMask.expand[miss]) | 0xffff0000u) & triMask);
      tgInst = n1.w;
    }
    // ---- triangle phase: one leaf triangle per lane --------------------------------------------------
    if (active && (tg.y & 0xffffu)) {
      const int b = __ffs(tg.y) - 1;
      tg.y &= tg.y - 1;
      const float4* tp = reinterpret_cast<const float4*>(tris + (tg.x + __popc((tg.y >> 16) & ((1u << b) - 1u))));
      const float4 v0 = __ldg(tp + 0), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
      if (DETAIL) tc.tris++;
      const uint32_t iw = tgInst;
      if (iw != curInst) {
        // world -> object (fused arithmetic, the same expressions as oracle traceInstance())
        const float4* ip = reinterpret_cast<const float4*>(sc.inst + ((iw & 0x7fffffffu) - 1u));
        const float4 r0 = __ldg(ip + 0), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
        if (DETAIL) tc.insts++;
        oo.x = cfma(r0.z, r.oz, cfma(r0.y, r.oy, cfma(r0.x, r.ox, r0.w)));
        oo.y = cfma(r1.z, r.oz, cfma(r1.y, r.oy, cfma(r1.x, r.ox, r1.w)));
        oo.z = cfma(r2.z, r.oz, cfma(r2.y, r.oy, cfma(r2.x, r.ox, r2.w)));
        od.x = cfma(r0.z, r.dz, cfma(r0.y, r.dy, cmul(r0.x, r.dx)));
        od.y = cfma(r1.z, r.dz, cfma(r1.y, r.dy, cmul(r1.x, r.dx)));
        od.z = cfma(r2.z, r.dz, cfma(r2.y, r.dy, cmul(r2.x, r.dx)));
        curInst = iw;
      }
      const int32_t inst = int32_t((iw & 0x7fffffffu) - 1u);
      // Moller-Trumbore, fused arithmetic, same expressions as oracle intersectTri()
      const V3 E1 = mk3(e1.x, e1.y, e1.z), E2 = mk3(e2.x, e2.y, e2.z);
      const V3 pv = fcross(od, E2);
      const float det = fdot(E1, pv);
      const float inv = __frcp_rn(det);  // correctly rounded 1 / det, the oracle's `1.0f / det`
      const V3 tv = csub3(oo, mk3(v0.x, v0.y, v0.z));
      const float u = cmul(fdot(tv, pv), inv);
      const V3 qv = fcross(tv, E1);
      const float v = cmul(fdot(od, qv), inv);
      const float t = cmul(fdot(E2, qv), inv);
      const int32_t prim = int32_t(__float_as_uint(v0.w));
      bool ok = det != 0.0f && (u >= 0.0f && u <= 1.0f) && (v >= 0.0f && cadd(u, v) <= 1.0f) && t > tmin;
      ok = ok && (t < hit.t || (t == hit.t && hit.inst >= 0 &&
                                (inst < hit.inst || (inst == hit.inst && prim < hit.prim))));
      if (!ANY && ok && (iw >> 31)) {
        const uint32_t g = sc.instSsbo[inst].geometryIndex;
        const uint32_t mi = __ldg(sc.geoms[g].matIndex + prim);
        const float alpha = sc.mats[mi].alpha;
        const uint32_t seed = __float_as_uint(a.seedSrc[a.queue[qpos]].w);
        if (alpha == 0.0f || anyHitRnd(seed, uint32_t(inst), uint32_t(prim)) > alpha) ok = false;
      }
      if (ok) {
        hit.t = t;
        hit.u = u;
        hit.v = v;
        hit.inst = inst;
        hit.prim = prim;
        hit.front = det > 0.0f ? 1u : 0u;
        if (ANY) finished = true;
      }
    }
    // ---- pop: lanes with nothing left in hand take the next group from their stack ----------------
    if (active && !finished && !(tg.y & 0xffffu) && !(ng.y & 0xff000000u)) {
      if (sp == 0) {
        finished = true;
      } else {
        ng = top;
        --sp;
        if (sp > KF_STACK_SHARED) top = stack[sp - 1 - KF_STACK_SHARED];
        else if (sp > 0) top = sStack[sp - 1][tid];
      }
    }

    if (finished) {
      if (ANY) {
        a.hitB[qpos] = hit.inst >= 0 ? 1 : 0;
      } else {
        a.hitA[qpos] = make_float4(hit.t, hit.u, hit.v, __int_as_float(hit.prim));
        a.hitB[qpos] = hit.inst < 0 ? -1 : int(uint32_t(hit.inst) | (hit.front << 31));
      }
      active = false;
    }
  }

  if (DETAIL) {
    atomicAdd(a.counters + a.detailBase + 0, (unsigned long long)tc.nodes);
    atomicAdd(a.counters + a.detailBase + 1, (unsigned long long)tc.tris);
    atomicAdd(a.counters + a.detailBase + 2, (unsigned long long)tc.insts);
    atomicAdd(a.counters + 13, (unsigned long long)tc.insts);
  }
}

}  // namespace kf
