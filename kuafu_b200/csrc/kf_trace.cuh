// kf_trace.cuh -- persistent-warp ray traversal stage of the wavefront scheduler (replaces
// traceRayEXT and the driver/RT-core traversal behind it: reference PathTrace.rgen:97-107 closest
// hit, PathTrace.rchit:186-197 occlusion).
//
// One lane walks one ray through the two-level compressed 8-wide BVH, but lanes are *persistent*:
// a lane whose ray has terminated does not wait for the slowest ray of its warp, it takes the next
// ray from the stage queue (one warp-aggregated atomic per refill).  With incoherent bounce rays the
// per-ray traversal length varies by an order of magnitude, so this is what keeps lanes busy.
//
// Every loop iteration runs the phases below warp-wide, each lane taking part in those that match
// its state (so that lanes doing the same kind of work do it in the same instructions):
//   [instance]  (top level) the nearest pending child is an InstNode: re-test its world box against
//               the current closest hit, then world -> object ray transform and enter its BLAS root
//   [node]      take the nearest pending child node, fetch its 80 bytes (5 x 16 B loads), slab-test
//               its 8 quantised child boxes
//   [pop]       next group from the stack; a sentinel entry returns to the top level
//   [triangle]  (bottom level) Moller-Trumbore on one leaf triangle of the current group
//
// Results do not depend on traversal order: boxes are conservative and equal-t ties resolve to the
// lowest (instance, primitive) pair (oracle deviation D3), so hit buffers stay bit-exact with the
// CPU oracle whatever the scheduling.
#pragma once

#include <cuda_fp16.h>

#include "kf_common.cuh"
#include "kf_traverse.cuh"

namespace kf {

#ifndef KF_STACK_SHARED
#define KF_STACK_SHARED 8  // stack entries per lane kept in shared memory (8 KB per block)
#endif
// Both are compile-time constants (as kernel arguments they cost a constant load and a compare per
// iteration: 0.4 % of the stage); re-measured in round 2 after the node test got cheaper: 6 / 10 / 12 idle
// lanes and periods 1 / 3 all lose 0.2 - 3 %.
#define KF_INST_PERIOD 2  // measured on config 3: 1 -> 2 takes 2.8 % off the closest-hit stage, 4 loses it again
#ifndef KF_REFILL_IDLE
#define KF_REFILL_IDLE 8  // refill the warp once this many lanes are without a ray
#endif

struct TraceArgs {
  SceneDev sc;
  const uint32_t* queue;     // slots to trace
  const uint32_t* count;     // number of queued slots (device)
  uint32_t* fetch;           // next queue position to hand out (zeroed by an earlier stage)
  const float4* rayO;        // origin.xyz per slot
  const float4* rayD;        // direction.xyz (+ tmax in .w for occlusion rays) per slot
  const float4* seedSrc;     // stateW: .w = ray seed bits (closest hit, non-opaque geometry only)
  // results are written in QUEUE order (entry i of the queue -> hitA[i], hitB[i]), so that the stage that
  // consumes them reads them as a stream next to the queue itself instead of gathering them by slot
  float4* hitA;              // closest hit: t, u, v, prim bits
  int* hitB;                 // closest hit: inst | front << 31, -1 on miss; occlusion: 1 occluded / 0
  uint32_t* clear0;          // counters this stage resets for later stages (may be null)
  uint32_t* clear1;
  uint32_t* clear2;
  unsigned long long* counters;
  int rayCounter;            // counters[] index that receives the number of rays of this stage
  int detailBase;            // counters[] index of (nodes, tris, insts) for detail accounting
};

// Closest hit (ANY == false) or first hit (ANY == true; TerminateOnFirstHit | Opaque |
// SkipClosestHitShader) for every ray of the queue.
template <bool ANY, bool DETAIL>
// Occlusion rays carry no (u, v, prim) and fit 64 registers (8 blocks / SM) without spilling; the
// closest-hit variant runs best at 72 registers (7 blocks / SM).
__global__ void __launch_bounds__(128, ANY ? 8 : 7) k_wf_trace(TraceArgs a) {
  const uint32_t count = *a.count;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (a.clear0) *a.clear0 = 0;
    if (a.clear1) *a.clear1 = 0;
    if (a.clear2) *a.clear2 = 0;
    atomicAdd(a.counters + a.rayCounter, (unsigned long long)count);
  }
  const SceneDev& sc = a.sc;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t laneLt = (1u << lane) - 1u;
  const float tmin = 0.001f;

  TravCounters tc{0, 0, 0, 0, 0};
  uint2 stack[KF_STACK];
  int sp = 0;
  bool active = false, exhausted = false;
  uint32_t qpos = 0;  // queue position of the lane's ray
  RaySetup r = setupRay(mk3(0.0f), mk3(1.0f));
  // The world-space ray setup waits in shared memory (written once, when the ray is fetched) while
  // the lane is inside a bottom-level structure (r then holds the object-space ray): ten registers
  // less per lane.
  __shared__ float sWorld[10][128];
  __shared__ MaskTables sMask;
  // The instance the lane's ray must not enter, or ~0: a bounce ray that leaves a convex geometry to the
  // front side of the triangle it starts on cannot meet that geometry again (k_batch_convex), and the shade
  // stage says so in the spare word of the ray's origin.  Without it every such ray walks the bottom
  // level of its own surface around its origin, finds the triangle it stands on and its neighbours, and
  // rejects them all.
  __shared__ uint32_t sSkip[128];
  fillMaskTables(sMask);
  __syncthreads();
  const uint32_t tid = threadIdx.x;
  Hit hit;
  hit.t = 0.0f; hit.u = hit.v = 0.0f; hit.inst = hit.prim = -1; hit.front = 0;
  const Node8* nodes = sc.tlasNodes;
  const Tri48* tris = nullptr;
  bool inBlas = false, nonOpaque = false;
  int32_t curInst = -1;
  uint2 ng = make_uint2(0u, 0u), tg = make_uint2(0u, 0u);

  bool finished = false;
  int iter = 0;
  // The first KF_STACK_SHARED entries of the stack live in shared memory, deeper ones (rare) in local
  // memory, and the top entry is mirrored in registers: a pop hands out the register copy at once
  // and reads the entry below it.  (ncu, all-local version: the reload of the entry below was the
  // largest single long-scoreboard stall of the kernel, 6.4 % of all samples, and the stack lines
  // competed with the BVH for L1.)
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
  // Next node group from the lane's stack.  spBlas is the stack height at instance entry: popping
  // at that height means the bottom level is exhausted and the ray returns to world space.  An empty
  // stack ends the ray.
  int spBlas = 0;
  auto popGroup = [&]() {
    if (inBlas && sp == spBlas) {
      r.ox = sWorld[0][tid]; r.oy = sWorld[1][tid]; r.oz = sWorld[2][tid];
      r.dx = sWorld[3][tid]; r.dy = sWorld[4][tid]; r.dz = sWorld[5][tid];
      r.ix = sWorld[6][tid]; r.iy = sWorld[7][tid]; r.iz = sWorld[8][tid];
      r.octinv = __float_as_uint(sWorld[9][tid]);
      nodes = sc.tlasNodes;
      inBlas = false;
    }
    if (sp == 0) {
      finished = true;
      return;
    }
    ng = top;
    --sp;
    if (sp > KF_STACK_SHARED) top = stack[sp - 1 - KF_STACK_SHARED];
    else if (sp > 0) top = sStack[sp - 1][tid];
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
            const V3 o = mk3(o4.x, o4.y, o4.z);
            const V3 d = mk3(d4.x, d4.y, d4.z);
            hit.t = ANY ? d4.w : 10000.0f;
            hit.u = hit.v = 0.0f;
            hit.inst = -1;
            hit.prim = -1;
            hit.front = 0;
            {
              const uint32_t ob = __float_as_uint(o4.w);
              sSkip[tid] = ((ob >> (ANY ? 31 : 30)) & 1u) ? (ob & 0x1fffffffu) : 0xffffffffu;
            }
            r = setupRay(o, d);
            // the world-space setup is parked once per ray; popGroup() restores it on every return
            sWorld[0][tid] = r.ox; sWorld[1][tid] = r.oy; sWorld[2][tid] = r.oz;
            sWorld[3][tid] = r.dx; sWorld[4][tid] = r.dy; sWorld[5][tid] = r.dz;
            sWorld[6][tid] = r.ix; sWorld[7][tid] = r.iy; sWorld[8][tid] = r.iz;
            sWorld[9][tid] = __uint_as_float(r.octinv);
            nodes = sc.tlasNodes;
            inBlas = false;
            nonOpaque = false;
            curInst = -1;
            sp = 0;
            // the root is "child 7 ^ octinv of a virtual parent": a real node, present mask empty
            ng = make_uint2(0u, sc.tlasNodes ? (0x80000000u | 0x0000ff00u) : 0u);
            tg = make_uint2(0u, 0u);
            active = true;
          }
        }
        if (base + want >= count) exhausted = true;
      }
      if (exhausted && __ballot_sync(0xffffffffu, active) == 0u) break;
    }
    finished = false;

    // ---- instance phase (top level): the nearest pending child is an InstNode ------------------
    // Few lanes want it in any one iteration, yet the whole warp pays its ~130 instructions; taking
    // it only every instPeriod-th iteration lets the requests pile up (a waiting lane idles).
    static_assert(KF_INST_PERIOD == 2, "the phase counter below is a parity bit");
    iter ^= 1;
    if (iter == 0 && active && !inBlas && (ng.y & 0xff000000u)) {
      const uint32_t hits = ng.y;
      const int p = 31 - __clz(hits);
      const uint32_t cslot = uint32_t(p - 24) ^ r.octinv;
      if (!((hits >> (8 + cslot)) & 1u)) {
        ng.y &= ~(1u << p);
        const uint32_t rel = __popc(hits & 0xffu & ((1u << cslot) - 1u));
        const uint4* ip = reinterpret_cast<const uint4*>(nodes + (ng.x + rel));  // 32-bit index sum: one wide multiply-add
        const uint4 w4 = __ldg(ip + 4);
        if (DETAIL) tc.insts++;
        // re-test the instance's own world box against what is now the closest hit
        const float2 lxy = __half22float2(*reinterpret_cast<const __half2*>(&w4.x));
        const float2 lzhx = __half22float2(*reinterpret_cast<const __half2*>(&w4.y));
        const float2 hyz = __half22float2(*reinterpret_cast<const __half2*>(&w4.z));
        const float ax0 = (lxy.x - r.ox) * r.ix, ax1 = (lzhx.y - r.ox) * r.ix;
        const float ay0 = (lxy.y - r.oy) * r.iy, ay1 = (hyz.x - r.oy) * r.iy;
        const float az0 = (lzhx.x - r.oz) * r.iz, az1 = (hyz.y - r.oz) * r.iz;
        const float t0 = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), tmin));
        const float t1 = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fminf(fmaxf(az0, az1), hit.t));
        const ulonglong2 ptrs = __ldg(reinterpret_cast<const ulonglong2*>(ip + 3));
        if (t0 <= t1 && ptrs.x != 0ull && w4.w != sSkip[tid]) {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(ip) + 0);
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(ip) + 1);
          const float4 r2 = __ldg(reinterpret_cast<const float4*>(ip) + 2);
          // bounding sphere of the geometry, in the record in front of its root
          const float4 sph = __ldg(reinterpret_cast<const float4*>(ptrs.x) - 5);
          // world -> object (fused arithmetic, the same expressions as oracle traceInstance())
          const V3 o = mk3(r.ox, r.oy, r.oz), d = mk3(r.dx, r.dy, r.dz);
          V3 oo, od;
          oo.x = cfma(r0.z, o.z, cfma(r0.y, o.y, cfma(r0.x, o.x, r0.w)));
          oo.y = cfma(r1.z, o.z, cfma(r1.y, o.y, cfma(r1.x, o.x, r1.w)));
          oo.z = cfma(r2.z, o.z, cfma(r2.y, o.y, cfma(r2.x, o.x, r2.w)));
          od.x = cfma(r0.z, d.z, cfma(r0.y, d.y, cmul(r0.x, d.x)));
          od.y = cfma(r1.z, d.z, cfma(r1.y, d.y, cmul(r1.x, d.x)));
          od.z = cfma(r2.z, d.z, cfma(r2.y, d.y, cmul(r2.x, d.x)));
          // Object-space sphere test, conservative: the ray is turned away only when its line passes
          // the centre at more than the radius plus a margin that outweighs the rounding of this
          // computation (relative 1e-4 on r^2, 2e-5 on |origin - centre|^2: the closest-approach vector
          // is a difference of two vectors of that size), or when the sphere lies behind the origin.
          const float cx = oo.x - sph.x, cy = oo.y - sph.y, cz = oo.z - sph.z;
          const float oc2 = cx * cx + cy * cy + cz * cz;
          const float rp = sph.w * 1.0001f + 2e-5f * oc2 + 1e-30f;
          bool enter = true;
          if (oc2 > rp) {
            const float A = od.x * od.x + od.y * od.y + od.z * od.z;
            const float B = cx * od.x + cy * od.y + cz * od.z;
            const float s = B * __frcp_rn_approx(A);
            const float wx = cx - s * od.x, wy = cy - s * od.y, wz = cz - s * od.z;
            enter = !(B > 0.0f) && !(wx * wx + wy * wy + wz * wz > rp);
          }
          if (enter) {
            if (DETAIL) tc.entries++;
            // what is left of this top-level node waits on the stack, below the bottom-level entries
            if (ng.y & 0xff000000u) push(ng);
            spBlas = sp;
            r = setupRay(oo, od);
            nodes = reinterpret_cast<const Node8*>(ptrs.x);
            nonOpaque = (ptrs.y & 1ull) != 0;
            tris = reinterpret_cast<const Tri48*>(ptrs.y & ~1ull);
            curInst = int32_t(w4.w);
            inBlas = true;
            ng = make_uint2(0u, 0x80000000u);  // bottom-level root; steps in the node phase below
          }
        }
      }
    }

    // ---- node phase: lanes without pending triangles take one node step -------------------------
    if (active && !(tg.y & 0xffffu) && (ng.y & 0xff000000u)) {
      const uint32_t hits = ng.y;
      const int p = 31 - __clz(hits);
      const uint32_t cslot = uint32_t(p - 24) ^ r.octinv;
      if (inBlas || ((hits >> (8 + cslot)) & 1u)) {
        ng.y &= ~(1u << p);
        if (ng.y & 0xff000000u) push(ng);
        const uint32_t rel = __popc(hits & 0xffu & ((1u << cslot) - 1u));
        uint32_t childBase, primBase, imask, triMask;
        const uint32_t miss = intersectNode(nodes + (ng.x + rel), r, tmin, hit.t, childBase, primBase, imask, triMask);
        if (DETAIL) tc.nodes++;
        if (DETAIL && !inBlas) tc.tlasNodes++;
        const uint32_t inner = sMask.perm[r.octinv][imask & ~miss];
        // top level: imask = present children, primBase = which of them are real nodes (triMask = 0)
        ng = make_uint2(childBase, (inner << 24) | imask | (inBlas ? 0u : (primBase & 0xffu) << 8));
        tg = make_uint2(primBase, (uint32_t(sMask.expand[miss]) | 0xffff0000u) & triMask);
      }
    }
    // ---- triangle phase (bottom level): one leaf triangle per lane -------------------------------
    if (active && !finished && (tg.y & 0xffffu)) {
      const int b = __ffs(tg.y) - 1;
      tg.y &= tg.y - 1;
      const float4* tp = reinterpret_cast<const float4*>(tris + (tg.x + __popc((tg.y >> 16) & ((1u << b) - 1u))));
      const float4 v0 = __ldg(tp + 0), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
      if (DETAIL) tc.tris++;
      // Moller-Trumbore, fused arithmetic, same expressions as oracle intersectTri()
      const V3 dd = mk3(r.dx, r.dy, r.dz);
      const V3 E1 = mk3(e1.x, e1.y, e1.z), E2 = mk3(e2.x, e2.y, e2.z);
      const V3 pv = fcross(dd, E2);
      const float det = fdot(E1, pv);
      const float inv = __frcp_rn(det);  // correctly rounded 1 / det, the oracle's `1.0f / det`
      const V3 tv = csub3(mk3(r.ox, r.oy, r.oz), mk3(v0.x, v0.y, v0.z));
      const float u = cmul(fdot(tv, pv), inv);
      const V3 qv = fcross(tv, E1);
      const float v = cmul(fdot(dd, qv), inv);
      const float t = cmul(fdot(E2, qv), inv);
      const int32_t prim = int32_t(__float_as_uint(v0.w));
      bool ok = det != 0.0f && (u >= 0.0f && u <= 1.0f) && (v >= 0.0f && cadd(u, v) <= 1.0f) && t > tmin;
      ok = ok && (t < hit.t || (t == hit.t && hit.inst >= 0 &&
                                (curInst < hit.inst || (curInst == hit.inst && prim < hit.prim))));
      if (!ANY && ok && nonOpaque) {
        const uint32_t g = sc.instSsbo[curInst].geometryIndex;
        const uint32_t mi = __ldg(sc.geoms[g].matIndex + prim);
        const float alpha = sc.mats[mi].alpha;
        const uint32_t seed = __float_as_uint(a.seedSrc[a.queue[qpos]].w);
        if (alpha == 0.0f || anyHitRnd(seed, uint32_t(curInst), uint32_t(prim)) > alpha) ok = false;
      }
      if (ok) {
        hit.t = t;
        hit.u = u;
        hit.v = v;
        hit.inst = curInst;
        hit.prim = prim;
        hit.front = det > 0.0f ? 1u : 0u;
        if (ANY) finished = true;
      }
    }
    // ---- pop: lanes with nothing left in hand take the next group from their stack ----------------
    if (active && !finished && !(tg.y & 0xffffu) && !(ng.y & 0xff000000u)) popGroup();

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
    atomicAdd(a.counters + 12, (unsigned long long)tc.tlasNodes);
    atomicAdd(a.counters + 13, (unsigned long long)tc.entries);
  }
}

}  // namespace kf
