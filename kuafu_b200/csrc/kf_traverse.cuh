// kf_traverse.cuh -- two-level traversal of the compressed 8-wide BVH (replaces traceRayEXT and the
// driver/RT-core traversal behind it: reference PathTrace.rgen:97-107, PathTrace.rchit:186-197).
//
// One thread walks one ray.  State is a small stack of 8-byte "groups":
//   node group      (childBase, hits<<24 | imask)   -- internal children still to visit, in
//                                                      octant priority order (highest bit first)
//   primitive group (primBase, 24-bit mask)         -- triangles (BLAS) or instance entries (TLAS)
//   sentinel        (x, 0)                          -- marks the return from a BLAS to the TLAS
#pragma once

#include "kf_common.cuh"

namespace kf {

#define KF_STACK 48

struct TravCounters {
  uint32_t nodes, tris, insts;
};

struct RaySetup {
  float ox, oy, oz;
  float dx, dy, dz;
  float ix, iy, iz;  // safe reciprocal direction (box tests only)
  uint32_t octinv;   // bit k set when direction k is >= 0
};

// Reciprocals feed the (conservative) box tests only, so the approximate MUFU.RCP is enough.
KF_D float boxRcp(float x) {
  const float eps = 1e-20f;
  return __fdividef(1.0f, fabsf(x) > eps ? x : copysignf(eps, x));
}
KF_D RaySetup setupRay(V3 o, V3 d) {
  RaySetup r;
  r.ox = o.x; r.oy = o.y; r.oz = o.z;
  r.dx = d.x; r.dy = d.y; r.dz = d.z;
  r.ix = boxRcp(d.x);
  r.iy = boxRcp(d.y);
  r.iz = boxRcp(d.z);
  r.octinv = (r.ix >= 0.0f ? 1u : 0u) | (r.iy >= 0.0f ? 2u : 0u) | (r.iz >= 0.0f ? 4u : 0u);
  return r;
}

// Quantised plane byte j of word w as the float 1 + q * 2^-15: one PRMT drops the byte into the
// second mantissa byte of 1.0f, so the dequantisation never touches the (quarter-rate) int->float
// conversion pipe.  t = f * A + B with A = 2^15 * scale * idir, B = (origin term) - A.
KF_D float planeFloat(uint32_t w, int j) {
  return __uint_as_float(__byte_perm(w, 0x3F800000u, 0x7604u | (uint32_t(j) << 4)));
}

// Intersects the 8 quantised child boxes of `node`; returns the CWBVH hit mask:
// bits 24..31 internal children by traversal priority, bits 0..23 leaf primitives.
KF_D uint32_t intersectNode(const Node8* __restrict__ node, const RaySetup& r, float tmin, float tmax,
                            uint32_t& childBase, uint32_t& primBase, uint32_t& imask) {
  const uint4* q = reinterpret_cast<const uint4*>(node);
  const uint4 n0 = __ldg(q + 0);
  const uint4 n1 = __ldg(q + 1);
  const uint4 n2 = __ldg(q + 2);
  const uint4 n3 = __ldg(q + 3);
  const uint4 n4 = __ldg(q + 4);
  childBase = n1.x;
  primBase = n1.y;
  imask = n0.w >> 24;
  // per-axis grid step 2^(e-127), pre-multiplied by 2^15 (see planeFloat)
  const float sx = __uint_as_float(((n0.w & 0xffu) + 15u) << 23);
  const float sy = __uint_as_float((((n0.w >> 8) & 0xffu) + 15u) << 23);
  const float sz = __uint_as_float((((n0.w >> 16) & 0xffu) + 15u) << 23);
  const float ax = sx * r.ix, ay = sy * r.iy, az = sz * r.iz;
  const float bx = (__uint_as_float(n0.x) - r.ox) * r.ix - ax;
  const float by = (__uint_as_float(n0.y) - r.oy) * r.iy - ay;
  const float bz = (__uint_as_float(n0.z) - r.oz) * r.iz - az;
  // near/far plane words per axis (two words = 8 children)
  const bool nx = r.ix < 0.0f, ny = r.iy < 0.0f, nz = r.iz < 0.0f;
  const uint32_t lox[2] = {n2.x, n2.y}, loy[2] = {n2.z, n2.w}, loz[2] = {n3.x, n3.y};
  const uint32_t hix[2] = {n3.z, n3.w}, hiy[2] = {n4.x, n4.y}, hiz[2] = {n4.z, n4.w};
  const uint32_t meta[2] = {n1.z, n1.w};
  uint32_t hitmask = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t nearx = nx ? hix[h] : lox[h], farx = nx ? lox[h] : hix[h];
    const uint32_t neary = ny ? hiy[h] : loy[h], fary = ny ? loy[h] : hiy[h];
    const uint32_t nearz = nz ? hiz[h] : loz[h], farz = nz ? loz[h] : hiz[h];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t m = (meta[h] >> (8 * j)) & 0xffu;
      const float t0x = fmaf(planeFloat(nearx, j), ax, bx);
      const float t0y = fmaf(planeFloat(neary, j), ay, by);
      const float t0z = fmaf(planeFloat(nearz, j), az, bz);
      const float t1x = fmaf(planeFloat(farx, j), ax, bx);
      const float t1y = fmaf(planeFloat(fary, j), ay, by);
      const float t1z = fmaf(planeFloat(farz, j), az, bz);
      const float t0 = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
      const float t1 = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
      if (m != 0 && t0 <= t1) {
        const int slot = h * 4 + j;
        if ((imask >> slot) & 1u)
          hitmask |= 1u << (24 + (slot ^ r.octinv));
        else
          hitmask |= (m >> 5) << (m & 31u);
      }
    }
  }
  return hitmask;
}

// Order-independent surrogate of the stochastic any-hit draw (reference PathTrace.rahit:30-48;
// declared deviation D5, see oracle/kf_oracle.cpp).
KF_D float anyHitRnd(uint32_t seed, uint32_t inst, uint32_t prim) {
  uint32_t h = tea(seed ^ (prim * 0x9e3779b9u), inst);
  return float(h & 0x00FFFFFFu) * (1.0f / 16777216.0f);
}

// Closest hit (ANY == false; any-hit alpha test on non-opaque geometry) or first hit (ANY == true;
// gl_RayFlagsTerminateOnFirstHit | Opaque | SkipClosestHitShader) in the open interval (tmin, tmax).
template <bool ANY, bool DETAIL>
KF_D bool traverse(const SceneDev& sc, V3 o, V3 d, float tmin, float tmax, uint32_t seed, Hit& hit,
                   TravCounters& tc) {
  uint2 stack[KF_STACK];
  int sp = 0;
  hit.t = tmax;
  hit.u = hit.v = 0.0f;
  hit.inst = -1;
  hit.prim = -1;
  hit.front = 0;
  if (sc.tlasNodes == nullptr) return false;

  RaySetup r = setupRay(o, d);
  const Node8* nodes = sc.tlasNodes;
  const Tri48* tris = nullptr;
  bool inBlas = false, nonOpaque = false;
  int32_t curInst = -1;
  uint2 ng = make_uint2(0u, 0x80000000u);
  uint2 tg = make_uint2(0u, 0u);

  for (;;) {
    if (ng.y & 0xff000000u) {
      const uint32_t hits = ng.y;
      const int p = 31 - __clz(hits);
      ng.y &= ~(1u << p);
      if (ng.y & 0xff000000u) {
        if (sp < KF_STACK) stack[sp++] = ng;
      }
      const uint32_t slot = uint32_t(p - 24) ^ r.octinv;
      const uint32_t rel = __popc(hits & 0xffu & ((1u << slot) - 1u));
      uint32_t childBase, primBase, imask;
      const uint32_t hm = intersectNode(nodes + ng.x + rel, r, tmin, hit.t, childBase, primBase, imask);
      if (DETAIL) tc.nodes++;
      ng = make_uint2(childBase, (hm & 0xff000000u) | imask);
      tg = make_uint2(primBase, hm & 0x00ffffffu);
    } else {
      tg = ng;  // a popped primitive group (or nothing)
      ng = make_uint2(0u, 0u);
    }

    if (inBlas) {
      while (tg.y) {
        const int b = __ffs(tg.y) - 1;
        tg.y &= tg.y - 1;
        const float4* tp = reinterpret_cast<const float4*>(tris + tg.x + b);
        const float4 a = __ldg(tp + 0), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
        if (DETAIL) tc.tris++;
        // Moller-Trumbore, contract arithmetic, same operation order as oracle intersectTri()
        const V3 dd = mk3(r.dx, r.dy, r.dz);
        const V3 E1 = mk3(e1.x, e1.y, e1.z), E2 = mk3(e2.x, e2.y, e2.z);
        const V3 pv = ccross(dd, E2);
        const float det = cdot(E1, pv);
        if (det == 0.0f) continue;
        const float inv = cdiv(1.0f, det);
        const V3 tv = csub3(mk3(r.ox, r.oy, r.oz), mk3(a.x, a.y, a.z));
        const float u = cmul(cdot(tv, pv), inv);
        if (!(u >= 0.0f && u <= 1.0f)) continue;
        const V3 qv = ccross(tv, E1);
        const float v = cmul(cdot(dd, qv), inv);
        if (!(v >= 0.0f && cadd(u, v) <= 1.0f)) continue;
        const float t = cmul(cdot(E2, qv), inv);
        if (!(t > tmin)) continue;
        const int32_t prim = int32_t(__float_as_uint(a.w));
        const bool closer =
            t < hit.t || (t == hit.t && hit.inst >= 0 &&
                          (curInst < hit.inst || (curInst == hit.inst && prim < hit.prim)));
        if (!closer) continue;
        if (!ANY && nonOpaque) {
          const uint32_t g = sc.instSsbo[curInst].geometryIndex;
          const uint32_t mi = __ldg(sc.geoms[g].matIndex + prim);
          const float alpha = sc.mats[mi].alpha;
          if (alpha == 0.0f) continue;
          if (anyHitRnd(seed, uint32_t(curInst), uint32_t(prim)) > alpha) continue;
        }
        hit.t = t;
        hit.u = u;
        hit.v = v;
        hit.inst = curInst;
        hit.prim = prim;
        hit.front = det > 0.0f ? 1u : 0u;
        if (ANY) return true;
      }
    } else {
      while (tg.y) {
        const int b = __ffs(tg.y) - 1;
        tg.y &= tg.y - 1;
        const uint32_t ii = __ldg(sc.tlasInstIdx + tg.x + b);
        const float4* ip = reinterpret_cast<const float4*>(sc.inst + ii);
        const float4 r0 = __ldg(ip + 0), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
        const ulonglong2 ptrs = __ldg(reinterpret_cast<const ulonglong2*>(ip + 3));
        if (ptrs.x == 0ull) continue;
        if (DETAIL) tc.insts++;
        // save what is left of this TLAS node, then the marker that brings us back
        if (tg.y && sp < KF_STACK) stack[sp++] = tg;
        if ((ng.y & 0xff000000u) && sp < KF_STACK) stack[sp++] = ng;
        if (sp < KF_STACK) stack[sp++] = make_uint2(0xffffffffu, 0u);
        // world -> object (contract arithmetic, oracle traceInstance())
        V3 oo, od;
        oo.x = cadd(cdot3(r0.x, r0.y, r0.z, o.x, o.y, o.z), r0.w);
        oo.y = cadd(cdot3(r1.x, r1.y, r1.z, o.x, o.y, o.z), r1.w);
        oo.z = cadd(cdot3(r2.x, r2.y, r2.z, o.x, o.y, o.z), r2.w);
        od.x = cdot3(r0.x, r0.y, r0.z, d.x, d.y, d.z);
        od.y = cdot3(r1.x, r1.y, r1.z, d.x, d.y, d.z);
        od.z = cdot3(r2.x, r2.y, r2.z, d.x, d.y, d.z);
        r = setupRay(oo, od);
        nodes = reinterpret_cast<const Node8*>(ptrs.x);
        nonOpaque = (ptrs.y & 1ull) != 0;
        tris = reinterpret_cast<const Tri48*>(ptrs.y & ~1ull);
        curInst = int32_t(ii);
        inBlas = true;
        ng = make_uint2(0u, 0x80000000u);
        tg = make_uint2(0u, 0u);
        break;
      }
      if (inBlas) continue;
    }

    if (!(ng.y & 0xff000000u)) {
      bool done = false;
      for (;;) {
        if (sp == 0) { done = true; break; }
        const uint2 e = stack[--sp];
        if (e.y == 0u) {  // sentinel: back to the top level
          r = setupRay(o, d);
          nodes = sc.tlasNodes;
          inBlas = false;
          continue;
        }
        ng = e;
        break;
      }
      if (done) break;
    }
  }
  return hit.inst >= 0;
}

}  // namespace kf
