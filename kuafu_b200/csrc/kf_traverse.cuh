// kf_traverse.cuh -- building blocks of the two-level traversal of the compressed 8-wide BVH (ray
// setup, the 8-box slab test of one node); the traversal loop itself is the persistent kernel in
// kf_trace.cuh.
//
// Traversal state is a small stack of 8-byte "groups":
//   node group      (childBase, hits<<24 | kind<<8 | mask) -- children still to visit, in octant
//                    priority order (highest bit first); mask = internal children (bottom level) or
//                    present children (top level, where kind marks the real nodes among them)
//   primitive group (primBase, 24-bit mask)                -- leaf triangles (bottom level)
//   sentinel        (x, 0)                                 -- marks the return to the top level
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

// Reciprocals feed the (conservative) box tests only, so the approximate MUFU.RCP is enough; the
// clamp keeps a zero direction component finite (+-1e20) so that no slab product becomes NaN.
KF_D float __frcp_rn_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
KF_D float boxRcp(float x) { return fminf(fmaxf(__frcp_rn_approx(x), -1e20f), 1e20f); }
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
// The constant goes in PRMT's first operand (always a register, loaded once) and the data word in
// the third, so that the selector can be an immediate; the other way round ptxas makes the constant
// the immediate and re-materialises the selector into a register before each of the 48 PRMTs of a node.
KF_D float planeFloat(uint32_t w, int j) {
  return __uint_as_float(__byte_perm(0x3F800000u, w, 0x3200u | (uint32_t(4 + j) << 4)));
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
  const uint32_t octinv4 = r.octinv * 0x01010101u;
  uint32_t hitmask = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t nearx = nx ? hix[h] : lox[h], farx = nx ? lox[h] : hix[h];
    const uint32_t neary = ny ? hiy[h] : loy[h], fary = ny ? loy[h] : hiy[h];
    const uint32_t nearz = nz ? hiz[h] : loz[h], farz = nz ? loz[h] : hiz[h];
    // four children at a time: where does a hit go in the mask, and which bits does it set?
    // internal children (meta = 0x20 | 24 + slot) land on bit 24 + (slot ^ octinv), i.e. in traversal
    // priority order; leaves set `unary count` bits from their offset; empty slots (meta 0) set none.
    const uint32_t meta4 = meta[h];
    const uint32_t inner4 = ((meta4 & (meta4 << 1)) & 0x10101010u) >> 4;  // 1 per internal child
    const uint32_t index4 = (meta4 ^ (octinv4 & (inner4 * 0xffu))) & 0x1f1f1f1fu;
    const uint32_t bits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float t0x = fmaf(planeFloat(nearx, j), ax, bx);
      const float t0y = fmaf(planeFloat(neary, j), ay, by);
      const float t0z = fmaf(planeFloat(nearz, j), az, bz);
      const float t1x = fmaf(planeFloat(farx, j), ax, bx);
      const float t1y = fmaf(planeFloat(fary, j), ay, by);
      const float t1z = fmaf(planeFloat(farz, j), az, bz);
      const float t0 = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
      const float t1 = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
      const uint32_t bits = __byte_perm(bits4, 0u, 0x4440u | uint32_t(j)), index = __byte_perm(index4, 0u, 0x4440u | uint32_t(j));
      hitmask |= (t0 <= t1) ? (bits << index) : 0u;
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

}  // namespace kf
