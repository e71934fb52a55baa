// kf_traverse.cuh -- building blocks of the two-level traversal of the compressed 8-wide BVH (ray
// setup, the 8-box slab test of one node); the traversal loop itself is the persistent kernel in
// kf_trace.cuh.
//
// Traversal state:
//   node group      (childBase, hits<<24 | kind<<8 | mask) -- children still to visit, in octant
//                    priority order (highest bit first); mask = internal children (bottom level) or
//                    present children (top level, where kind marks the real nodes among them);
//                    node groups wait on the lane's stack
//   triangle group  (primBase, triMask<<16 | hits)         -- leaf triangles of the last node step
//                    (bottom level, never stacked): bit b of hits is triangle
//                    primBase + popc(triMask below b)
#pragma once

#include "kf_common.cuh"

namespace kf {

#define KF_STACK 48

struct TravCounters {
  uint32_t nodes, tris, insts, tlasNodes, entries;
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

// Hit-mask tables (shared memory, filled by every block of the traversal kernel).  The slab test
// yields one *miss* bit per child slot; what the traversal needs is
//   internal children by traversal priority: slot s at bit s ^ octinv   -> perm[octinv][hit slots]
//   leaf triangles: the two-bit field of every hit slot                 -> expand[missed slots]
// Built arithmetically that costs ~50 ALU-pipe instructions per node (byte extraction, variable
// shifts, selects): the ALU pipe is the busiest unit of this kernel (ncu: 66 % active against 28 % for
// the FMA pipe), the load/store path has room, so two table lookups replace them.
struct MaskTables {
  uint8_t perm[8][256];
  uint16_t expand[256];
};
KF_D void fillMaskTables(MaskTables& t) {
  for (uint32_t i = threadIdx.x; i < 8u * 256u; i += blockDim.x) {
    const uint32_t o = i >> 8, h = i & 0xffu;
    uint32_t m = 0;
    for (uint32_t s = 0; s < 8; s++)
      if ((h >> s) & 1u) m |= 1u << (s ^ o);
    t.perm[o][h] = uint8_t(m);
  }
  for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
    uint32_t m = 0;
    for (uint32_t s = 0; s < 8; s++)
      if (!((i >> s) & 1u)) m |= 3u << (2 * s);
    t.expand[i] = uint16_t(m);
  }
}

// Slab-tests the 8 quantised child boxes of a node given as its five 16-byte words; returns the mask
// of the child slots the ray MISSES (bit s = slot s; empty slots may report either, the caller masks
// with imask / triMask).
KF_D uint32_t intersectNodeWords(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4,
                                 const RaySetup& r, float tmin, float tmax, uint32_t& childBase,
                                 uint32_t& primBase, uint32_t& imask, uint32_t& triMask) {
  childBase = n1.x;
  primBase = n1.y;
  triMask = n1.z;
  imask = n0.w >> 24;
  // per-axis grid step 2^(e-127) times the 2^15 of planeFloat(): the builder stores e + 15, so a shift
  // and a mask per axis put it in the exponent field
  const float sx = __uint_as_float((n0.w << 23) & 0x7f800000u);
  const float sy = __uint_as_float((n0.w << 15) & 0x7f800000u);
  const float sz = __uint_as_float((n0.w << 7) & 0x7f800000u);
  const float ax = sx * r.ix, ay = sy * r.iy, az = sz * r.iz;
  const float bx = (__uint_as_float(n0.x) - r.ox) * r.ix - ax;
  const float by = (__uint_as_float(n0.y) - r.oy) * r.iy - ay;
  const float bz = (__uint_as_float(n0.z) - r.oz) * r.iz - az;
  // near/far plane words per axis (two words = 8 children)
  const bool nx = r.ix < 0.0f, ny = r.iy < 0.0f, nz = r.iz < 0.0f;
  const uint32_t lox[2] = {n2.x, n2.y}, loy[2] = {n2.z, n2.w}, loz[2] = {n3.x, n3.y};
  const uint32_t hix[2] = {n3.z, n3.w}, hiy[2] = {n4.x, n4.y}, hiz[2] = {n4.z, n4.w};
  // children 7 .. 0, so that the funnel shift below leaves the verdict of child s at bit s
  uint32_t miss = 0;
#pragma unroll
  for (int h = 1; h >= 0; h--) {
    const uint32_t nearx = nx ? hix[h] : lox[h], farx = nx ? lox[h] : hix[h];
    const uint32_t neary = ny ? hiy[h] : loy[h], fary = ny ? loy[h] : hiy[h];
    const uint32_t nearz = nz ? hiz[h] : loz[h], farz = nz ? loz[h] : hiz[h];
#pragma unroll
    for (int j = 3; j >= 0; j--) {
      const float t0x = fmaf(planeFloat(nearx, j), ax, bx);
      const float t0y = fmaf(planeFloat(neary, j), ay, by);
      const float t0z = fmaf(planeFloat(nearz, j), az, bz);
      const float t1x = fmaf(planeFloat(farx, j), ax, bx);
      const float t1y = fmaf(planeFloat(fary, j), ay, by);
      const float t1z = fmaf(planeFloat(farz, j), az, bz);
      // max(t0, tmin) <= min(t1, tmax): a three-input and a two-input min / max per side, then the sign of
      // one subtraction (+0 on equality, never NaN: all operands are finite) shifted into the mask --
      // 5 instructions per child where three sign-bit subtractions + OR + byte packing took 7 (measured on
      // config 3: -3.3 % on both traversal stages).
      const float t0 = fmaxf(fmaxf(fmaxf(t0x, t0y), t0z), tmin);
      const float t1 = fminf(fminf(fminf(t1x, t1y), t1z), tmax);
      miss = __funnelshift_l(__float_as_uint(t1 - t0), miss, 1);  // miss = miss << 1 | sign
    }
  }
  return miss & 0xffu;
}

// The same test with the node fetched here (5 x 16 B loads).
KF_D uint32_t intersectNode(const Node8* __restrict__ node, const RaySetup& r, float tmin, float tmax,
                            uint32_t& childBase, uint32_t& primBase, uint32_t& imask, uint32_t& triMask) {
  const uint4* q = reinterpret_cast<const uint4*>(node);
  const uint4 n0 = __ldg(q + 0);
  const uint4 n1 = __ldg(q + 1);
  const uint4 n2 = __ldg(q + 2);
  const uint4 n3 = __ldg(q + 3);
  const uint4 n4 = __ldg(q + 4);
  return intersectNodeWords(n0, n1, n2, n3, n4, r, tmin, tmax, childBase, primBase, imask, triMask);
}

// Order-independent surrogate of the stochastic any-hit draw (reference PathTrace.rahit:30-48;
// declared deviation D5, see oracle/kf_oracle.cpp).
KF_D float anyHitRnd(uint32_t seed, uint32_t inst, uint32_t prim) {
  uint32_t h = tea(seed ^ (prim * 0x9e3779b9u), inst);
  return float(h & 0x00FFFFFFu) * (1.0f / 16777216.0f);
}

}  // namespace kf
