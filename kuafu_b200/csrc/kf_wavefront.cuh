// kf_wavefront.cuh -- wavefront path tracing: the reference's raygen -> closest-hit -> shadow
// recursion (PathTrace.rgen:30-141, PathTrace.rchit:322-469) unrolled into per-stage kernels over
// compacted queues, so that every warp runs one stage with (almost) all lanes alive.
//
// Per batch of S samples of every pixel (slot = sample * nPixelSlots + tiled pixel index):
//
//   k_wf_raygen            camera ray + RNG seeds, path state, queue 0
//   for depth = 0 .. maxPathDepth:
//     k_wf_trace<false>    closest hit of every queued ray            (kf_trace.cuh, persistent lanes)
//     k_wf_shade           miss / emissive -> path ends; surface -> BSDF sample, first light that
//                          needs an occlusion ray (speculative contribution), or advance directly
//     k_wf_trace<true>     occlusion ray of every queued light sample (x L light rounds)
//     k_wf_shadow_resolve  commit the light if unoccluded, walk to the next light (multi-light
//                          scenes), then advance: Russian roulette, enqueue next bounce
//   k_wf_finish            per pixel: add the S sample colours in sample order onto the sum buffer
//
// The per-path LCG stream is consumed in exactly the reference order: lobe choice, lobe sample,
// per light [perturbation, (if unoccluded) calcDirectContribution draw], Russian roulette.  The
// contribution of the light whose occlusion ray is in flight is evaluated speculatively on a copy of
// the seed; the shadow kernel commits (contribution, advanced seed) only when the ray is unoccluded.
#pragma once

#include "kf_common.cuh"
#include "kf_shade.cuh"
#include "kf_trace.cuh"
#include "kf_traverse.cuh"

namespace kf {

// 7 resident blocks = 72 registers without spills; 8 (64 registers) spills since the hit / miss deal
// was added and measures 1.6 % slower, 10 (48 registers) 12 % slower.
#ifndef KF_SHADE_THREADS
#define KF_SHADE_THREADS 128  // the hit / miss deal works on one block's worth of queue entries
#endif
#ifndef KF_SHADE_MIN_BLOCKS
#define KF_SHADE_MIN_BLOCKS (896 / KF_SHADE_THREADS)
#endif

// Surface context spilled between light iterations (multi-light scenes only), 6 x float4.
struct WfBuffers {
  float4* rayO;     // origin.xyz, -
  float4* rayD;     // direction.xyz, -
  float4* hitA;     // t, u, v, prim (bits)           } indexed by QUEUE position of the traversal stage that
  int* hitB;        // inst | front << 31, -1 on miss } wrote them (kf_trace.cuh), not by slot
  float4* stateW;   // throughput weight.xyz, seed (bits)
  float4* stateC;   // colour.xyz, -
  float4* shadowL;  // L.xyz, maxDist
  float4* shadowC;  // speculative contribution.xyz, seed after calcDirect (bits)
  float4* ctx;      // 6 per slot (multi-light): N|f, V|a2, diffuse|isInside, specular|k, transmission|-, shadowAcc|-
  uint32_t* queue[2];
  uint32_t* shadowQueue[2];
  uint32_t* counts;  // [0..1] extension queue sizes, [2..3] shadow queue sizes, [4] closest-hit / [5] occlusion fetch cursor
};

struct WfArgs {
  SceneDev sc;
  WfBuffers b;
  const KfrtCamera* cams;
  uint32_t nCams, w, h, tilesX, tilesY;  // 8 x 4 pixel tiles
  uint32_t slotsPerSample;               // nCams * tilesX * tilesY * 32
  uint32_t batchBegin, batchCount;       // global sample index of the first sample, samples in batch
  uint32_t firstSample;                  // sampleBegin of the kfrtRender call
  KfrtPushConstants pc;
  uint32_t clockBase;
  float4* sum;
  float4* albedo;
  float4* normal;
  int2* hitIds;
  float* hitT;
  float* depth;
  unsigned long long* counters;
};

KF_D void slotToPixel(const WfArgs& a, uint32_t pixelSlot, uint32_t& cam, uint32_t& x, uint32_t& y) {
  const uint32_t lane = pixelSlot & 31u;
  uint32_t tile = pixelSlot >> 5;
  const uint32_t tilesPerCam = a.tilesX * a.tilesY;
  cam = tile / tilesPerCam;
  tile -= cam * tilesPerCam;
  const uint32_t ty = tile / a.tilesX, tx = tile - ty * a.tilesX;
  x = tx * 8 + (lane & 7u);
  y = ty * 4 + (lane >> 3);
}

// ---------------------------------------------------------------------------------------------
// One thread per pixel slot, looping over the samples of the batch: the pixel's seed hash (shared by all
// samples, rgen:23) is made once and its jitter stream advanced two draws per sample instead of being
// re-hashed and skipped ahead for every (sample, pixel) pair, and the queue comes out as runs of 256
// consecutive pixels of one sample followed by the same pixels of the next sample -- rays that walk the
// same part of the hierarchy sit next to each other in the first traversal stage.
__global__ void __launch_bounds__(256) k_wf_raygen(WfArgs a) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.b.counts[1] = 0;
    a.b.counts[2] = 0;
    a.b.counts[3] = 0;
  }
  const uint32_t pixelSlot = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = pixelSlot < a.slotsPerSample;
  uint32_t cam = 0, x = 0, y = 0;
  if (valid) {
    slotToPixel(a, pixelSlot, cam, x, y);
    valid = x < a.w && y < a.h;
  }
  // Queue positions: the block reserves room for all its pixels x all samples of the batch with ONE atomic
  // (a warp-aggregated append per sample was a million atomics on one address per launch) and lays its
  // entries out sample by sample -- runs of up to 256 consecutive pixels of one sample followed by the same
  // pixels of the next one, the order the first traversal and shade stages like.
  __shared__ uint32_t sWarpValid[8];
  __shared__ uint32_t sBase;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t validMask = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) sWarpValid[warp] = __popc(validMask);
  __syncthreads();
  uint32_t before = 0, blockValid = 0;
#pragma unroll
  for (uint32_t w = 0; w < 8; w++) {
    const uint32_t v = sWarpValid[w];
    if (w < warp) before += v;
    blockValid += v;
  }
  if (threadIdx.x == 0) sBase = atomicAdd(a.b.counts + 0, blockValid * a.batchCount);
  __syncthreads();
  if (!valid) return;
  const uint32_t rank = before + __popc(validMask & ((1u << lane) - 1u));
  const uint32_t mapping = y * a.w + x;
  // the pixel-jitter stream is shared by all samples of the pixel (rgen:30-37): skip the two draws of
  // each sample before this batch once, then two draws per sample
  uint32_t seed = lcgSkip(tea(mapping, a.clockBase), 2u * a.batchBegin);
  for (uint32_t s = 0; s < a.batchCount; s++) {
    const uint32_t slot = s * a.slotsPerSample + pixelSlot;
    const uint32_t i = a.batchBegin + s;  // global sample index
    uint32_t raySeed = tea(mapping, a.clockBase + 1u + i);
    V3 o, d;
    cameraRay(a.cams + cam, x, y, a.w, a.h, seed, raySeed, o, d);  // draws the sample's two jitter numbers from seed
    a.b.rayO[slot] = make_float4(o.x, o.y, o.z, 0.0f);
    a.b.rayD[slot] = make_float4(d.x, d.y, d.z, 0.0f);
    a.b.stateW[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(raySeed));
    a.b.stateC[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    a.b.queue[0][sBase + s * blockValid + rank] = slot;
  }
}

// The next light whose occlusion ray can matter, with its speculative contribution.  A light that would
// contribute exactly nothing AND draws no random number in calcDirect (a surface of a transmissive material
// seen from inside: PathTrace.rchit:103-171 returns zero before its first draw) leaves colour and seed the
// same whether its ray is occluded or not, so the ray is not traced: 21 % of the occlusion rays of config 3,
// whose glass spheres are lit from inside by every path that crosses them.  Such rays are counted in
// counters[11] (KfrtCounters.shadowRaysSkipped), not in shadowRays.
// A ray that leaves a convex geometry to the front side of the triangle it starts on cannot meet that
// geometry again (k_batch_convex): the traversal stages skip the instance it starts on.  The test keeps a
// margin (sine of the elevation above the triangle's plane > 1e-3): a grazing ray takes the full walk, so that
// what rounding makes of a ray along its own surface stays what the oracle makes of it.  Ng: M^-T (e1 x e2).
KF_D bool leavesSurface(V3 Ng, V3 d) {
  const float s = dot(Ng, d);
  return s > 0.0f && s * s > 1e-6f * dot(Ng, Ng) * dot(d, d);
}

KF_D bool nextRelevantLight(const SceneDev& sc, const Surface& sf, uint32_t& seed, int& k, V3& L, float& maxDist,
                            V3& lightEmission, uint32_t& texFetches, V3& contrib, uint32_t& specSeed, uint32_t& skipped,
                            bool cull) {
  while (nextLight(sc, sf, seed, k, L, maxDist, lightEmission, texFetches)) {
    specSeed = seed;
    contrib = calcDirect(sf, L, lightEmission, specSeed);
    if (!cull || anyNe(contrib, mk3(0.0f)) || specSeed != seed) return true;
    skipped++;
    k++;
  }
  return false;
}

KF_D void storeCtx(float4* __restrict__ c, const Surface& sf, int k, V3 acc) {
  c[0] = make_float4(sf.N.x, sf.N.y, sf.N.z, sf.f);
  c[1] = make_float4(sf.V.x, sf.V.y, sf.V.z, sf.a2);
  c[2] = make_float4(sf.diffuseColor.x, sf.diffuseColor.y, sf.diffuseColor.z, __uint_as_float(sf.isInside));
  c[3] = make_float4(sf.specularColor.x, sf.specularColor.y, sf.specularColor.z, __int_as_float(k));
  c[4] = make_float4(sf.transmissionColor.x, sf.transmissionColor.y, sf.transmissionColor.z, 0.0f);
  c[5] = make_float4(acc.x, acc.y, acc.z, 0.0f);
}
KF_D void loadCtx(const float4* __restrict__ c, Surface& sf, int& k, V3& acc) {
  const float4 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4], c5 = c[5];
  sf.N = mk3(c0.x, c0.y, c0.z);
  sf.f = c0.w;
  sf.V = mk3(c1.x, c1.y, c1.z);
  sf.a2 = c1.w;
  sf.diffuseColor = mk3(c2.x, c2.y, c2.z);
  sf.isInside = __float_as_uint(c2.w);
  sf.specularColor = mk3(c3.x, c3.y, c3.z);
  k = __float_as_int(c3.w);
  sf.transmissionColor = mk3(c4.x, c4.y, c4.z);
  acc = mk3(c5.x, c5.y, c5.z);
}

// ---------------------------------------------------------------------------------------------
// CULL: light samples whose occlusion ray cannot matter are not traced (nextRelevantLight); OWN: rays that
// leave a convex geometry outwards skip the instance they start on (leavesSurface).  Both are the product's
// defaults; they are template parameters because as run-time flags they cost the stage 6 % and 4 %.
template <bool MULTI, bool DETAIL, bool CULL, bool OWN>
__global__ void __launch_bounds__(KF_SHADE_THREADS, KF_SHADE_MIN_BLOCKS) k_wf_shade(WfArgs a, int q, uint32_t depth) {
  const uint32_t count = a.b.counts[q];
  const uint32_t* __restrict__ queue = a.b.queue[q];
  const uint32_t stride = gridDim.x * blockDim.x;
  unsigned long long texTotal = 0;
  uint32_t hitTotal = 0, skipTotal = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.b.counts[4] = 0;  // fetch cursors of the next closest-hit / occlusion stages
    a.b.counts[5] = 0;
  }
  // The block re-deals its 128 queue entries before shading them: hits first, misses last.  A miss
  // is a short path (environment lookup), a hit a long one (gather, textures, BSDF, light sample); in
  // bounce order the two are mixed within every warp, so every warp walked the long path.  After the
  // deal whole warps are miss-only and skip it.
  // (two copies of the staging arrays, used in turn: a warp still reading round n is separated from
  // the writers of round n + 2 by the two barriers of round n + 1)
  constexpr uint32_t WARPS = KF_SHADE_THREADS / 32;
  __shared__ uint32_t sSlotBuf[2][KF_SHADE_THREADS];
  __shared__ uint32_t sQposBuf[2][KF_SHADE_THREADS];
  __shared__ int sHitBuf[2][KF_SHADE_THREADS];
  __shared__ uint32_t sClassBuf[2][2][WARPS];  // per warp: hits, misses
  __shared__ float sNg[3][KF_SHADE_THREADS];    // geometric normal of a convex geometry's hit, until its light is chosen
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t round = 0;
  // The queue entry and hit record of the NEXT round are fetched while this round is shaded: the deal
  // below needs them before its first barrier, and with the loads issued a round ahead the barrier no
  // longer waits for memory (ncu, round 2: 3.5 barrier-stall cycles per issued instruction before).
  uint32_t nextSlot = 0;
  int nextHit = -1;
  {
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 < count) {
      nextSlot = queue[i0];
      nextHit = a.b.hitB[i0];
    }
  }
  for (uint32_t base = blockIdx.x * blockDim.x; base < count; base += stride, round ^= 1u) {
    uint32_t* sSlot = sSlotBuf[round];
    uint32_t* sQpos = sQposBuf[round];
    int* sHitB = sHitBuf[round];
    uint32_t(*sClass)[WARPS] = sClassBuf[round];
    bool valid;
    uint32_t slot = 0, qpos = 0;
    int hB = -1;
    {
      const uint32_t i = base + threadIdx.x;
      const bool have = i < count;
      const uint32_t mySlot = nextSlot;
      const int myHit = nextHit;
      {
        const uint32_t i2 = i + stride;
        if (i2 < count) {
          nextSlot = queue[i2];
          nextHit = a.b.hitB[i2];
        }
      }
      const uint32_t hitMask = __ballot_sync(0xffffffffu, have && myHit != -1);
      const uint32_t missMask = __ballot_sync(0xffffffffu, have && myHit == -1);
      if (lane == 0) {
        sClass[0][warp] = __popc(hitMask);
        sClass[1][warp] = __popc(missMask);
      }
      __syncthreads();
      uint32_t hitsBefore = 0, missesBefore = 0, hitsTotal = 0, missesTotal = 0;
#pragma unroll
      for (uint32_t w = 0; w < WARPS; w++) {
        const uint32_t h = sClass[0][w], m = sClass[1][w];
        if (w < warp) { hitsBefore += h; missesBefore += m; }
        hitsTotal += h;
        missesTotal += m;
      }
      if (have) {
        const uint32_t below = (1u << lane) - 1u;
        const uint32_t pos = myHit != -1 ? hitsBefore + __popc(hitMask & below)
                                         : hitsTotal + missesBefore + __popc(missMask & below);
        sSlot[pos] = mySlot;
        sQpos[pos] = i;
        sHitB[pos] = myHit;
      }
      __syncthreads();
      valid = threadIdx.x < hitsTotal + missesTotal;
      if (valid) {
        slot = sSlot[threadIdx.x];
        qpos = sQpos[threadIdx.x];
        hB = sHitB[threadIdx.x];
      }
    }
    bool toNext = false, toShadow = false, isHit = false;
    if (valid) {
      // a miss needs neither the ray's origin nor a hit record
      const float4 d4 = a.b.rayD[slot];
      float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f), hA = o4;
      if (hB != -1) {
        o4 = a.b.rayO[slot];
        hA = a.b.hitA[qpos];
      }
      const float4 sw = a.b.stateW[slot];
      V3 weight = mk3(sw.x, sw.y, sw.z);
      uint32_t seed = __float_as_uint(sw.w);
      // The path's colour is touched only when something is added to it: a miss, an emissive surface, or
      // -- to keep the reference's `color += emission * weight` with emission = 0 exact -- a weight that
      // is no longer finite (0 * Inf = NaN).  Every other hit leaves its 16 bytes where they are: 32 of
      // the 176 bytes a hit moved through this stage.
      auto addColor = [&](V3 term) {
        const float4 sc4 = a.b.stateC[slot];
        const V3 color = mk3(sc4.x, sc4.y, sc4.z) + term;
        a.b.stateC[slot] = make_float4(color.x, color.y, color.z, 0.0f);
      };
      auto finite3 = [](V3 v) { return isfinite(v.x) && isfinite(v.y) && isfinite(v.z); };
      const V3 ro = mk3(o4.x, o4.y, o4.z), rd = mk3(d4.x, d4.y, d4.z);
      uint32_t tex = 0;
      const uint32_t s = slot / a.slotsPerSample;
      const bool firstHitOutputs = depth == 0 && (a.batchBegin + s) == 0;  // sample 0, depth 0
      uint32_t cam = 0, px = 0, py = 0;
      size_t pi = 0;
      if (firstHitOutputs) {
        slotToPixel(a, slot - s * a.slotsPerSample, cam, px, py);
        pi = (size_t(cam) * a.h + py) * a.w + px;
      }
      if (hB == -1) {
        const V3 emission = shadeMiss(a.sc, a.pc, rd, tex);
        addColor(emission * weight);
        if (firstHitOutputs) {
          a.albedo[pi] = make_float4(0.f, 0.f, 0.f, 1.f);
          a.normal[pi] = make_float4(0.f, 0.f, 0.f, 1.f);
          a.hitIds[pi] = make_int2(-1, -1);
          a.hitT[pi] = 0.0f;
          a.depth[pi] = 0.0f;
        }
      } else {
        isHit = true;
        Hit hit;
        hit.t = hA.x;
        hit.u = hA.y;
        hit.v = hA.z;
        hit.prim = __float_as_int(hA.w);
        hit.inst = int(uint32_t(hB) & 0x7fffffffu);
        hit.front = uint32_t(hB) >> 31;
        if (firstHitOutputs) {
          const float Px = cadd(ro.x, cmul(rd.x, hit.t)), Py = cadd(ro.y, cmul(rd.y, hit.t)),
                      Pz = cadd(ro.z, cmul(rd.z, hit.t));
          const float* vm = a.cams[cam].view;
          a.hitIds[pi] = make_int2(hit.inst, hit.prim);
          a.hitT[pi] = hit.t;
          a.depth[pi] = -cadd(cadd(cadd(cmul(vm[2], Px), cmul(vm[6], Py)), cmul(vm[10], Pz)), vm[14]);
        }
        Surface sf;
        V3 L, w, albedo, emission;
        if (!shadeSurface(a.sc, hit, ro, rd, seed, sf, L, w, albedo, emission, tex,
                          OWN ? &sNg[0][threadIdx.x] : nullptr, KF_SHADE_THREADS)) {
          addColor(emission * weight);  // emissive surface ends the path (rchit:330-334)
          if (firstHitOutputs) {
            a.albedo[pi] = make_float4(0.f, 0.f, 0.f, 1.f);
            a.normal[pi] = make_float4(0.f, 0.f, 0.f, 1.f);
          }
        } else {
          if (firstHitOutputs) {
            a.albedo[pi] = make_float4(albedo.x, albedo.y, albedo.z, 1.f);
            a.normal[pi] = make_float4(sf.N.x, sf.N.y, sf.N.z, 1.f);
          }
          if (!finite3(weight)) addColor(mk3(0.0f) * weight);  // ray.emission = 0 (rgen:109 keeps NaN/Inf semantics)
          weight *= w;                  // rgen:110; the NEE sum is scaled by the post-BSDF weight
          // next extension ray (rchit:462-463); the occlusion rays share its origin, whose spare word says which
          // instance the two rays need not enter (leavesSurface):
          // instance | convex << 29 | extension ray skips it << 30 | occlusion ray skips it << 31
          uint32_t originBits = 0u;
          if (OWN) {
            // (the instance is read again from where the deal left it, the normal from where shadeSurface left it)
            const V3 Ng = mk3(sNg[0][threadIdx.x], sNg[1][threadIdx.x], sNg[2][threadIdx.x]);
            const uint32_t oInst = uint32_t(sHitB[threadIdx.x]) & 0x7fffffffu;
            if ((Ng.x != 0.0f || Ng.y != 0.0f || Ng.z != 0.0f) && oInst < (1u << 29))
              originBits = oInst | (1u << 29) | (leavesSurface(Ng, L) ? 1u << 30 : 0u);
          }
          a.b.rayD[slot] = make_float4(L.x, L.y, L.z, 0.0f);
          int k = 0;
          V3 Ls, le, contrib = mk3(0.0f);
          float maxDist;
          uint32_t specSeed = seed;
          bool need = nextRelevantLight(a.sc, sf, seed, k, Ls, maxDist, le, tex, contrib, specSeed, skipTotal, CULL);
          // A path whose weight the BSDF sample has just taken to exactly zero (a GGX direction below the
          // horizon, rchit:420-428) adds `shadow_color * 0` and then ends before its next random draw
          // (rgen:119): its light sample cannot matter either, unless the contribution is not finite.
          if (!MULTI && CULL && need && allEq(weight, mk3(0.0f)) && finite3(contrib)) {
            need = false;
            skipTotal++;
          }
          if (!MULTI && need && originBits &&
              leavesSurface(mk3(sNg[0][threadIdx.x], sNg[1][threadIdx.x], sNg[2][threadIdx.x]), Ls))
            originBits |= 1u << 31;
          originBits = (originBits >> 30) ? (originBits & ~(1u << 29)) : 0u;  // the spare word of the origin
          a.b.rayO[slot] = make_float4(sf.worldPos.x, sf.worldPos.y, sf.worldPos.z, __uint_as_float(originBits));
          if (need) {
            a.b.shadowL[slot] = make_float4(Ls.x, Ls.y, Ls.z, maxDist);
            a.b.shadowC[slot] = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(specSeed));
            if (MULTI) storeCtx(a.b.ctx + size_t(6) * slot, sf, k, mk3(0.0f));
            a.b.stateW[slot] = make_float4(weight.x, weight.y, weight.z, __uint_as_float(seed));
            toShadow = true;
          } else {
            // no light needs a ray: shadow_color = 0, advance right away
            if (!finite3(weight)) addColor(mk3(0.0f) * weight);
            toNext = advancePath(a.pc, depth, weight, seed);
            a.b.stateW[slot] = make_float4(weight.x, weight.y, weight.z, __uint_as_float(seed));
          }
        }
      }
      texTotal += tex;
    }
    queueAppend(a.b.queue[q ^ 1], a.b.counts + (q ^ 1), toNext, slot);
    queueAppend(a.b.shadowQueue[0], a.b.counts + 2, toShadow, slot);
    hitTotal += isHit ? 1u : 0u;
  }
  // (one atomic per warp for the whole launch: per round it was one more serialised same-address atomic)
  hitTotal = __reduce_add_sync(0xffffffffu, hitTotal);
  skipTotal = __reduce_add_sync(0xffffffffu, skipTotal);
  if ((threadIdx.x & 31u) == 0u && hitTotal) atomicAdd(a.counters + 3, (unsigned long long)hitTotal);
  if ((threadIdx.x & 31u) == 0u && skipTotal) atomicAdd(a.counters + 11, (unsigned long long)skipTotal);
  if (DETAIL) atomicAdd(a.counters + 7, texTotal);
}

// ---------------------------------------------------------------------------------------------
// sq: which shadow queue to read; q: the extension queue being filled for the next bounce.
template <bool MULTI, bool DETAIL, bool CULL>
__global__ void __launch_bounds__(128) k_wf_shadow_resolve(WfArgs a, int sq, int qNext, uint32_t depth) {
  const uint32_t count = a.b.counts[2 + sq];
  const uint32_t* __restrict__ queue = a.b.shadowQueue[sq];
  if (blockIdx.x == 0 && threadIdx.x == 0) a.b.counts[5] = 0;  // occlusion fetch cursor of the next round
  const uint32_t stride = gridDim.x * blockDim.x;
  unsigned long long texTotal = 0;
  uint32_t skipTotal = 0;
  __shared__ uint32_t sCnt[2][8];
  __shared__ uint32_t sBase[2];
  uint32_t appendRound = 0;
  // queue entry and occlusion verdict are fetched a round ahead of the state gathers that depend on them
  uint32_t nextSlot = 0;
  int nextOcc = 0;
  {
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 < count) {
      nextSlot = queue[i0];
      nextOcc = a.b.hitB[i0];
    }
  }
  for (uint32_t base = blockIdx.x * blockDim.x; base < count; base += stride) {
    const uint32_t i = base + threadIdx.x;
    const bool valid = i < count;
    bool toNext = false, toShadow = false;
    uint32_t slot = nextSlot;
    const bool occluded = nextOcc != 0;  // written by k_wf_trace<true> at the queue position
    if (i + stride < count) {
      nextSlot = queue[i + stride];
      nextOcc = a.b.hitB[i + stride];
    }
    if (valid) {
      const float4 sw = a.b.stateW[slot];
      V3 weight = mk3(sw.x, sw.y, sw.z);
      uint32_t seed = __float_as_uint(sw.w);
      V3 shadowColor = mk3(0.0f);
      if (!occluded) {
        // an occluded light leaves no trace: its speculative contribution is never read
        const float4 spec = a.b.shadowC[slot];
        shadowColor = mk3(spec.x, spec.y, spec.z);
        seed = __float_as_uint(spec.w);
      }
      bool pending = false;
      if (MULTI) {
        Surface sf;
        int k;
        V3 acc;
        loadCtx(a.b.ctx + size_t(6) * slot, sf, k, acc);
        const float4 o4 = a.b.rayO[slot];
        sf.worldPos = mk3(o4.x, o4.y, o4.z);
        acc += shadowColor;
        k++;
        V3 Ls, le;
        float maxDist;
        uint32_t tex = 0;
        V3 contrib = mk3(0.0f);
        uint32_t specSeed = seed;
        if (nextRelevantLight(a.sc, sf, seed, k, Ls, maxDist, le, tex, contrib, specSeed, skipTotal, CULL)) {
          a.b.shadowL[slot] = make_float4(Ls.x, Ls.y, Ls.z, maxDist);
          a.b.shadowC[slot] = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(specSeed));
          storeCtx(a.b.ctx + size_t(6) * slot, sf, k, acc);
          a.b.stateW[slot] = make_float4(weight.x, weight.y, weight.z, __uint_as_float(seed));
          pending = true;
          toShadow = true;
        }
        texTotal += tex;
        shadowColor = acc;
      }
      if (!pending) {
        // rgen:111; a zero shadow colour under a finite weight adds nothing, and the path's colour stays
        // where it is (0 * Inf = NaN still lands, as in the reference)
        if (anyNe(shadowColor, mk3(0.0f)) || !(isfinite(weight.x) && isfinite(weight.y) && isfinite(weight.z))) {
          const float4 sc4 = a.b.stateC[slot];
          const V3 color = mk3(sc4.x, sc4.y, sc4.z) + shadowColor * weight;
          a.b.stateC[slot] = make_float4(color.x, color.y, color.z, 0.0f);
        }
        toNext = advancePath(a.pc, depth, weight, seed);
        // (an occluded light below the Russian-roulette depth leaves weight and seed as they were)
        if (seed != __float_as_uint(sw.w) || anyNe(weight, mk3(sw.x, sw.y, sw.z)))
          a.b.stateW[slot] = make_float4(weight.x, weight.y, weight.z, __uint_as_float(seed));
      }
    }
    // one atomic per block and round (per warp it was the stage's bottleneck: 0.73 -> 0.47 ms per 8-spp frame)
    queueAppendBlock(a.b.queue[qNext], a.b.counts + qNext, toNext, slot, sCnt, sBase, appendRound);
    appendRound ^= 1u;
    if (MULTI) queueAppend(a.b.shadowQueue[sq ^ 1], a.b.counts + 2 + (sq ^ 1), toShadow, slot);
  }
  if (MULTI) {
    skipTotal = __reduce_add_sync(0xffffffffu, skipTotal);
    if ((threadIdx.x & 31u) == 0u && skipTotal) atomicAdd(a.counters + 11, (unsigned long long)skipTotal);
  }
  if (DETAIL && MULTI) atomicAdd(a.counters + 7, texTotal);
}

__global__ void k_wf_clear_count(uint32_t* c) { *c = 0; }

// ---------------------------------------------------------------------------------------------
// colours of the batch's samples are added in sample order, continuing the running sum of earlier
// batches, so the float association equals the reference's `colors += color` loop (rgen:140).
__global__ void __launch_bounds__(256) k_wf_finish(WfArgs a) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= a.slotsPerSample) return;
  uint32_t cam, x, y;
  slotToPixel(a, slot, cam, x, y);
  if (x >= a.w || y >= a.h) return;
  const size_t pi = (size_t(cam) * a.h + y) * a.w + x;
  V3 colors = mk3(0.0f);
  if (a.batchBegin != a.firstSample) {
    const float4 s = a.sum[pi];
    colors = mk3(s.x, s.y, s.z);
  }
  for (uint32_t s = 0; s < a.batchCount; s++) {
    const float4 c = a.b.stateC[size_t(s) * a.slotsPerSample + slot];
    colors += mk3(c.x, c.y, c.z);
  }
  a.sum[pi] = make_float4(colors.x, colors.y, colors.z, float(a.batchBegin + a.batchCount - a.firstSample));
}

}  // namespace kf
