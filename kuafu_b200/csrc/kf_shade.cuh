// kf_shade.cuh -- device restatement of the reference's closest-hit / miss shading
// (PathTrace.rchit, PathTrace.rmiss, base/Sampling.glsl, base/Random.glsl), split into resumable
// pieces so that next-event estimation can be deferred by a wavefront scheduler without changing
// the order in which the per-path LCG stream is consumed:
//
//   shadeSurface()  material fetch + textures + lobe choice + BSDF sample      (rchit:322-455)
//   nextLight()     walks directional -> point[32] -> active[8] lights in reference order and
//                   stops at the first one that needs an occlusion ray          (rchit:175-316)
//   calcDirect()    contribution of an unoccluded light                          (rchit:103-171)
//   shadeMiss()     environment / clear colour                                   (rmiss:12-31)
#pragma once

#include "kf_common.cuh"

namespace kf {

// ---------------------------------------------------------------------------------------------
// sampling + microfacet terms (reference base/Random.glsl:58-136, base/Sampling.glsl)
// ---------------------------------------------------------------------------------------------
KF_D V3 getPerpendicularVector(V3 u) {
  const float ax = fabsf(u.x), ay = fabsf(u.y), az = fabsf(u.z);
  const uint32_t xm = ((ax - ay) < 0 && (ax - az) < 0) ? 1u : 0u;
  const uint32_t ym = (ay - az) < 0 ? (1u ^ xm) : 0u;
  const uint32_t zm = 1u ^ (xm | ym);
  return cross(u, mk3(float(xm), float(ym), float(zm)));
}
KF_D float Schlick(float cosine, float ior) {
  float r0 = (1.0f - ior) / (1.0f + ior);
  r0 *= r0;
  // GLSL pow(x, y) is exp2(y * log2(x)) (Vulkan precision table: "inherited from exp2(y * log2(x))"),
  // undefined for x < 0: the clamp makes that case 0 instead of NaN (same definition as the oracle)
  return r0 + (1.0f - r0) * exp2f(5.0f * log2f(fmaxf(1.0f - cosine, 0.0f)));
}
KF_D float ggxNormalDistribution(float NdotH, float a2) {
  const float d = fmaxf(NdotH * NdotH * (a2 - 1) + 1, 1e-6f);
  return a2 / (d * d * KF_PI);
}
KF_D V3 sampleGGX(uint32_t& seed, float a2, V3 N) {
  const float rx = rnd(seed);
  const float ry = rnd(seed);
  const V3 B = getPerpendicularVector(N);
  const V3 T = cross(B, N);
  const float cosThetaH = sqrtf(clampf((1.0f - rx) / ((a2 - 1.0f) * rx + 1), 0, 1));
  const float sinThetaH = sqrtf(clampf(1.0f - cosThetaH * cosThetaH, 0, 1));
  const float phiH = ry * KF_PI * 2.0f;
  float sp, cp;
  sincosf(phiH, &sp, &cp);
  return T * (sinThetaH * cp) + B * (sinThetaH * sp) + N * cosThetaH;
}
KF_D float G1(float dotValue, float a2) {
  return (2 * dotValue) / (dotValue + sqrtf(a2 + (1 - a2) * (dotValue * dotValue)));
}
KF_D float GeometricShadowing(float NdotV, float NdotL, float a2) { return G1(NdotL, a2) * G1(NdotV, a2); }
KF_D V3 schlickFresnel(V3 f0, float lDotH) {
  return f0 + (mk3(1.0f) - f0) * exp2f((-5.55473f * lDotH - 6.98316f) * lDotH);
}
KF_D V3 transformLocalToWorld(V3 direction, V3 normal) {
  V3 tangent;
  if (fabsf(normal.x) > fabsf(normal.y))
    tangent = mk3(normal.z, 0, -normal.x) / sqrtf(normal.x * normal.x + normal.z * normal.z);
  else
    tangent = mk3(0, -normal.z, normal.y) / sqrtf(normal.y * normal.y + normal.z * normal.z);
  const V3 bitangent = cross(normal, tangent);
  return direction.x * tangent + direction.y * bitangent + direction.z * normal;
}
KF_D V3 cosineHemisphereSampling(uint32_t& seed, V3 normal) {
  const float u0 = rnd(seed);
  const float u1 = rnd(seed);
  const float sq = sqrtf(1.0f - u1);
  float s, c;
  sincosf(2 * KF_PI * u0, &s, &c);
  return transformLocalToWorld(mk3(c * sq, s * sq, sqrtf(u1)), normal);
}
KF_D V3 uniformSphereSampling(uint32_t& seed) {
  V3 p;
  do {
    const float a = rnd(seed), b = rnd(seed), c = rnd(seed);
    p = mk3(a, b, c) * 2.0f - mk3(1.0f);
  } while (dot(p, p) >= 1.0f);
  return p;
}
// returns the disk sample; bit-exact with the oracle (pure +,-,*)
KF_D void diskSampling(uint32_t& seed, float& px, float& py) {
  do {
    const float a = rnd(seed), b = rnd(seed);
    px = csub(cmul(2.0f, a), 1.0f);
    py = csub(cmul(2.0f, b), 1.0f);
  } while (cadd(cmul(px, px), cmul(py, py)) >= 1.0f);
}
KF_D V3 reflectv(V3 I, V3 N) { return I - (2.0f * dot(N, I)) * N; }
KF_D V3 refractv(V3 I, V3 N, float eta) {
  const float NdotI = dot(N, I);
  const float k = 1.0f - eta * eta * (1.0f - NdotI * NdotI);
  if (k < 0.0f) return mk3(0.0f);
  return eta * I - (eta * NdotI + sqrtf(k)) * N;
}

// ---------------------------------------------------------------------------------------------
// textures (reference vkCore.hpp:580-637: R8G8B8A8Srgb, bilinear, repeat, LOD 0)
// ---------------------------------------------------------------------------------------------
KF_D float wrap01(float u) {
  float f = u - floorf(u);
  if (!(f >= 0.0f && f <= 1.0f)) f = 0.0f;
  return f;
}
KF_D V3 texelLinear(const SceneDev& sc, const uchar4* __restrict__ texels, uint32_t w, int x, int y) {
  const uchar4 t = __ldg(texels + size_t(y) * w + x);
  return mk3(__ldg(sc.srgbToLinear + t.x), __ldg(sc.srgbToLinear + t.y), __ldg(sc.srgbToLinear + t.z));
}
KF_D V3 bilinear(const SceneDev& sc, const uchar4* __restrict__ texels, uint32_t W, uint32_t H, float x,
                 float y, bool repeat) {
  const float x0f = floorf(x), y0f = floorf(y);
  const float fx = x - x0f, fy = y - y0f;
  int x0 = int(x0f), y0 = int(y0f), x1 = x0 + 1, y1 = y0 + 1;
  if (repeat) {
    if (x0 < 0) x0 += int(W);
    if (y0 < 0) y0 += int(H);
    if (x1 >= int(W)) x1 -= int(W);
    if (y1 >= int(H)) y1 -= int(H);
  } else {
    x0 = max(0, min(int(W) - 1, x0));
    x1 = max(0, min(int(W) - 1, x1));
    y0 = max(0, min(int(H) - 1, y0));
    y1 = max(0, min(int(H) - 1, y1));
  }
  const V3 a = texelLinear(sc, texels, W, x0, y0), b = texelLinear(sc, texels, W, x1, y0);
  const V3 c = texelLinear(sc, texels, W, x0, y1), d = texelLinear(sc, texels, W, x1, y1);
  return (a * (1.0f - fx) + b * fx) * (1.0f - fy) + (c * (1.0f - fx) + d * fx) * fy;
}
KF_D V3 sampleTexture(const SceneDev& sc, int idx, float u, float v, uint32_t& texFetches) {
  if (idx < 0 || uint32_t(idx) >= sc.nTex) return mk3(0.0f);
  const TexRec t = sc.texs[idx];
  if (t.w == 0 || t.texels == nullptr) return mk3(0.0f);
  texFetches++;
  const float x = wrap01(u) * float(t.w) - 0.5f;
  const float y = wrap01(v) * float(t.h) - 0.5f;
  return bilinear(sc, t.texels, t.w, t.h, x, y, true);
}
KF_D V3 sampleCube(const SceneDev& sc, V3 r, uint32_t& texFetches) {
  if (sc.envFaces == nullptr || sc.envSize == 0) return mk3(0.0f);
  const float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
  int face;
  float s, t, ma;
  if (az >= ax && az >= ay) {
    face = r.z < 0 ? 5 : 4;
    s = r.z < 0 ? -r.x : r.x;
    t = -r.y;
    ma = az;
  } else if (ay >= ax) {
    face = r.y < 0 ? 3 : 2;
    s = r.x;
    t = r.y < 0 ? -r.z : r.z;
    ma = ay;
  } else {
    face = r.x < 0 ? 1 : 0;
    s = r.x < 0 ? r.z : -r.z;
    t = -r.y;
    ma = ax;
  }
  float u = 0.5f * (s / ma + 1.0f), v = 0.5f * (t / ma + 1.0f);
  if (!(u >= 0.0f && u <= 1.0f)) u = 0.0f;
  if (!(v >= 0.0f && v <= 1.0f)) v = 0.0f;
  texFetches++;
  const uint32_t S = sc.envSize;
  return bilinear(sc, sc.envFaces + size_t(face) * S * S, S, S, u * float(S) - 0.5f, v * float(S) - 0.5f, false);
}

// ---------------------------------------------------------------------------------------------
// Shading
// ---------------------------------------------------------------------------------------------
// What next-event estimation needs from the surface (84 bytes when spilled by the wavefront).
struct Surface {
  V3 worldPos, N, V;
  float f, a2;
  V3 diffuseColor, specularColor, transmissionColor;
  uint32_t isInside;
};

// reference PathTrace.rmiss:12-31
KF_D V3 shadeMiss(const SceneDev& sc, const KfrtPushConstants& pc, V3 dir, uint32_t& texFetches) {
  const V3 d2 = mk3(-dir.y, dir.z, -dir.x);
  if (pc.useEnvironmentMap) return sampleCube(sc, d2, texFetches);
  return mk3(pc.clearColor[0], pc.clearColor[1], pc.clearColor[2]) * pc.clearColor[3];
}

// reference PathTrace.rchit:66-98 + :322-455.  Returns false when the hit is emissive (path ends,
// `emission` valid); otherwise fills the surface context, the sampled direction L, the BSDF weight
// and the first-hit albedo.
// ngOut (shared memory, stride ngStride floats; may be NULL): receives the geometric normal e1 x e2 of the hit
// triangle of a CONVEX geometry, taken to world space like a shading normal (M^-T n, so that dot(Ng, d) has
// the sign of the object-space normal . M^-1 d), or zero for other geometries -- written here, where the
// rows of the inverse are in registers anyway, and read back by the caller once the outgoing directions are
// known (kept in registers across the BSDF code it costs the kernel spills).
KF_D bool shadeSurface(const SceneDev& sc, const Hit& h, V3 rayO, V3 rayD, uint32_t& seed, Surface& sf,
                       V3& L, V3& weight, V3& albedo, V3& emission, uint32_t& texFetches,
                       float* ngOut = nullptr, uint32_t ngStride = 0) {
  const float4* ip = reinterpret_cast<const float4*>(sc.inst + h.inst);
  const float4 r0 = __ldg(ip + 0), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
  const ulonglong2 p23 = __ldg(reinterpret_cast<const ulonglong2*>(ip + 4));
  // normals, texture coordinates and material index of the triangle: one 64-byte ShadeTri record
  // (the same bits the reference reads through indices -> vertices -> matIndices, rchit:52-98)
  const float4* sp = reinterpret_cast<const float4*>(reinterpret_cast<const ShadeTri*>(p23.y) + h.prim);
  const float4 s0 = __ldg(sp + 0), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2), s3 = __ldg(sp + 3);
  const float bx = 1.0f - h.u - h.v, by = h.u, bz = h.v;
  const V3 n0 = mk3(s0.x, s0.y, s0.z), n1 = mk3(s0.w, s1.x, s1.y), n2 = mk3(s1.z, s1.w, s2.x);
  const V3 ln = n0 * bx + n1 * by + n2 * bz;
  V3 wn;
  wn.x = (ln.x * r0.x + ln.y * r1.x) + ln.z * r2.x;
  wn.y = (ln.x * r0.y + ln.y * r1.y) + ln.z * r2.y;
  wn.z = (ln.x * r0.z + ln.y * r1.z) + ln.z * r2.z;
  V3 N = normalize(wn);
  if (ngOut) {
    float gx = 0.0f, gy = 0.0f, gz = 0.0f;
    if (p23.x != 0ull) {
      const float4 gn = __ldg(reinterpret_cast<const float4*>(p23.x) + h.prim);
      gx = (gn.x * r0.x + gn.y * r1.x) + gn.z * r2.x;
      gy = (gn.x * r0.y + gn.y * r1.y) + gn.z * r2.y;
      gz = (gn.x * r0.z + gn.y * r1.z) + gn.z * r2.z;
    }
    ngOut[0] = gx;
    ngOut[ngStride] = gy;
    ngOut[2 * ngStride] = gz;
  }
  const V3 worldPos = rayO + rayD * h.t;
  const float uvx = (s2.y * bx + s2.w * by) + s3.y * bz;
  const float uvy = (s2.z * bx + s3.x * by) + s3.z * bz;
  const uint32_t matIndex = __float_as_uint(s3.w);
  const float4* mp = reinterpret_cast<const float4*>(sc.mats + matIndex);
  const float4 m0 = __ldg(mp + 0), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2), m3 = __ldg(mp + 3);
  const int4 m4 = __ldg(reinterpret_cast<const int4*>(mp + 4));
  // m0 diffuse, m1 emission, m2 = alpha,metallic,specular,roughness, m3 = ior,transmission,difTex,metTex
  // m4 = roughTex, transTex, pad, pad
  const float matMetallic = m2.y, matSpecular = m2.z, matRoughness = m2.w, matIor = m3.x, matTransmission = m3.y;
  const int diffuseTex = __float_as_int(m3.z), metallicTex = __float_as_int(m3.w);
  const int roughnessTex = m4.x, transmissionTex = m4.y;

  emission = mk3(m1.x, m1.y, m1.z) * m1.w;
  if (anyNe(emission, mk3(0.0f))) return false;

  V3 baseColor = mk3(m0.x, m0.y, m0.z);
  if (diffuseTex >= 0) baseColor = sampleTexture(sc, diffuseTex, uvx, uvy, texFetches);
  baseColor = baseColor / KF_PI;
  const float metallic = metallicTex >= 0 ? sampleTexture(sc, metallicTex, uvx, uvy, texFetches).x : matMetallic;
  const float roughness = roughnessTex >= 0 ? sampleTexture(sc, roughnessTex, uvx, uvy, texFetches).x : matRoughness;
  const float a2 = roughness * roughness;
  const float transmission =
      transmissionTex >= 0 ? sampleTexture(sc, transmissionTex, uvx, uvy, texFetches).x : matTransmission;
  const float f = fmaxf(matIor, 1e-5f);
  const float diffuse_weight = (1.0f - clampf(metallic, 0.0f, 1.0f)) * (1.0f - clampf(transmission, 0.0f, 1.0f));
  const float final_transmission = clampf(transmission, 0.0f, 1.0f) * (1.0f - clampf(metallic, 0.0f, 1.0f));
  const float specular_weight = 1.0f - final_transmission;

  weight = mk3(0.0f);
  const V3 V = normalize(-rayD);
  const bool isInside = h.front == 0u;
  N = isInside ? -N : N;
  L = mk3(0.0f);
  const float NdotV = dot(N, V);

  const V3 diffuseColor = diffuse_weight * baseColor;
  const V3 specularColor =
      specular_weight * (baseColor * metallic + (matSpecular * 0.08f * mk3(1.0f)) * (1.0f - metallic));
  const V3 transmissionColor = transmission * baseColor;

  const float diffuseLum = length(diffuseColor);
  const float specularLum = length(specularColor);
  float probDiffuse = diffuseLum / (diffuseLum + specularLum);
  if (diffuseLum == 0 && specularLum == 0) {
    if (allEq(baseColor, mk3(0.0f))) {
      if (diffuse_weight == 1) probDiffuse = 1.0f;
      else if (diffuse_weight == 0) probDiffuse = 0.0f;
      else probDiffuse = 0.5f;
    } else
      probDiffuse = 0.0f;
  } else {
    probDiffuse *= specular_weight;
  }
  const bool chooseDiffuse = rnd(seed) < probDiffuse;
  if (chooseDiffuse) {
    L = cosineHemisphereSampling(seed, N);
    const float NdotL = clampf(dot(N, L), 0, 1);
    weight = KF_PI * diffuseColor * NdotL / probDiffuse;
  } else {
    if (final_transmission == 0) {
      const V3 H = sampleGGX(seed, a2, N);
      const float HdotV = dot(H, V);
      L = 2 * HdotV * H - V;
      const float NoV = fmaxf(dot(N, V), 1e-7f);
      const float NoL = dot(N, L);
      const float NoH = fmaxf(dot(N, H), 1e-7f);
      const float VoH = fmaxf(dot(V, H), 1e-7f);
      if (NoL >= 0) {
        const float G = GeometricShadowing(NoV, NoL, a2);
        const V3 F = schlickFresnel(specularColor, VoH);
        weight = KF_PI * F * G * VoH / (NoH * NoV * (1 - probDiffuse));
      } else
        weight = mk3(0.0f);
    } else {
      const float ior = isInside ? 1 / f : f;
      const float _dot = isInside ? NdotV * ior : NdotV;
      const V3 refractedL = refractv(-V, N, 1 / f);
      const float reflectProb = anyNe(refractedL, mk3(0.0f)) ? Schlick(_dot, f) : 1.0f;
      if (rnd(seed) >= reflectProb) {
        L = refractedL;
        weight = KF_PI * transmissionColor / (1 - probDiffuse);
      } else {
        L = reflectv(-V, N);
        weight = KF_PI * transmissionColor / (1 - probDiffuse);
      }
    }
  }
  sf.worldPos = worldPos;
  sf.N = N;
  sf.V = V;
  sf.f = f;
  sf.a2 = a2;
  sf.diffuseColor = diffuseColor;
  sf.specularColor = specularColor;
  sf.transmissionColor = transmissionColor;
  sf.isInside = isInside ? 1u : 0u;
  albedo = baseColor;
  return true;
}

// reference PathTrace.rchit:103-171 (the caller has established that the light is unoccluded)
KF_D V3 calcDirect(const Surface& sf, V3 L, V3 lightEmission, uint32_t& seed) {
  const V3 N = sf.N, V = sf.V;
  const float a2 = sf.a2;
  V3 weight = mk3(0.0f);
  if (anyNe(sf.transmissionColor, mk3(0.0f))) {
    if (!sf.isInside) {
      const float NdotV0 = dot(N, V);
      const V3 refractedL = refractv(-V, N, 1 / sf.f);
      const float reflectProb = anyNe(refractedL, mk3(0.0f)) ? Schlick(NdotV0, sf.f) : 1.0f;
      if (rnd(seed) <= reflectProb) {
        const V3 H = normalize(L + V);
        const float NdotL = dot(N, L);
        const float NdotH = dot(N, H);
        const float HdotV = dot(H, V);
        const float NdotV = fmaxf(dot(N, V), 1e-6f);
        const float LdotH = dot(L, H);
        const float D = ggxNormalDistribution(NdotH, a2);
        const float G = GeometricShadowing(NdotL, NdotV, a2);
        const V3 F = schlickFresnel(sf.transmissionColor, LdotH);
        weight = D * F * G * HdotV / NdotH * NdotV;
      }
    }
  } else {
    const V3 H = normalize(L + V);
    const float NdotL = dot(N, L);
    const float NdotH = dot(N, H);
    const float HdotV = dot(H, V);
    const float NdotV = fmaxf(dot(N, V), 1e-6f);
    const float LdotH = dot(L, H);
    const float D = ggxNormalDistribution(NdotH, a2);
    const float G = GeometricShadowing(NdotL, NdotV, a2);
    const V3 F = schlickFresnel(sf.specularColor, LdotH);
    const float diffuseLum = length(sf.diffuseColor);
    const float specularLum = length(sf.specularColor);
    float probDiffuse = diffuseLum / (diffuseLum + specularLum);
    if (diffuseLum == 0 && specularLum == 0) probDiffuse = 0.5f;
    const V3 diffuseWeight = sf.diffuseColor * mk3(NdotL);
    const V3 specularWeight = D * F * G * HdotV / NdotH * NdotV;
    weight = rnd(seed) < probDiffuse ? diffuseWeight * probDiffuse : specularWeight * (1 - probDiffuse);
  }
  return lightEmission * weight;
}

// Light slots in reference evaluation order: 0 directional, 1..32 point, 33..40 active.
#define KF_NUM_LIGHT_SLOTS 41

// Advances `k` to the next light that needs an occlusion ray (N.L > 0) and returns its direction,
// range and emission; returns false once all lights are exhausted.  Lights that are switched off,
// black, outside the projector cone or below the horizon contribute nothing and are skipped after
// consuming exactly the random numbers the reference consumes for them.
KF_D bool nextLight(const SceneDev& sc, const Surface& sf, uint32_t& seed, int& k, V3& L, float& maxDist,
                    V3& lightEmission, uint32_t& texFetches) {
  for (;; k++) {
    // jump to the next slot that is switched on at all (sc.lightMask, computed by kfrtSetLights from
    // the same conditions the reference tests before it draws any random number for a light)
    const unsigned long long pending = k < KF_NUM_LIGHT_SLOTS ? (sc.lightMask >> k) : 0ull;
    if (pending == 0ull) {
      k = KF_NUM_LIGHT_SLOTS;
      return false;
    }
    k += __ffsll((long long)pending) - 1;
    if (k == 0) {  // rchit:204-223
      const float4 dir = __ldg(reinterpret_cast<const float4*>(sc.dl->direction));
      const float4 rgbs = __ldg(reinterpret_cast<const float4*>(sc.dl->rgbs));
      lightEmission = mk3(rgbs.x, rgbs.y, rgbs.z) * rgbs.w;
      if (allEq(lightEmission, mk3(0.0f))) continue;
      L = -mk3(dir.x, dir.y, dir.z);
      if (dir.w != 0) {
        const float a = rnd(seed), b = rnd(seed), c = rnd(seed);
        L = normalize(L + dir.w * mk3(a, b, c));
      }
      maxDist = 1e6f;
    } else if (k <= 32) {  // rchit:227-258
      const int i = k - 1;
      const float4 rgbs = __ldg(reinterpret_cast<const float4*>(sc.pl->rgbs[i]));
      if (!(rgbs.w > 0)) continue;
      const V3 rgb = mk3(rgbs.x, rgbs.y, rgbs.z);
      if (length(rgb * rgbs.w) == 0) continue;
      const float4 posr = __ldg(reinterpret_cast<const float4*>(sc.pl->posr[i]));
      V3 lpos = mk3(posr.x, posr.y, posr.z);
      if (posr.w != 0) {
        const V3 perturb = uniformSphereSampling(seed);
        lpos += posr.w * normalize(perturb);
      }
      L = lpos - sf.worldPos;
      const float d = length(L);
      L = normalize(L);
      lightEmission = rgb * rgbs.w / d / d;
      maxDist = d;
    } else {  // rchit:262-316
      const int i = k - 33;
      const float4 front = __ldg(reinterpret_cast<const float4*>(sc.al->front[i]));
      if (!(front.w > 0)) continue;
      const float4 rgbs = __ldg(reinterpret_cast<const float4*>(sc.al->rgbs[i]));
      const V3 rgb = mk3(rgbs.x, rgbs.y, rgbs.z);
      if (length(rgb * rgbs.w) == 0) continue;
      const float4 sftp = __ldg(reinterpret_cast<const float4*>(sc.al->sftp[i]));
      const float4 pos = __ldg(reinterpret_cast<const float4*>(sc.al->position[i]));
      V3 lpos = mk3(pos.x, pos.y, pos.z);
      if (sftp.x != 0) {
        const V3 perturb = uniformSphereSampling(seed);
        lpos += sftp.x * normalize(perturb);
      }
      L = lpos - sf.worldPos;
      const float d = length(L);
      L = normalize(L);
      const V3 alightDir = normalize(mk3(front.x, front.y, front.z));
      const float halfAngle = clampf(sftp.y, 0, KF_PI) / 2;
      const float cos_ = dot(alightDir, -L);
      if (!(cos_ > cosf(halfAngle))) continue;
      const int texID = int(sftp.z);
      V3 color = rgb;
      if (texID >= 0) {
        // proj * view * vec4(worldPos, 1) groups as (proj * view) * p; the product is made once in
        // kfrtSetLights
        const float* pv = sc.alProjView + 16 * i;
        float cp[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
          cp[r] = ((__ldg(pv + r) * sf.worldPos.x + __ldg(pv + 4 + r) * sf.worldPos.y) + __ldg(pv + 8 + r) * sf.worldPos.z) +
                  __ldg(pv + 12 + r) * 1.0f;
        const float tu = cp[0] / cp[3], tv = cp[1] / cp[3];
        color *= sampleTexture(sc, texID, tu * 0.5f + 0.5f, tv * 0.5f + 0.5f, texFetches);
      }
      lightEmission = color * rgbs.w / d / d;
      maxDist = d;
    }
    // traceShadowRay (rchit:175-200): below the horizon counts as shadowed, no ray, no random draw
    if (dot(sf.N, L) > 0.0f) return true;
  }
}

// Camera ray of one sample (reference PathTrace.rgen:35-56), contract arithmetic: the primary-ray
// bits must equal the oracle's so that 1-spp hit buffers can be compared bit-exactly.
KF_D void cameraRay(const KfrtCamera* __restrict__ cam, uint32_t x, uint32_t y, uint32_t w, uint32_t h,
                    uint32_t& pixelSeed, uint32_t& raySeed, V3& o, V3& d) {
  const float jx = rnd(pixelSeed);
  const float jy = rnd(pixelSeed);
  const float px = cadd(float(x), jx), py = cadd(float(y), jy);
  const float nx = cdiv(px, float(w)), ny = cdiv(py, float(h));
  const float dx = csub(cmul(nx, 2.0f), 1.0f), dy = csub(cmul(ny, 2.0f), 1.0f);
  const float aperture = cam->position[3];
  const float focus = cam->front[3];
  float ox, oy;
  diskSampling(raySeed, ox, oy);
  ox = cmul(cdiv(aperture, 2.0f), ox);
  oy = cmul(cdiv(aperture, 2.0f), oy);
  float target[4], origin[4], direction[4];
  cmulMat4(cam->projectionInverse, dx, dy, 1.0f, 1.0f, target);
  if (aperture > 0.0f) {
    cmulMat4(cam->viewInverse, ox, oy, 0.0f, 1.0f, origin);
    const V3 t = mk3(csub(cmul(target[0], focus), ox), csub(cmul(target[1], focus), oy),
                     csub(cmul(target[2], focus), 0.0f));
    const V3 dd = cnormalize(t);
    cmulMat4(cam->viewInverse, dd.x, dd.y, dd.z, 0.0f, direction);
  } else {
    cmulMat4(cam->viewInverse, 0.0f, 0.0f, 0.0f, 1.0f, origin);
    const V3 dd = cnormalize(mk3(target[0], target[1], target[2]));
    cmulMat4(cam->viewInverse, dd.x, dd.y, dd.z, 0.0f, direction);
  }
  o = mk3(origin[0], origin[1], origin[2]);
  d = mk3(direction[0], direction[1], direction[2]);
}

}  // namespace kf
