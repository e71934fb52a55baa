// kf_ref_host.cpp -- TEST INFRASTRUCTURE.  The CPU "ray-tracing pipeline" that runs the reference's
// own shaders (compiled from /root/reference/resources/shaders through glsl2cpp.py + glsl_shim.hpp)
// the way vkCmdTraceRaysKHR would: one raygen invocation per pixel (src/core/rt/rt.cpp:637-666),
// traceRayEXT dispatching to the closest-hit / any-hit / miss stages of the shader binding table
// (rt.cpp:553-600: miss index 0 = PathTrace.rmiss, 1 = PathTraceShadow.rmiss, one hit group), and
// the descriptor sets bound as the reference binds them:
//   set 0: 0 TLAS, 1 image, 2 albedoImage, 3 normalImage                    (rt.cpp:668-706)
//   set 1: 0 camera UBO, 1 geometry instances, 2 environment cube, 3 directional light,
//          4 point lights, 5 active lights                                  (scene.cpp:467-560)
//   set 2: 0 vertices[], 1 indices[], 2 matIndices[], 3 textures[], 4 materials
// The output library (oracle/_ref/libkf_ref.so) is what the hand-written oracle is pinned against.
//
// Black boxes (not shader code in the reference either): acceleration-structure traversal, the
// triangle test and the texture samplers come from libkf_oracle.so (kfo_trace, kfo_sample_*), and
// clockARB() is the oracle's deterministic surrogate D1 (pixel stream = clockBase, stream of sample
// i = clockBase + 1 + i).
#include <atomic>
#include <thread>
#include <vector>

#include "glsl_shim.hpp"

extern "C" {
struct KfoHitOut {
  float t, u, v;
  int32_t inst, prim, front;
  float worldToObject[12];
};
struct KfoSceneView {
  uint32_t nGeoms, nMats, nInsts, nTex;
  const void* const* verts;
  const void* const* idx;
  const void* const* matIndex;
  const void* mats;
  const void* insts;
  const void* dl;
  const void* pl;
  const void* al;
  int32_t hasEnv;
};
typedef int (*AnyHitFn)(void* user, int32_t inst, int32_t prim);
int kfo_prepare(void* h);
int kfo_scene_view(void* h, KfoSceneView* v);
int kfo_trace(void* h, const float* o, const float* d, float tmin, float tmax, int anyHit, int terminateOnFirst,
              int brute, AnyHitFn fn, void* user, KfoHitOut* out);
void kfo_sample_texture(void* h, int32_t index, float u, float v, float* rgb);
void kfo_sample_cube(void* h, const float* dir, float* rgb);
}

namespace glsl {

struct PtrBlock {  // a buffer block whose only member is an unsized array
  const void* p;
};

struct Pipeline {
  void* scene = nullptr;
  const void* bindings[3][8] = {};
  const void* pushConstants = nullptr;
  int brute = 0;
  uint32_t clockBase = 0, clockCalls = 0, launchTraces = 0;
  uint64_t extRays = 0, shRays = 0, extHits = 0;
  Invocation launch;
  // primary hit of the first sample (the additive hit buffers of kf_rt.h)
  int32_t primInst = -1, primPrim = -1;
  float primT = 0.0f;
  // bound objects
  accelerationStructureEXT tlas{0};
  image2D image{}, albedoImage{}, normalImage{};
  samplerCube env{};
  PtrBlock instances{}, materials{};
  std::vector<PtrBlock> vertices, indices, matIndices;
  std::vector<sampler2D> textures;

  void trace(StageBase& caller, uint flags, uint missIndex, vec3 o, float tmin, vec3 d, float tmax, void* payload);
};

const void* StageBase::bindingPtr(int set, int binding) const { return pl->bindings[set][binding]; }
const void* StageBase::pushConstantPtr() const { return pl->pushConstants; }
uint64_t StageBase::clockARB() {
  // PathTrace.rgen:23 (pixel seed), :29 (timeStart, unused), :32 (one per sample)
  uint32_t k = pl->clockCalls++;
  return k == 0 ? pl->clockBase : pl->clockBase + (k - 1);
}
void StageBase::traceRayEXT(const accelerationStructureEXT&, uint rayFlags, uint, uint, uint, uint missIndex,
                            vec3 origin, float tMin, vec3 direction, float tMax, int payloadLocation) {
  pl->trace(*this, rayFlags, missIndex, origin, tMin, direction, tMax, payloadAt(payloadLocation));
}
vec4 texture(const sampler2D& s, vec2 uv) {
  float c[3];
  kfo_sample_texture(s.pl->scene, s.index, uv.x, uv.y, c);
  return vec4(c[0], c[1], c[2], 1.0f);
}
vec4 texture(const samplerCube& s, vec3 dir) {
  float c[3];
  kfo_sample_cube(s.pl->scene, &dir.x, c);
  return vec4(c[0], c[1], c[2], 1.0f);
}

#include KF_REF_STAGES  // the reference's shaders, translated at build time

struct AnyHitCall {
  Pipeline* pl;
  void* payload;
};
static int anyHitThunk(void* user, int32_t inst, int32_t prim) {
  AnyHitCall& c = *static_cast<AnyHitCall*>(user);
  Invocation iv = c.pl->launch;
  iv.instanceID = inst;
  iv.primitiveID = prim;
  Stage_rahit stage(c.pl, c.payload, iv);
  stage.main();
  return stage.ignoreIntersection_ ? 0 : 1;
}

void Pipeline::trace(StageBase&, uint flags, uint missIndex, vec3 o, float tmin, vec3 d, float tmax, void* payload) {
  const bool opaque = (flags & StageBase::gl_RayFlagsOpaqueEXT) != 0;
  const bool terminate = (flags & StageBase::gl_RayFlagsTerminateOnFirstHitEXT) != 0;
  const bool skipClosestHit = (flags & StageBase::gl_RayFlagsSkipClosestHitShaderEXT) != 0;
  if (terminate) shRays++; else extRays++;
  AnyHitCall ah{this, payload};
  KfoHitOut h;
  const int found = kfo_trace(scene, &o.x, &d.x, tmin, tmax, opaque ? 0 : 1, terminate ? 1 : 0, brute,
                              opaque ? nullptr : anyHitThunk, &ah, &h);
  const bool primary = !terminate && launchTraces++ == 0;
  Invocation iv = launch;
  iv.worldRayOrigin = o;
  iv.worldRayDirection = d;
  if (found) {
    if (primary) { primInst = h.inst; primPrim = h.prim; primT = h.t; }
    if (!terminate) extHits++;
    if (skipClosestHit) return;
    iv.instanceID = h.inst;
    iv.primitiveID = h.prim;
    iv.hitKind = h.front ? StageBase::gl_HitKindFrontFacingTriangleEXT : StageBase::gl_HitKindBackFacingTriangleEXT;
    iv.hitT = h.t;
    for (int c = 0; c < 4; c++)
      iv.worldToObject.c[c] = vec3(h.worldToObject[0 * 4 + c], h.worldToObject[1 * 4 + c], h.worldToObject[2 * 4 + c]);
    Stage_rchit stage(this, payload, iv);
    stage.attribs = vec3(h.u, h.v, 0.0f);
    stage.main();
  } else if (missIndex == 0) {
    Stage_rmiss stage(this, payload, iv);
    stage.main();
  } else {
    Stage_rmiss_shadow stage(this, payload, iv);
    stage.main();
  }
}

}  // namespace glsl

using namespace glsl;

extern "C" {

// Renders nCams cameras of w x h pixels with the reference's shaders.  `scene` is a libkf_oracle.so
// scene handle (kfo_create + kfo_set_*).  image / albedo / normal: camera-major rgba32f, the three
// storage images of PathTrace.rgen (image is read when pushConstants.frameCount > 0).  hitIds /
// hitT (may be NULL): primary hit of the first sample.  counters (may be NULL): [paths,
// extensionRays, shadowRays, extensionHits].  threads <= 0: all cores.
__attribute__((visibility("default")))
int kfref_render(void* scene, const void* cams, uint32_t nCams, uint32_t w, uint32_t h, const void* pushConstants,
                 uint32_t clockBase, int brute, int threads, float* image, float* albedo, float* normal,
                 int32_t* hitIds, float* hitT, uint64_t* counters) {
  kfo_prepare(scene);
  KfoSceneView view;
  kfo_scene_view(scene, &view);
  const uint32_t spp = static_cast<const uint32_t*>(pushConstants)[5];  // Constants::sampleRatePerPixel
  int nt = threads > 0 ? threads : int(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  std::atomic<uint64_t> nextRow{0}, cExt{0}, cSh{0}, cHit{0};
  const uint64_t totalRows = uint64_t(nCams) * h;

  auto worker = [&]() {
    Pipeline pl;
    pl.scene = scene;
    pl.pushConstants = pushConstants;
    pl.brute = brute;
    pl.clockBase = clockBase;
    pl.env.pl = &pl;
    pl.instances.p = view.insts;
    pl.materials.p = view.mats;
    for (uint32_t g = 0; g < view.nGeoms; g++) {
      pl.vertices.push_back({view.verts[g]});
      pl.indices.push_back({view.idx[g]});
      pl.matIndices.push_back({view.matIndex[g]});
    }
    for (uint32_t t = 0; t < view.nTex; t++) pl.textures.push_back({&pl, int(t)});
    pl.bindings[0][0] = &pl.tlas;
    pl.bindings[0][1] = &pl.image;
    pl.bindings[0][2] = &pl.albedoImage;
    pl.bindings[0][3] = &pl.normalImage;
    pl.bindings[1][1] = &pl.instances;
    pl.bindings[1][2] = &pl.env;
    pl.bindings[1][3] = view.dl;
    pl.bindings[1][4] = view.pl;
    pl.bindings[1][5] = view.al;
    pl.bindings[2][0] = pl.vertices.data();
    pl.bindings[2][1] = pl.indices.data();
    pl.bindings[2][2] = pl.matIndices.data();
    pl.bindings[2][3] = pl.textures.data();
    pl.bindings[2][4] = &pl.materials;
    for (;;) {
      const uint64_t row = nextRow.fetch_add(1);
      if (row >= totalRows) break;
      const uint32_t c = uint32_t(row / h), y = uint32_t(row % h);
      const size_t off = size_t(c) * w * h;
      pl.bindings[1][0] = static_cast<const char*>(cams) + size_t(c) * 320;  // CameraUBO, camera.hpp:181-192
      pl.image = {image + 4 * off, int(w), int(h)};
      pl.albedoImage = {albedo + 4 * off, int(w), int(h)};
      pl.normalImage = {normal + 4 * off, int(w), int(h)};
      for (uint32_t x = 0; x < w; x++) {
        pl.launch = Invocation();
        pl.launch.launchID = {x, y, 0};
        pl.launch.launchSize = {w, h, 1};
        pl.clockCalls = 0;
        pl.launchTraces = 0;
        pl.primInst = pl.primPrim = -1;
        pl.primT = 0.0f;
        Stage_rgen stage(&pl, nullptr, pl.launch);
        stage.main();
        const size_t pi = off + size_t(y) * w + x;
        if (hitIds) { hitIds[2 * pi] = pl.primInst; hitIds[2 * pi + 1] = pl.primPrim; }
        if (hitT) hitT[pi] = pl.primT;
      }
    }
    cExt += pl.extRays;
    cSh += pl.shRays;
    cHit += pl.extHits;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  if (counters) {
    counters[0] = uint64_t(nCams) * w * h * spp;
    counters[1] = cExt;
    counters[2] = cSh;
    counters[3] = cHit;
  }
  return 0;
}

}  // extern "C"
