// Frame orchestration over the kf_rt.h C ABI (shape of reference src/core/context/context.cpp:
// 298-350 update() and :602-684 recordSwapchainCommandBuffers()).
#include "core/context/context.hpp"

#include "image_io.hpp"
#include "kf_rt.h"

namespace kuafu {

Context::~Context() {
  // the cameras go first, while mLastCameras is still alive for their ~Camera -> forgetCamera()
  mCurrentScene = nullptr;
  mScenes.clear();
  mLastCameras.clear();
  if (mRt) kfrtDestroy(mRt);
  mRt = nullptr;
}

// Error codes of the C ABI become the std::runtime_error the reference throws from KF_CRITICAL /
// its limit checks (include/stdafx.hpp:61-64, src/core/scene.cpp:60-64,115-119).
void Context::check(int rc, const char* what) {
  if (rc == KFRT_OK) return;
  const std::string msg = std::string(what) + ": " + kfrtLastError(mRt);
  global::logger->critical(msg);
  throw std::runtime_error(msg);
}

void Context::init() {
  KF_ASSERT(pConfig, "Context::init without a Config");
  if (pConfig->isHostOnly()) {
    if (mCurrentScene) mCurrentScene->init();
    return;
  }
  if (!mRt) {
    KfrtContext* h = nullptr;
    const int rc = kfrtCreate(pConfig->getDeviceOrdinal(), &h);
    if (rc != KFRT_OK) {
      const std::string msg = std::string("kfrtCreate: ") + kfrtLastError(nullptr);
      global::logger->critical(msg);
      throw std::runtime_error(msg);
    }
    mRt = h;
  }
  check(kfrtSetLimits(mRt, uint32_t(pConfig->mMaxGeometry), uint32_t(pConfig->mMaxGeometryInstances),
                      uint32_t(pConfig->mMaxTextures), uint32_t(pConfig->mMaxMaterials)),
        "kfrtSetLimits");
  if (mCurrentScene) mCurrentScene->init();
}

void Context::pack() {
  KF_ASSERT(mCurrentScene, "Context::pack without a scene");
  Scene& s = *mCurrentScene;
  if (s.mUseEnvironmentMap && s.mEnvironmentMapTexturePath != "mem:" && !s.mEnvironmentMapTexturePath.empty()) {
    const std::string& p = s.mEnvironmentMapTexturePath;
    const bool ktx = p.size() > 4 && p.compare(p.size() - 4, 4, ".ktx") == 0;
    if (!ktx || !io::loadKtxCubeRGBA8(p, s.mWire.envSize, s.mWire.envFaces))
      throw std::runtime_error("cubemap format not supported: " + p);
  }
  s.packMaterialsAndTextures();
  s.packInstances();
  s.packLights();
}

int Context::predictFrameCount() const { return pConfig->mAccumulateFrames ? global::frameCount + 1 : -1; }

KfrtPushConstants Context::makePushConstants(int frameCount) const {
  KfrtPushConstants pc;
  const glm::vec4 cc = mCurrentScene->getClearColor();
  pc.clearColor[0] = cc.x; pc.clearColor[1] = cc.y; pc.clearColor[2] = cc.z; pc.clearColor[3] = cc.w;
  pc.frameCount = frameCount;
  pc.sampleRatePerPixel = pConfig->mPerPixelSampleRate;
  pc.maxPathDepth = pConfig->mPathDepth;
  pc.useEnvironmentMap = mCurrentScene->mUseEnvironmentMap ? 1u : 0u;
  pc.russianRoulette = pConfig->mRussianRoulette ? 1u : 0u;
  pc.russianRouletteMinBounces = pConfig->mRussianRouletteMinBounces;
  pc.nextEventEstimation = pConfig->mNextEventEstimation ? 1u : 0u;
  pc.nextEventEstimationMinBounces = pConfig->mNextEventEstimationMinBounces;
  return pc;
}

void Context::update() {
  if (!mRt && pConfig && pConfig->isHostOnly())
    KF_CRITICAL("This renderer was created host-only (no CUDA device); rendering is not possible. There is no CPU fallback.");
  KF_ASSERT(mRt && mCurrentScene, "Context::update before init()");
  Scene& s = *mCurrentScene;
  if (pConfig->mMaxGeometryChanged || pConfig->mMaxGeometryInstancesChanged || pConfig->mMaxTexturesChanged) {
    check(kfrtSetLimits(mRt, uint32_t(pConfig->mMaxGeometry), uint32_t(pConfig->mMaxGeometryInstances),
                        uint32_t(pConfig->mMaxTextures), uint32_t(pConfig->mMaxMaterials)),
          "kfrtSetLimits");
    pConfig->mMaxGeometryChanged = pConfig->mMaxGeometryInstancesChanged = pConfig->mMaxTexturesChanged = false;
  }
  if (mUploadedScene != &s) {  // switching scenes re-uploads everything ("this is heavy")
    mLastLights.clear();
    check(kfrtClearGeometries(mRt), "kfrtClearGeometries");
    for (auto& g : s.mGeometries)
      if (g) g->initialized = false;
    s.mUploadGeometries = s.mUploadGeometryInstancesToBuffer = true;
    s.mUploadEnvironmentMap = true;
    mUploadedScene = &s;
  }

  if (s.mUploadEnvironmentMap) {
    s.mUploadEnvironmentMap = false;
    if (!s.mUseEnvironmentMap || s.mEnvironmentMapTexturePath.empty()) {
      check(kfrtClearEnvironment(mRt), "kfrtClearEnvironment");
    } else {
      if (s.mEnvironmentMapTexturePath != "mem:") {
        const std::string& p = s.mEnvironmentMapTexturePath;
        const bool ktx = p.size() > 4 && p.compare(p.size() - 4, 4, ".ktx") == 0;
        if (!ktx || !io::loadKtxCubeRGBA8(p, s.mWire.envSize, s.mWire.envFaces))
          throw std::runtime_error("cubemap format not supported: " + p);
      }
      const uint8_t* faces[6];
      for (int f = 0; f < 6; f++) faces[f] = s.mWire.envFaces[f].data();
      check(kfrtSetEnvironmentCube(mRt, faces, s.mWire.envSize), "kfrtSetEnvironmentCube");
    }
  }

  if (s.mUploadGeometries) {  // also (re)loads the projector textures, like the reference
    s.mUploadGeometries = false;
    s.packMaterialsAndTextures();
    check(kfrtUploadMaterials(mRt, reinterpret_cast<const KfrtMaterial*>(s.mWire.materials.data()),
                              uint32_t(s.mWire.materials.size())),
          "kfrtUploadMaterials");
    for (size_t t = 0; t < s.mWire.textures.size(); t++)
      check(kfrtUploadTexture(mRt, uint32_t(t), s.mWire.textures[t].rgba.data(), s.mWire.textures[t].width,
                              s.mWire.textures[t].height),
            "kfrtUploadTexture");
    for (size_t i = 0; i < s.mGeometries.size(); ++i) {
      Geometry* g = s.mGeometries[i].get();
      if (!g || g->initialized) continue;
      check(kfrtUploadGeometry(mRt, uint32_t(i), reinterpret_cast<const KfrtVertex*>(g->vertices.data()),
                               uint32_t(g->vertices.size()), g->indices.data(), uint32_t(g->indices.size()),
                               g->matIndex.data(), uint32_t(g->matIndex.size()), g->isOpaque ? 1 : 0,
                               g->hideRender ? 1 : 0),
            "kfrtUploadGeometry");
      g->initialized = true;
      s.mUploadGeometryInstancesToBuffer = true;  // a new bottom level needs a new top level
    }
  }

  if (s.mUploadGeometryInstancesToBuffer) {
    s.mUploadGeometryInstancesToBuffer = false;
    s.packInstances();
    check(kfrtBuildBlas(mRt), "kfrtBuildBlas");  // only geometries uploaded since the last build
    check(kfrtSetInstances(mRt, reinterpret_cast<const KfrtInstance*>(s.mWire.instances.data()),
                           uint32_t(s.mWire.instances.size())),
          "kfrtSetInstances");
    check(kfrtBuildTlas(mRt), "kfrtBuildTlas");
    mLastTransforms.assign(16 * s.mWire.instances.size(), 0.0f);
    for (size_t i = 0; i < s.mWire.instances.size(); i++)
      std::memcpy(&mLastTransforms[16 * i], &s.mWire.instances[i].transform, 64);
  } else {
    // per-frame path: GeometryInstance::transform is read every frame (reference rt.cpp:372-495)
    s.packInstances();
    std::vector<float> tr(16 * s.mWire.instances.size());
    for (size_t i = 0; i < s.mWire.instances.size(); i++) std::memcpy(&tr[16 * i], &s.mWire.instances[i].transform, 64);
    if (tr != mLastTransforms) {
      check(kfrtRefitTlas(mRt, tr.data(), uint32_t(s.mWire.instances.size())), "kfrtRefitTlas");
      mLastTransforms.swap(tr);
    }
  }

  s.packLights();
  {  // the light blocks are rewritten every frame like the reference's UBOs, uploaded when they changed
    std::vector<uint8_t> now(sizeof(s.mWire.directional) + sizeof(s.mWire.points) + sizeof(s.mWire.actives));
    std::memcpy(now.data(), &s.mWire.directional, sizeof(s.mWire.directional));
    std::memcpy(now.data() + sizeof(s.mWire.directional), &s.mWire.points, sizeof(s.mWire.points));
    std::memcpy(now.data() + sizeof(s.mWire.directional) + sizeof(s.mWire.points), &s.mWire.actives, sizeof(s.mWire.actives));
    if (now != mLastLights) {
      check(kfrtSetLights(mRt, reinterpret_cast<const KfrtDirectionalLight*>(&s.mWire.directional),
                          reinterpret_cast<const KfrtPointLights*>(&s.mWire.points),
                          reinterpret_cast<const KfrtActiveLights*>(&s.mWire.actives)),
            "kfrtSetLights");
      mLastLights.swap(now);
    }
  }

  global::frameCount = predictFrameCount();
}

// A camera whose frame is about to be overwritten on the device keeps a host copy, so that
// render(A); render(B); A->downloadLatestFrame() behaves like the reference's per-camera images.
void Context::stashDisplacedFrames(const std::vector<Camera*>& next) {
  for (Camera* c : mLastCameras) {
    if (!c || !c->mFrames.valid || c->mFrames.owner != this || c->mFrames.serial != mSerial) continue;
    if (std::find(next.begin(), next.end(), c) != next.end()) continue;
    c->mFrames.stash.resize(size_t(c->getWidth()) * c->getHeight() * 4);
    check(kfrtDownloadBGRA8(mRt, c->mFrames.slot, c->mFrames.stash.data(), c->mFrames.stash.size()),
          "kfrtDownloadBGRA8");
  }
}

void Context::forgetCamera(Camera* camera) {
  mLastCameras.erase(std::remove(mLastCameras.begin(), mLastCameras.end(), camera), mLastCameras.end());
}

// A scene that is being destroyed: nothing of it may be looked at again (its cameras unregister
// themselves in ~Camera; the uploaded-scene marker must not match a later scene at the same address).
void Context::forgetScene(Scene* scene) {
  if (mUploadedScene == scene) mUploadedScene = nullptr;
  if (mCurrentScene == scene) mCurrentScene = nullptr;
}

void Context::trace(const std::vector<Camera*>& cameras) {
  KF_ASSERT(!cameras.empty(), "Trying to render with an invalid camera!");
  const int w = cameras[0]->getWidth(), h = cameras[0]->getHeight();
  std::vector<CameraUBO> ubos;
  for (Camera* c : cameras) {
    KF_ASSERT(c, "Trying to render with an invalid camera!");
    KF_ASSERT(c->getWidth() == w && c->getHeight() == h, "Cameras rendered together must share one size");
    ubos.push_back(mCurrentScene->packCamera(c));
  }
  stashDisplacedFrames(cameras);
  KfrtPushConstants pc = makePushConstants(global::frameCount);
  const uint32_t s0 = mSharded ? std::min(mSampleBegin, pc.sampleRatePerPixel) : 0u;
  const uint32_t s1 = mSharded ? std::min(mSampleEnd, pc.sampleRatePerPixel) : pc.sampleRatePerPixel;
  check(kfrtRender(mRt, reinterpret_cast<const KfrtCamera*>(ubos.data()), uint32_t(ubos.size()), uint32_t(w),
                   uint32_t(h), &pc, s0, s1, mClockBase),
        "kfrtRender");
  if (!(mSharded && mDeferResolve)) check(kfrtResolve(mRt), "kfrtResolve");
  mClockBase += pc.sampleRatePerPixel + 1u;
  ++mSerial;
  for (size_t i = 0; i < cameras.size(); i++) {
    FrameStore& f = cameras[i]->mFrames;
    f.owner = this;
    f.slot = uint32_t(i);
    f.serial = mSerial;
    f.valid = true;
    f.stash.clear();
  }
  mLastCameras = cameras;
}

void Context::setSampleShard(uint32_t begin, uint32_t end, bool deferResolve) {
  KF_ASSERT(begin <= end, "sample shard must be an ordered range");
  mSharded = true;
  mSampleBegin = begin;
  mSampleEnd = end;
  mDeferResolve = deferResolve;
}
void Context::clearSampleShard() { mSharded = mDeferResolve = false; }
void Context::resolve() { check(kfrtResolve(mRt), "kfrtResolve"); }

void Context::render() {
  update();
  trace({mCurrentScene->getCamera()});
}

void Context::renderCameras(const std::vector<Camera*>& cameras) {
  update();
  trace(cameras);
}
}  // namespace kuafu
