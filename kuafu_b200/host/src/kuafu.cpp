// Public facade (reference src/kuafu.cpp:27-129).  Offscreen only: `setPresent(true)` is accepted
// but no window is opened -- a B200 node has no display.
#include "kuafu.hpp"

namespace kuafu {
Kuafu::Kuafu(std::shared_ptr<Config> config) {
  mContext.pConfig = config ? std::move(config) : std::make_shared<Config>();
  if (mContext.pConfig->getAssetsPath().empty() && !Config::sDefaultAssetsPath.empty())
    mContext.pConfig->setAssetsPath(Config::sDefaultAssetsPath);
  if (mContext.pConfig->getPresent()) {
    KF_INFO("Present mode requested; this build renders offscreen only.");
    mContext.pConfig->setPresent(false);
  }
  Scene* scene = createScene();
  mContext.mCurrentScene = scene;
  Camera* cam = scene->createCamera(mContext.pConfig->mInitialWidth, mContext.pConfig->mInitialHeight);
  scene->setCamera(cam);
  mContext.init();
  reset();
}

bool Kuafu::isRunning() const {
  if (!mRunning) KF_INFO("Shutting down Kuafu.");
  return mRunning;
}

void Kuafu::run() {
  if (!mRunning) return;
  mContext.getCamera()->update();
  mContext.render();
}

void Kuafu::run(const std::vector<Camera*>& cameras) {
  if (!mRunning) return;
  for (Camera* c : cameras)
    if (c) c->update();
  mContext.renderCameras(cameras);
}

std::pair<size_t, size_t> Kuafu::cameraShard(size_t nCameras, int rank, int world) {
  KF_ASSERT(world > 0 && rank >= 0 && rank < world, "cameraShard: rank must lie in [0, world)");
  const size_t r = size_t(rank), w = size_t(world);
  return {nCameras * r / w, nCameras * (r + 1) / w};
}

std::vector<size_t> Kuafu::cameraShardIndices(size_t nCameras, int rank, int world, bool interleaved) {
  KF_ASSERT(world > 0 && rank >= 0 && rank < world, "cameraShardIndices: rank must lie in [0, world)");
  std::vector<size_t> out;
  if (interleaved) {
    for (size_t c = size_t(rank); c < nCameras; c += size_t(world)) out.push_back(c);
  } else {
    const auto range = cameraShard(nCameras, rank, world);
    for (size_t c = range.first; c < range.second; c++) out.push_back(c);
  }
  return out;
}

std::vector<uint8_t> Kuafu::downloadLatestFrame(Camera* cam) {
  KF_ASSERT(cam, "Invalid call to Camera::downloadLatestFrame");
  return cam->downloadLatestFrame();
}

void Kuafu::setWindow(std::shared_ptr<Window> window) {
  pWindow = window;
  mContext.pWindow = pWindow;
}
void Kuafu::setWindow(int width, int height, const char* title, uint32_t flags) {
  setWindow(std::make_shared<Window>(width, height, title, flags));
}
void Kuafu::setGui(std::shared_ptr<Gui> gui) { pGUI = std::move(gui); }

void Kuafu::reset() { global::frameCount = -1; }

void Kuafu::setScene(Scene* scene) {
  if (mContext.mCurrentScene == scene && scene->initialized) return;
  KF_INFO("Switching scene, this is heavy...");
  KF_ASSERT(scene, "Trying to set an invalid scene!");
  auto it = std::find_if(mContext.mScenes.begin(), mContext.mScenes.end(), [scene](auto& s) { return s.get() == scene; });
  KF_ASSERT(it != mContext.mScenes.end(), "Scene does not belong to this renderer");
  scene->init();
  mContext.mCurrentScene = scene;
  global::frameCount = -1;
}

Scene* Kuafu::createScene() {
  mContext.mScenes.emplace_back(new Scene(mContext.pConfig));
  return mContext.mScenes.back().get();
}

void Kuafu::removeScene(Scene* scene) {
  KF_ASSERT(scene, "Trying to remove an invalid scene!");
  auto it = std::find_if(mContext.mScenes.begin(), mContext.mScenes.end(), [scene](auto& s) { return s.get() == scene; });
  KF_ASSERT(it != mContext.mScenes.end(), "Scene does not belong to this renderer");
  mContext.forgetScene(scene);
  mContext.mScenes.erase(it);
}
}  // namespace kuafu
