// Frame orchestration (reference include/core/context/context.hpp, src/core/context/context.cpp:
// 298-350 update(), :602-684 recordSwapchainCommandBuffers()).  Same shape -- dirty flags -> upload ->
// build or refit -> uniforms -> frameCount++ -> trace -> output stage -- but every device action
// goes through the kf_rt.h C ABI instead of Vulkan.
#pragma once
#include "core/scene.hpp"

struct KfrtContext;
struct KfrtPushConstants;

namespace kuafu {
class Kuafu;
class Window;
class Gui;

class KUAFU_API Context {
 public:
  friend Kuafu;
  friend Camera;
  friend Scene;

  Context() = default;
  ~Context();
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;

  void init();
  void update();
  void render();
  /// Additive: one update(), then all cameras (same size) traced in one launch.
  void renderCameras(const std::vector<Camera*>& cameras);

  /// Additive: host-side packing of the current scene into wire format, no device work.
  void pack();
  /// The push-constant block a frame with the given frameCount uses (reference context.cpp:605-614).
  KfrtPushConstants makePushConstants(int frameCount) const;
  /// frameCount the next update() will produce.
  int predictFrameCount() const;

  /// Additive (spp sharding): this process traces only samples [begin, end) of every
  /// pixel; with deferResolve the frame is finished by resolve() after the caller has summed the
  /// SUM32F buffers of all ranks (kfrtReduceNccl or its own collective on kfrtGetDeviceBuffer).
  void setSampleShard(uint32_t begin, uint32_t end, bool deferResolve);
  void clearSampleShard();
  void resolve();

  Camera* getCamera() { return mCurrentScene ? mCurrentScene->getCamera() : nullptr; }
  KfrtContext* getDevice() { return mRt; }
  /// Surrogate of the reference's clockARB() seeds: advances by spp + 1 per rendered frame.
  uint32_t getClockBase() const { return mClockBase; }
  void setClockBase(uint32_t c) { mClockBase = c; }

  std::shared_ptr<Config> pConfig = nullptr;
  std::shared_ptr<Window> pWindow = nullptr;
  std::vector<std::unique_ptr<Scene>> mScenes;
  Scene* mCurrentScene = nullptr;

 private:
  void check(int rc, const char* what);
  void trace(const std::vector<Camera*>& cameras);
  void stashDisplacedFrames(const std::vector<Camera*>& next);
  void forgetCamera(Camera* camera);
  void forgetScene(Scene* scene);

  KfrtContext* mRt = nullptr;
  uint32_t mClockBase = 0;
  uint64_t mSerial = 0;
  std::vector<Camera*> mLastCameras;
  std::vector<float> mLastTransforms;
  std::vector<uint8_t> mLastLights;  // the light blocks last uploaded
  Scene* mUploadedScene = nullptr;
  bool mSharded = false, mDeferResolve = false;
  uint32_t mSampleBegin = 0, mSampleEnd = 0;
};
}  // namespace kuafu
