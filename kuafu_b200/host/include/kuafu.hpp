// kuafu.hpp -- public facade of the B200-native Kuafu renderer.  Same surface as the reference's
// include/kuafu.hpp:31-93 (Kuafu, Scene, Camera, Geometry, NiceMaterial, lights, Config,
// downloadLatestFrame), so SAPIEN-side code compiles unchanged; underneath, the Vulkan-RT pipeline
// is replaced by the CUDA core behind include/kf_rt.h.
#pragma once

#include "core/context/context.hpp"
#include "core/gui.hpp"
#include "core/window.hpp"
#include "stdafx.hpp"

namespace kuafu {
class KUAFU_API Kuafu {
  std::shared_ptr<Window> pWindow = nullptr;
  std::shared_ptr<Gui> pGUI = nullptr;
  kuafu::Context mContext;
  bool mRunning = true;

  void reset();

 public:
  explicit Kuafu(std::shared_ptr<Config> config = nullptr);

  void run();
  /// Additive: render several cameras of the current scene in one launch (one TLAS refit).
  void run(const std::vector<Camera*>& cameras);
  /// Additive (camera-batch split): the cameras process `rank` of `world` renders when a
  /// batch of `nCameras` is split across GPUs -- a contiguous range [first, second), sizes differing by
  /// at most one.  Every process holds the whole scene and refits its own top level once per run();
  /// no collective is involved.
  static std::pair<size_t, size_t> cameraShard(size_t nCameras, int rank, int world);
  /// The same split as a list of camera indices; interleaved: rank, rank + world, rank + 2 world, ...
  /// instead of a contiguous range.  Neighbouring cameras of a rig see similar amounts of scene, so an
  /// interleaved split balances the ranks where a contiguous one leaves the busiest views on one GPU.
  static std::vector<size_t> cameraShardIndices(size_t nCameras, int rank, int world, bool interleaved);

  [[nodiscard]] bool isRunning() const;

  [[nodiscard]] std::vector<uint8_t> downloadLatestFrame(Camera* cam);

  void setWindow(std::shared_ptr<Window> other);
  void setWindow(int width, int height, const char* title = "App", uint32_t flags = 0);
  [[nodiscard]] auto getWindow() const { return pWindow; }
  void setGui(std::shared_ptr<Gui> gui);

  inline auto& getConfig() { return *mContext.pConfig; }
  inline Scene* getScene() { return mContext.mCurrentScene; }
  inline Context& getContext() { return mContext; }

  void setScene(Scene* scene);
  Scene* createScene();
  void removeScene(Scene* scene);
};
}  // namespace kuafu
