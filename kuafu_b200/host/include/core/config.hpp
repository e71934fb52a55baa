// Renderer settings (reference include/core/config.hpp:35-177, src/core/config.cpp).  Values reach
// the kernels through the push-constant block built every frame (KfrtPushConstants).
#pragma once
#include "stdafx.hpp"

namespace kuafu {
class Context;
class Kuafu;
class Scene;

class KUAFU_API Config {
 public:
  friend class Context;
  friend class Kuafu;
  friend class Scene;

  auto getPathDepth() const -> uint32_t { return mPathDepth; }
  void setPathDepth(uint32_t recursionDepth);

  bool getRussianRoulette() { return mRussianRoulette; }
  void setRussianRoulette(bool flag);

  bool getNextEventEstimation() { return mNextEventEstimation; }
  void setNextEventEstimation(bool flag);

  uint32_t getNextEventEstimationMinBounces() { return mNextEventEstimationMinBounces; }
  void setNextEventEstimationMinBounces(uint32_t minBounces);

  uint32_t getRussianRouletteMinBounces() { return mRussianRouletteMinBounces; }
  void setRussianRouletteMinBounces(uint32_t minBounces);

  auto getMaxPathDepth() const -> uint32_t { return mMaxPathDepth; }

  static std::string sDefaultAssetsPath;
  static void setDefaultAssetsPath(std::string path);

  auto getAssetsPath() const -> std::string_view { return mAssetsPath; }
  void setAssetsPath(int argc, char* argv[]);
  void setAssetsPath(std::string_view path);

  void setAutomaticPipelineRefresh(bool flag);

  void setGeometryLimit(size_t amount);
  void setGeometryInstanceLimit(uint32_t amount);
  void setTextureLimit(size_t amount);
  void setMaterialLimit(size_t amount);

  void setPerPixelSampleRate(uint32_t sampleRate);
  auto getPerPixelSampleRate() const -> uint32_t { return mPerPixelSampleRate; }

  void setUseDenoiser(bool useDenoiser = true);
  auto isUsingDenoiser() const -> bool { return mUseDenoiser; }

  void setAccumulatingFrames(bool flag);
  auto isAccumulatingFrames() const -> bool { return mAccumulateFrames; }

  void triggerPipelineRefresh() { mPipelineNeedsRefresh = true; }
  void triggerSwapchainRefresh() { mSwapchainNeedsRefresh = true; }

  float getVariance() { return mVariance; }
  void updateVariance(bool flag);

  inline void setPresent(bool present) { mPresent = present; }
  inline bool getPresent() { return mPresent; }

  inline void setInitialWidth(int w) { mInitialWidth = w; }
  inline void setInitialHeight(int h) { mInitialHeight = h; }

  /// Additive: CUDA device ordinal the renderer runs on (one process per GPU).
  inline void setDeviceOrdinal(int d) { mDeviceOrdinal = d; }
  inline int getDeviceOrdinal() const { return mDeviceOrdinal; }
  /// Additive: build and pack scenes without touching a device (tooling); rendering then throws.
  inline void setHostOnly(bool flag) { mHostOnly = flag; }
  inline bool isHostOnly() const { return mHostOnly; }

 private:
  int mInitialWidth = 800;
  int mInitialHeight = 600;
  int mDeviceOrdinal = 0;
  bool mHostOnly = false;

  bool mPresent = true;  ///< the reference opens a window when true; this build is offscreen-only
  bool mUseDenoiser = false;

  bool mPipelineNeedsRefresh = false;
  bool mSwapchainNeedsRefresh = false;

  size_t mMaxGeometryInstances = 256;
  bool mMaxGeometryInstancesChanged = false;
  size_t mMaxGeometry = 128;
  bool mMaxGeometryChanged = false;
  size_t mMaxTextures = 128;
  bool mMaxTexturesChanged = false;
  size_t mMaxMaterials = 256;

  std::string mAssetsPath;

  uint32_t mMaxPathDepth = 12;
  uint32_t mPathDepth = 8;
  uint32_t mPerPixelSampleRate = 32;
  uint32_t mRussianRouletteMinBounces = 4;

  bool mNextEventEstimation = true;             // pushed, never read by the reference shaders
  uint32_t mNextEventEstimationMinBounces = 0;  // pushed, never read by the reference shaders

  float mVariance = 0.0F;
  bool mUpdateVariance = false;

  bool mAccumulateFrames = true;
  bool mRussianRoulette = true;
  bool mAutomaticPipelineRefresh = false;
  bool mAutomaticSwapchainRefresh = false;
};
}  // namespace kuafu
