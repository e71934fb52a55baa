// Config setters (reference src/core/config.cpp:25-155), including its quirks: setPathDepth() warns
// above the maximum but still stores the requested value, limits are stored as amount + 1.
#include "core/config.hpp"

#include "core/context/global.hpp"

namespace kuafu {
std::string Config::sDefaultAssetsPath;

void Config::setDefaultAssetsPath(std::string p) { sDefaultAssetsPath = std::move(p); }

void Config::setPathDepth(uint32_t recursionDepth) {
  if (recursionDepth > mMaxPathDepth)
    KF_WARN("Exceeded maximum path depth of ", mMaxPathDepth, ". Using highest possible value instead.");
  mPathDepth = recursionDepth;  // the reference overwrites its clamp (config.cpp:31-41)
}

void Config::setNextEventEstimation(bool flag) { mNextEventEstimation = flag; }
void Config::setNextEventEstimationMinBounces(uint32_t n) { mNextEventEstimationMinBounces = n; }
void Config::setRussianRoulette(bool flag) { mRussianRoulette = flag; }
void Config::setRussianRouletteMinBounces(uint32_t n) { mRussianRouletteMinBounces = n; }

static std::string withForwardSlashes(std::string s) {
  std::replace(s.begin(), s.end(), '\\', '/');
  return s;
}

void Config::setAssetsPath(int argc, char* argv[]) {
  std::string joined;
  for (int i = 0; i < argc; ++i) joined += argv[i];
  joined = withForwardSlashes(joined);
  mAssetsPath = joined.substr(0, joined.find_last_of('/') + 1);
  global::assetsPath = mAssetsPath;
}

void Config::setAssetsPath(std::string_view path) {
  if (path.empty()) return;
  mAssetsPath = withForwardSlashes(std::string(path));
  if (mAssetsPath.back() != '/') mAssetsPath += '/';
  global::assetsPath = mAssetsPath;
}

void Config::setAutomaticPipelineRefresh(bool flag) { mAutomaticPipelineRefresh = flag; }

void Config::setGeometryInstanceLimit(uint32_t amount) {
  if (amount == 0) {
    amount = 1;
    KF_WARN("Can not use value 0 for the maximum amount of geometry instances. Using 1 instead.");
  }
  mMaxGeometryInstances = amount;
  mMaxGeometryInstancesChanged = true;
}

void Config::setGeometryLimit(size_t amount) {
  if (amount < 16 || amount % 4 != 0) amount = 16;
  mMaxGeometry = amount + 1;
  mMaxGeometryChanged = true;
}

void Config::setTextureLimit(size_t amount) {
  if (amount == 0) KF_WARN("Can not use value 0 for the maximum amount of textures. Using 1 instead.");
  mMaxTextures = amount + 1;
  mMaxTexturesChanged = true;
}

void Config::setMaterialLimit(size_t amount) {
  if (amount == 0) KF_WARN("Can not use value 0 for the maximum amount of materials. Using 1 instead.");
  mMaxMaterials = amount + 1;
}

void Config::setUseDenoiser(bool useDenoiser) {
  if (useDenoiser) KF_WARN("The OptiX denoiser is out of scope of this build; the flag is stored but ignored.");
  mUseDenoiser = useDenoiser;
}

void Config::setPerPixelSampleRate(uint32_t sampleRate) { mPerPixelSampleRate = sampleRate; }
void Config::setAccumulatingFrames(bool flag) { mAccumulateFrames = flag; }
void Config::updateVariance(bool flag) { mUpdateVariance = flag; }
}  // namespace kuafu
