// Process-wide state of the kuafu API (reference include/core/context/global.hpp:27-39,
// src/core/context/global.cpp).  The material registry and frameCount determine results, so their
// semantics are kept: materials are appended by Geometry creation and indexed by matIndex.
#pragma once
#include "core/geometry.hpp"

namespace kuafu::global {
KUAFU_API extern std::shared_ptr<Logger> logger;

KUAFU_API extern int frameCount;

KUAFU_API extern std::string assetsPath;
KUAFU_API extern uint32_t materialIndex;
KUAFU_API extern uint32_t textureIndex;
KUAFU_API extern std::vector<NiceMaterial> materials;

const size_t maxResources = 2;
const size_t maxPointLights = 32;
const size_t maxActiveLights = 8;

namespace keys {
KUAFU_API extern bool eW, eA, eS, eD, eQ, eE, eX, eY, eZ, eC, eSpace, eLeftShift, eLeftCtrl, eB, eL;
}

/// Additive: textures can be handed over in memory instead of through a file.  The returned name
/// ("mem:<name>") is accepted wherever the API takes a texture path.
KUAFU_API std::string registerMemoryTexture(const std::string& name, uint32_t width, uint32_t height,
                                            const uint8_t* rgba8);
struct MemoryTexture {
  uint32_t width = 0, height = 0;
  std::vector<uint8_t> rgba;
};
KUAFU_API const MemoryTexture* findMemoryTexture(const std::string& path);
}  // namespace kuafu::global
