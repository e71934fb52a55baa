// Process-wide state (reference src/core/context/global.cpp:26-54).
#include "core/context/global.hpp"

#include <map>

namespace kuafu::global {
std::shared_ptr<Logger> logger = std::make_shared<Logger>();
int frameCount = -1;
std::string assetsPath;
uint32_t materialIndex = 0;
uint32_t textureIndex = 0;
std::vector<NiceMaterial> materials;

namespace keys {
bool eW = false, eA = false, eS = false, eD = false, eQ = false, eE = false, eX = false, eY = false,
     eZ = false, eC = false, eSpace = false, eLeftShift = false, eLeftCtrl = false, eB = false, eL = false;
}

static std::map<std::string, MemoryTexture>& memoryTextures() {
  static std::map<std::string, MemoryTexture> m;
  return m;
}

std::string registerMemoryTexture(const std::string& name, uint32_t width, uint32_t height,
                                  const uint8_t* rgba8) {
  MemoryTexture t;
  t.width = width;
  t.height = height;
  t.rgba.assign(rgba8, rgba8 + size_t(width) * height * 4);
  const std::string key = "mem:" + name;
  memoryTextures()[key] = std::move(t);
  return key;
}

const MemoryTexture* findMemoryTexture(const std::string& path) {
  auto it = memoryTextures().find(path);
  return it == memoryTextures().end() ? nullptr : &it->second;
}
}  // namespace kuafu::global
