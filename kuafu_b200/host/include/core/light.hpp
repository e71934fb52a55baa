// Light descriptions and their uniform-buffer mirrors (reference include/core/light.hpp:9-58).
#pragma once
#include "core/context/global.hpp"

namespace kuafu {

struct DirectionalLight {
  glm::vec3 direction = {0, 0, -1};
  glm::vec3 color = {1, 1, 1};
  float strength = 0.0;
  float softness = 0.0;
};

struct PointLight {
  glm::vec3 position = {0, 0, 0};
  glm::vec3 color = {1, 1, 1};
  float radius = 1;
  float strength = 0.0;
};

/// Textured projector.  Without a texture (or when it fails to load) it degrades to a spot light.
struct ActiveLight {
  glm::mat4 viewMat{};
  glm::vec3 color = {1, 1, 1};
  float fov = 0.0;
  float strength = 0.0;
  float softness = 0.0;
  std::string texPath;
  int texID = -1;
};

struct DirectionalLightUBO {
  glm::vec4 direction{};  // direction + softness
  glm::vec4 rgbs{};       // rgb + strength
};

struct PointLightsUBO {
  glm::vec4 posr[global::maxPointLights]{};  // position + radius
  glm::vec4 rgbs[global::maxPointLights]{};  // rgb + strength
};

struct ActiveLightsUBO {
  glm::mat4 viewMat[global::maxActiveLights]{};
  glm::mat4 projMat[global::maxActiveLights]{};
  glm::vec4 front[global::maxActiveLights]{};     // front + in-use flag
  glm::vec4 rgbs[global::maxActiveLights]{};      // rgb + strength
  glm::vec4 position[global::maxActiveLights]{};  // position
  glm::vec4 sftp[global::maxActiveLights]{};      // softness, fov, texID, padding
};
static_assert(sizeof(DirectionalLightUBO) == 32 && sizeof(PointLightsUBO) == 1024 &&
                  sizeof(ActiveLightsUBO) == 1536,
              "light wire formats");
}  // namespace kuafu
