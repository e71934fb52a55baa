// Vertex wire format: 48 bytes, identical to reference include/core/context/vertex.hpp:28-33 and to
// KfrtVertex (include/kf_rt.h).  The Vulkan binding/attribute descriptions of the reference have no
// counterpart here.
#pragma once
#include "stdafx.hpp"

namespace kuafu {
struct Vertex {
  glm::vec3 pos{0.0f};
  glm::vec3 normal{0.0f};
  glm::vec3 color{0.0f};
  glm::vec2 texCoord{0.0f};
  float padding0 = 0.0f;

  bool operator==(const Vertex& o) const {
    return pos == o.pos && color == o.color && texCoord == o.texCoord && normal == o.normal;
  }
};
static_assert(sizeof(Vertex) == 48, "Vertex must stay 48 bytes (shader/kernel wire format)");
}  // namespace kuafu
