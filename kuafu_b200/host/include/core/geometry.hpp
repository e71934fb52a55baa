// Geometry / material model of the kuafu API (reference include/core/geometry.hpp:11-108).
#pragma once
#include "core/context/vertex.hpp"

namespace kuafu {

/// Simplified Blender PrincipledBSDF parameters.
struct NiceMaterial {
  glm::vec3 diffuseColor = glm::vec3(1.0F);
  float alpha = 1.0F;

  std::string diffuseTexPath;
  std::string metallicTexPath;
  std::string roughnessTexPath;
  std::string transmissionTexPath;

  float metallic = 0.0F;
  float specular = 0.5F;
  float roughness = 0.5F;
  float ior = 1.4F;
  float transmission = 0.0F;

  glm::vec3 emission = glm::vec3(1.0F);
  float emissionStrength = 0.0F;

  friend bool operator==(const NiceMaterial& m1, const NiceMaterial& m2);
};

struct KUAFU_API Geometry {
  void setMaterial(const NiceMaterial& material);
  void recalculateNormals();

  std::vector<Vertex> vertices;
  std::vector<uint32_t> indices;
  std::vector<uint32_t> matIndex;  ///< one entry per index (the reference over-allocates the same way)
  std::string path;
  bool initialized = false;  ///< set once the geometry has been uploaded to the device
  bool dynamic = false;
  bool isOpaque = true;
  bool hideRender = false;
};

struct KUAFU_API GeometryInstance {
  void setTransform(const glm::mat4& transform);

  glm::mat4 transform = glm::mat4(1.0F);
  int geometryIndex = -1;
  std::shared_ptr<Geometry> geometry = nullptr;
};

KUAFU_API std::vector<std::shared_ptr<Geometry>> loadScene(std::string_view fname, bool dynamic);
KUAFU_API std::shared_ptr<Geometry> loadObj(std::string_view path, bool dynamic = false);

KUAFU_API std::shared_ptr<GeometryInstance> instance(const std::shared_ptr<Geometry>& geometry,
                                                     const glm::mat4& transform = glm::mat4(1.0F));

/// 80-byte material record read by the kernels (== KfrtMaterial).
struct NiceMaterialSSBO {
  glm::vec4 diffuse = glm::vec4(1.0F, 1.0F, 1.0F, 0.0F);
  glm::vec4 emission = glm::vec4(1.0F, 1.0F, 1.0F, 0.0F);
  float alpha = 1.0F;
  float metallic = 0.0F;
  float specular = 0.5F;
  float roughness = 0.5F;
  float ior = 1.4F;
  float transmission = 0.0F;
  int diffuseTexIdx = -1;
  int metallicTexIdx = -1;
  int roughnessTexIdx = -1;
  int transmissionTexIdx = -1;
  int padding0 = 0;
  int padding1 = 0;
};
static_assert(sizeof(NiceMaterialSSBO) == 80, "material wire format");

/// 80-byte instance record (== KfrtInstance).
struct GeometryInstanceSSBO {
  glm::mat4 transform = glm::mat4(1.0F);
  uint32_t geometryIndex = 0;
  uint32_t padding0 = 0;
  uint32_t padding1 = 0;
  uint32_t padding2 = 0;
};
static_assert(sizeof(GeometryInstanceSSBO) == 80, "instance wire format");

KUAFU_API std::shared_ptr<Geometry> createYZPlane(bool dynamic = true, NiceMaterial mat = {});
KUAFU_API std::shared_ptr<Geometry> createCube(bool dynamic = true, NiceMaterial mat = {});
KUAFU_API std::shared_ptr<Geometry> createSphere(bool dynamic = true, NiceMaterial mat = {});
KUAFU_API std::shared_ptr<Geometry> createCapsule(float halfHeight = 1., float radius = 1., bool dynamic = true,
                                                  NiceMaterial mat = {});
}  // namespace kuafu
