// Scene container of the kuafu API (reference include/core/scene.hpp:40-295, src/core/scene.cpp).
// Ownership follows the reference: geometries, instances and lights are shared_ptrs shared with the
// caller and mutated in place; cameras are owned by the scene and handed out as raw pointers.
#pragma once
#include "core/camera.hpp"
#include "core/config.hpp"
#include "core/light.hpp"

namespace kuafu {
class Context;
class Kuafu;

/// Host copy of everything the device needs, in wire format (what the reference keeps in its
/// Vulkan uniform/storage buffers).
struct WireScene {
  std::vector<NiceMaterialSSBO> materials;
  struct Tex {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba;
  };
  std::vector<Tex> textures;
  std::vector<GeometryInstanceSSBO> instances;
  DirectionalLightUBO directional{};
  PointLightsUBO points{};
  ActiveLightsUBO actives{};
  std::vector<uint8_t> envFaces[6];
  uint32_t envSize = 0;
};

class KUAFU_API Scene {
 public:
  friend Context;
  friend Kuafu;

  Scene() = delete;

  auto getGeometries() const -> const std::vector<std::shared_ptr<Geometry>>&;
  auto getGeometryInstances() const -> const std::vector<std::shared_ptr<GeometryInstance>>&;
  auto getGeometryInstance(size_t index) const -> std::shared_ptr<GeometryInstance>;

  void submitGeometryInstance(std::shared_ptr<GeometryInstance> geometryInstance);
  void submitGeometryInstance(const GeometryInstance& geometryInstance);
  void setGeometryInstances(const std::vector<std::shared_ptr<GeometryInstance>>& geometryInstances);
  void removeGeometryInstance(const std::shared_ptr<GeometryInstance>& geometryInstance);
  void removeGeometryInstances(const std::vector<std::shared_ptr<GeometryInstance>>& geometryInstances);
  void clearGeometryInstances();

  void submitGeometry(std::shared_ptr<Geometry> geometry);
  void submitGeometry(const Geometry& geometry);
  void setGeometries(const std::vector<std::shared_ptr<Geometry>>& geometries);
  void removeGeometry(std::shared_ptr<Geometry> geometry);
  void removeGeometry(uint32_t geometryIndex);
  void clearGeometries();

  void setClearColor(const glm::vec4& clearColor);
  inline auto getClearColor() const { return mClearColor; }

  auto findGeometry(std::string_view path) const -> std::shared_ptr<Geometry>;

  /// `.ktx` (KTX1, uncompressed RGBA8 cube) like the reference; anything else throws.
  void setEnvironmentMap(std::string_view path);
  /// Additive: six square RGBA8 faces in Vulkan cube order (+X,-X,+Y,-Y,+Z,-Z).
  void setEnvironmentMapFaces(const uint8_t* const faces[6], uint32_t size);
  void removeEnvironmentMap();

  Camera* createCamera(int width, int height);
  void removeCamera(Camera* camera);
  void setCamera(Camera* camera);
  Camera* getCamera() const { return mCurrentCamera; }

  inline auto getGeometryInstanceCount() { return mGeometries.size(); }

  inline void markGeometriesChanged() { mUploadGeometries = true; }
  inline void markGeometryInstancesChanged() { mUploadGeometryInstancesToBuffer = true; }

  inline void setDirectionalLight(std::shared_ptr<DirectionalLight> light) { pDirectionalLight = light; }
  inline void removeDirectionalLight() { pDirectionalLight = nullptr; }

  void addPointLight(const std::shared_ptr<PointLight>& light);
  void removePointLight(const std::shared_ptr<PointLight>& light);
  void addActiveLight(const std::shared_ptr<ActiveLight>& light);
  void removeActiveLight(const std::shared_ptr<ActiveLight>& light);

  void init();

  /// Additive: the packed wire buffers of the last Context::update() (test / tooling hook).
  const WireScene& wire() const { return mWire; }
  CameraUBO packCamera(const Camera* camera) const;

 private:
  explicit Scene(std::shared_ptr<Config> pConfig);

  void packLights();
  void packMaterialsAndTextures();
  void packInstances();

  bool initialized = false;
  glm::vec4 mClearColor = glm::vec4(0.F, 0.F, 0.F, 1.F);

  std::vector<std::shared_ptr<Geometry>> mGeometries;
  std::vector<std::shared_ptr<GeometryInstance>> mGeometryInstances;

  std::shared_ptr<DirectionalLight> pDirectionalLight;
  std::vector<std::shared_ptr<PointLight>> pPointLights;
  std::vector<std::shared_ptr<ActiveLight>> pActiveLights;

  std::string mEnvironmentMapTexturePath;
  bool mUseEnvironmentMap = false;

  bool mUploadGeometryInstancesToBuffer = false;
  bool mUploadEnvironmentMap = false;
  bool mUploadGeometries = false;

  std::vector<std::unique_ptr<Camera>> mRegisteredCameras;
  Camera* mCurrentCamera = nullptr;
  std::shared_ptr<Config> pConfig = nullptr;

  WireScene mWire;
};
}  // namespace kuafu
