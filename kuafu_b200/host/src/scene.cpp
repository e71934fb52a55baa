// Scene container and the packing of its contents into wire format (reference src/core/scene.cpp:
// 58-230 container logic, :236-292 uniform packing, :311-467 material / texture / instance packing).
#include "core/scene.hpp"

#include "core/context/context.hpp"
#include "image_io.hpp"

namespace kuafu {

Scene::Scene(std::shared_ptr<Config> config) : pConfig(std::move(config)) { setCamera(createCamera(1, 1)); }

auto Scene::getGeometries() const -> const std::vector<std::shared_ptr<Geometry>>& { return mGeometries; }
auto Scene::getGeometryInstances() const -> const std::vector<std::shared_ptr<GeometryInstance>>& {
  return mGeometryInstances;
}
auto Scene::getGeometryInstance(size_t index) const -> std::shared_ptr<GeometryInstance> {
  if (index >= mGeometryInstances.size()) throw std::runtime_error("Geometry Instances out of bound.");
  return mGeometryInstances[index];
}

void Scene::submitGeometryInstance(std::shared_ptr<GeometryInstance> inst) {
  if (mGeometryInstances.size() > pConfig->mMaxGeometryInstances)
    throw std::runtime_error("Failed to submit geometry instance because instance buffer size has been exceeded.");
  for (size_t i = 0; i < mGeometries.size(); i++)
    if (mGeometries[i] == inst->geometry) inst->geometryIndex = int(i);
  KF_ASSERT(inst->geometryIndex >= 0, "Geometry not submitted!");
  mGeometryInstances.push_back(std::move(inst));
  markGeometryInstancesChanged();
}
void Scene::submitGeometryInstance(const GeometryInstance& inst) {
  submitGeometryInstance(std::make_shared<GeometryInstance>(inst));
}
void Scene::setGeometryInstances(const std::vector<std::shared_ptr<GeometryInstance>>& insts) {
  mGeometryInstances.clear();
  mGeometryInstances.reserve(insts.size());
  for (const auto& i : insts) submitGeometryInstance(i);
  markGeometryInstancesChanged();
}
void Scene::removeGeometryInstance(const std::shared_ptr<GeometryInstance>& inst) {
  KF_ASSERT(inst, "Deleting an invalid geometry instance!");
  mGeometryInstances.erase(std::remove(mGeometryInstances.begin(), mGeometryInstances.end(), inst),
                           mGeometryInstances.end());
  markGeometryInstancesChanged();
}
void Scene::removeGeometryInstances(const std::vector<std::shared_ptr<GeometryInstance>>& insts) {
  for (const auto& i : insts) removeGeometryInstance(i);
}
void Scene::clearGeometryInstances() {
  mGeometryInstances.clear();
  markGeometryInstancesChanged();
}

void Scene::submitGeometry(std::shared_ptr<Geometry> geometry) {
  if (mGeometries.size() >= pConfig->mMaxGeometry)
    throw std::runtime_error("Failed to submit geometry because geometries buffer size has been exceeded.");
  mGeometries.push_back(std::move(geometry));
  markGeometriesChanged();
}
void Scene::submitGeometry(const Geometry& geometry) { submitGeometry(std::make_shared<Geometry>(geometry)); }
void Scene::setGeometries(const std::vector<std::shared_ptr<Geometry>>& geometries) {
  mGeometries.clear();
  mGeometries.reserve(geometries.size());
  for (const auto& g : geometries) submitGeometry(g);
  markGeometriesChanged();
}
void Scene::removeGeometry(std::shared_ptr<Geometry>) { throw std::runtime_error("Not implemented!"); }
void Scene::removeGeometry(uint32_t) { throw std::runtime_error("Not implemented!"); }
void Scene::clearGeometries() {
  mGeometries.clear();
  mGeometryInstances.clear();
  markGeometriesChanged();
  markGeometryInstancesChanged();
}

void Scene::setClearColor(const glm::vec4& clearColor) {
  mClearColor = clearColor;
  global::frameCount = -1;
}

auto Scene::findGeometry(std::string_view path) const -> std::shared_ptr<Geometry> {
  for (const auto& g : mGeometries)
    if (g && g->path == path) return g;
  return nullptr;
}

void Scene::setEnvironmentMap(std::string_view path) {
  mEnvironmentMapTexturePath = path;
  mUseEnvironmentMap = true;
  mUploadEnvironmentMap = true;
}
void Scene::setEnvironmentMapFaces(const uint8_t* const faces[6], uint32_t size) {
  for (int f = 0; f < 6; f++) mWire.envFaces[f].assign(faces[f], faces[f] + size_t(size) * size * 4);
  mWire.envSize = size;
  mEnvironmentMapTexturePath = "mem:";
  mUseEnvironmentMap = true;
  mUploadEnvironmentMap = true;
}
void Scene::removeEnvironmentMap() {
  mUseEnvironmentMap = false;
  global::frameCount = -1;
}

Camera* Scene::createCamera(int width, int height) {
  mRegisteredCameras.emplace_back(new Camera(width, height, glm::vec3(0.f, 0.f, 0.f)));
  return mRegisteredCameras.back().get();
}
void Scene::removeCamera(Camera* camera) {
  KF_ASSERT(camera, "Trying to remove an invalid camera!");
  auto it = std::find_if(mRegisteredCameras.begin(), mRegisteredCameras.end(),
                         [camera](auto& c) { return c.get() == camera; });
  KF_ASSERT(it != mRegisteredCameras.end(), "Trying to remove an camera which does not belong to the scene!");
  if (camera->mFrames.owner) camera->mFrames.owner->forgetCamera(camera);
  if (mCurrentCamera == camera) {
    KF_INFO("Removing the active camera. This may cause problems.");
    mCurrentCamera = nullptr;
  }
  mRegisteredCameras.erase(it);
}
void Scene::setCamera(Camera* camera) {
  KF_ASSERT(camera, "Trying to set an invalid camera!");
  auto it = std::find_if(mRegisteredCameras.begin(), mRegisteredCameras.end(),
                         [camera](auto& c) { return c.get() == camera; });
  KF_ASSERT(it != mRegisteredCameras.end(), "Trying to set an camera which does not belong to the scene!");
  mCurrentCamera = camera;
  global::frameCount = -1;
}

void Scene::addPointLight(const std::shared_ptr<PointLight>& light) {
  if (pPointLights.size() >= global::maxPointLights)
    KF_WARN("Reached max point light number. The light will not be added!");
  else
    pPointLights.push_back(light);
}
void Scene::removePointLight(const std::shared_ptr<PointLight>& light) {
  KF_ASSERT(light, "Deleting an invalid light!");
  pPointLights.erase(std::remove(pPointLights.begin(), pPointLights.end(), light), pPointLights.end());
}
void Scene::addActiveLight(const std::shared_ptr<ActiveLight>& light) {
  if (pActiveLights.size() >= global::maxActiveLights)
    KF_WARN("Reached max active light number. The light will not be added!");
  else {
    pActiveLights.push_back(light);
    markGeometriesChanged();  // its texture is uploaded together with the material textures
  }
}
void Scene::removeActiveLight(const std::shared_ptr<ActiveLight>& light) {
  KF_ASSERT(light, "Deleting an invalid light!");
  pActiveLights.erase(std::remove(pActiveLights.begin(), pActiveLights.end(), light), pActiveLights.end());
}

void Scene::init() {
  markGeometriesChanged();
  markGeometryInstancesChanged();
  for (auto& g : mGeometries)
    if (g) g->initialized = false;
  mUploadEnvironmentMap = mUseEnvironmentMap;
  initialized = true;
}

// ---- packing ---------------------------------------------------------------------------------

CameraUBO Scene::packCamera(const Camera* cam) const {  // reference scene.cpp:236-250
  KF_ASSERT(cam, "Trying to render with an invalid camera!");
  CameraUBO u;
  u.view = cam->getViewMatrix();
  u.viewInverse = cam->getViewInverseMatrix();
  u.projection = cam->getProjectionMatrix();
  u.projectionInverse = cam->getProjectionInverseMatrix();
  u.position = glm::vec4(cam->getPosition(), cam->getAperture());
  u.front = glm::vec4(cam->getFront(), cam->getFocalLength());
  return u;
}

void Scene::packLights() {  // reference scene.cpp:252-292
  if (pDirectionalLight) {
    mWire.directional.direction = glm::vec4(glm::normalize(pDirectionalLight->direction), pDirectionalLight->softness);
    mWire.directional.rgbs = glm::vec4(pDirectionalLight->color, pDirectionalLight->strength);
  } else {
    mWire.directional.rgbs[3] = 0.0f;
  }
  for (size_t i = 0; i < global::maxPointLights; i++) {
    if (i < pPointLights.size()) {
      KF_ASSERT(pPointLights[i], "Invalid point light!");
      mWire.points.posr[i] = glm::vec4(pPointLights[i]->position, pPointLights[i]->radius);
      mWire.points.rgbs[i] = glm::vec4(pPointLights[i]->color, pPointLights[i]->strength);
    } else {
      mWire.points.rgbs[i][3] = 0.0f;
    }
  }
  for (size_t i = 0; i < global::maxActiveLights; i++) {
    if (i < pActiveLights.size()) {
      KF_ASSERT(pActiveLights[i], "Invalid active light!");
      const ActiveLight& l = *pActiveLights[i];
      const glm::mat4 inv = glm::inverse(l.viewMat);
      mWire.actives.viewMat[i] = l.viewMat;
      mWire.actives.projMat[i] = glm::perspective(l.fov, 1.F, 0.01F, 1000.0F);
      mWire.actives.front[i] = glm::vec4(-inv[2][0], -inv[2][1], -inv[2][2], 1.0f);
      mWire.actives.rgbs[i] = glm::vec4(l.color, l.strength);
      mWire.actives.position[i] = glm::vec4(inv[3][0], inv[3][1], inv[3][2], 0.0f);
      mWire.actives.sftp[i] = glm::vec4(l.softness, l.fov, float(l.texID), 0.0f);
      if (l.softness > 0) KF_WARN("FIXME: softness of active light is incorrectly implemented!");
    } else {
      mWire.actives.front[i][3] = 0.0f;
    }
  }
}

// Materials of the process-wide registry -> 80-byte records, loading every referenced texture
// (reference scene.cpp:314-413: four optional textures per material, then the projector patterns).
void Scene::packMaterialsAndTextures() {
  mWire.materials.clear();
  mWire.textures.clear();
  global::textureIndex = 0;
  auto load = [&](const std::string& path, const char* what, const char* fallback) -> int {
    if (path.empty()) return -1;
    WireScene::Tex t;
    if (!io::loadTextureRGBA8(path, t.width, t.height, t.rgba)) {
      KF_WARN("Failed to load ", what, " texture: ", path, ", ", fallback, " will be used!");
      return -1;
    }
    if (mWire.textures.size() >= pConfig->mMaxTextures) throw std::runtime_error("Texture limit exceeded.");
    mWire.textures.push_back(std::move(t));
    return int(global::textureIndex++);
  };
  for (const NiceMaterial& m : global::materials) {
    NiceMaterialSSBO s;
    s.diffuse = glm::vec4(m.diffuseColor, 0.0f);
    s.emission = glm::vec4(m.emission, m.emissionStrength);
    s.alpha = m.alpha;
    s.metallic = m.metallic;
    s.specular = m.specular;
    s.roughness = m.roughness;
    s.ior = m.ior;
    s.transmission = m.transmission;
    s.diffuseTexIdx = load(m.diffuseTexPath, "diffuse", "base color");
    s.metallicTexIdx = load(m.metallicTexPath, "metallic", "metallic value");
    s.roughnessTexIdx = load(m.roughnessTexPath, "roughness", "roughness value");
    s.transmissionTexIdx = load(m.transmissionTexPath, "transmission", "transmission value");
    mWire.materials.push_back(s);
  }
  for (auto& light : pActiveLights) {
    light->texID = load(light->texPath, "active light", "spot light");
  }
}

void Scene::packInstances() {  // reference scene.cpp:444-465
  mWire.instances.resize(mGeometryInstances.size());
  for (size_t i = 0; i < mGeometryInstances.size(); i++) {
    mWire.instances[i].transform = mGeometryInstances[i]->transform;
    mWire.instances[i].geometryIndex = uint32_t(mGeometryInstances[i]->geometryIndex);
  }
}
}  // namespace kuafu
