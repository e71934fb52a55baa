// Facade API behaviour against the reference's public headers (include/kuafu.hpp, include/core/
// scene.hpp, config.hpp): ownership, limits and error messages, exercised host-only (no device).
// Built and run by tests/test_cpu_facade_api.py.
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>

#include "kuafu.hpp"

using namespace kuafu;

static int failures = 0;
#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      failures++;                                                       \
    }                                                                   \
  } while (0)

static bool throwsWith(const std::function<void()>& f, const char* needle) {
  try {
    f();
  } catch (const std::runtime_error& e) {
    if (std::strstr(e.what(), needle)) return true;
    std::printf("  threw \"%s\", expected \"%s\"\n", e.what(), needle);
    return false;
  }
  return false;
}

int main() {
  auto config = std::make_shared<Config>();
  config->setHostOnly(true);
  config->setGeometryLimit(3);            // reference include/core/config.hpp:157-163
  config->setGeometryInstanceLimit(4);
  Kuafu renderer(config);
  Scene* scene = renderer.getScene();
  CHECK(scene != nullptr);

  // cameras are owned by the scene, handed out as raw pointers (scene.hpp:117-121)
  Camera* cam = scene->createCamera(64, 48);
  scene->setCamera(cam);
  CHECK(scene->getCamera() == cam);
  CHECK(cam->getWidth() == 64 && cam->getHeight() == 48);

  NiceMaterial mat;
  mat.diffuseColor = glm::vec3(0.5f);
  auto cube = createCube(true, mat);
  auto sphere = createSphere(true, mat);
  auto plane = createYZPlane(true, mat);
  scene->submitGeometry(cube);
  scene->submitGeometry(sphere);
  scene->submitGeometry(plane);
  // config.cpp:101-120: a geometry limit below 16 is raised to 16, and one slot is added; so the
  // 18th geometry is the first one refused (scene.cpp:115-119)
  for (int k = 3; k < 17; k++) scene->submitGeometry(createCube(true, mat));
  CHECK(throwsWith([&] { scene->submitGeometry(createCube(true, mat)); }, "geometries buffer size has been exceeded"));
  // scene.cpp:152-158
  CHECK(throwsWith([&] { scene->removeGeometry(cube); }, "Not implemented!"));
  CHECK(throwsWith([&] { scene->removeGeometry(0u); }, "Not implemented!"));

  auto i0 = instance(cube, glm::mat4(1.0f));
  auto i1 = instance(sphere, glm::translate(glm::mat4(1.0f), glm::vec3(2.0f, 0.0f, 0.0f)));
  auto i2 = instance(plane, glm::mat4(1.0f));
  auto i3 = instance(cube, glm::translate(glm::mat4(1.0f), glm::vec3(-2.0f, 0.0f, 0.0f)));
  auto i4 = instance(sphere, glm::translate(glm::mat4(1.0f), glm::vec3(0.0f, 0.0f, 2.0f)));
  for (auto& i : {i0, i1, i2, i3, i4}) scene->submitGeometryInstance(i);
  // scene.cpp:60-64 compares with '>': a limit of 4 admits five instances, the sixth is refused
  CHECK(throwsWith([&] { scene->submitGeometryInstance(instance(cube, glm::mat4(1.0f))); }, "instance buffer size has been exceeded"));
  scene->removeGeometryInstance(i3);
  scene->submitGeometryInstance(instance(sphere, glm::mat4(1.0f)));  // room again

  // instances are shared with the caller, who moves them in place (scene.hpp:156-187, rt.cpp:127)
  i1->setTransform(glm::translate(glm::mat4(1.0f), glm::vec3(0.0f, 3.0f, 0.0f)));
  CHECK(i1->transform[3][1] == 3.0f);

  // a host-only renderer packs scenes but cannot render: no CPU fallback
  CHECK(throwsWith([&] { renderer.run(); }, "host-only"));

  // materials live in the process-wide registry, indexed per primitive (global.hpp:33-35)
  CHECK(!cube->matIndex.empty() && cube->matIndex.front() < global::materials.size());
  NiceMaterial glass = mat;
  glass.alpha = 0.5f;
  cube->setMaterial(glass);
  CHECK(!cube->isOpaque && global::materials[cube->matIndex.front()].alpha == 0.5f);

  // a second scene, then removal of the first (kuafu.hpp:79-82)
  Scene* other = renderer.createScene();
  renderer.setScene(other);
  CHECK(renderer.getScene() == other);
  renderer.removeScene(scene);

  std::printf(failures ? "%d failure(s)\n" : "facade api ok\n", failures);
  return failures ? 1 : 0;
}
