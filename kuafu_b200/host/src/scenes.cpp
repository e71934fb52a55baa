#include "scenes.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>

namespace kuafu::scenes {
namespace {

// Raw mt19937 words -> [0,1): distributions are implementation-defined, the engine is not.
struct Rng {
  std::mt19937 eng;
  explicit Rng(uint32_t seed) : eng(seed) {}
  float uniform() { return float(eng() >> 8) * (1.0f / 16777216.0f); }
  float range(float lo, float hi) { return lo + (hi - lo) * uniform(); }
};

glm::mat4 trs(glm::vec3 t, float angleDeg, glm::vec3 axis, glm::vec3 s) {
  glm::mat4 m = glm::translate(glm::mat4(1.0F), t);
  if (angleDeg != 0.0f) m = glm::rotate(m, glm::radians(angleDeg), axis);
  return glm::scale(m, s);
}

NiceMaterial material(glm::vec3 diffuse, float specular, float metallic, float roughness, float ior = 1.4f,
                      float transmission = 0.0f) {
  NiceMaterial m;
  m.diffuseColor = diffuse;
  m.specular = specular;
  m.metallic = metallic;
  m.roughness = roughness;
  m.ior = ior;
  m.alpha = 1.0F;
  m.transmission = transmission;
  return m;
}

// Wavefront text of a Geometry: one v / vt / vn per vertex, faces as i/i/i; %.9g round-trips a float.
// vt carries 1 - v because loadScene() flips it back (aiProcess_FlipUVs).
void writeObj(const Geometry& g, const std::string& path) {
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("cannot write " + path);
  for (const Vertex& v : g.vertices) {
    std::fprintf(f, "v %.9g %.9g %.9g\n", v.pos.x, v.pos.y, v.pos.z);
    std::fprintf(f, "vt %.9g %.9g\n", v.texCoord.x, 1.0f - v.texCoord.y);
    std::fprintf(f, "vn %.9g %.9g %.9g\n", v.normal.x, v.normal.y, v.normal.z);
  }
  for (size_t i = 0; i + 2 < g.indices.size(); i += 3) {
    const unsigned a = g.indices[i] + 1, b = g.indices[i + 1] + 1, c = g.indices[i + 2] + 1;
    std::fprintf(f, "f %u/%u/%u %u/%u/%u %u/%u/%u\n", a, a, a, b, b, b, c, c, c);
  }
  std::fclose(f);
}

void applyConfig(Kuafu& r, const Recipe& rc, int spp, int depth, bool rr) {
  Config& c = r.getConfig();
  c.setPerPixelSampleRate(uint32_t(rc.spp > 0 ? rc.spp : spp));
  c.setPathDepth(uint32_t(rc.depth >= 0 ? rc.depth : depth));
  c.setRussianRoulette(rr);
  c.setRussianRouletteMinBounces(4);
}

Camera* mainCamera(Kuafu& r, const Recipe& rc, int w, int h) {
  Scene* s = r.getScene();
  Camera* cam = s->createCamera(rc.width > 0 ? rc.width : w, rc.height > 0 ? rc.height : h);
  s->setCamera(cam);
  return cam;
}

// Where the reference's own assets are, for the "_ref" recipes (they cannot ship with this repository).
std::string referenceResource(const std::string& relative) {
  const char* root = std::getenv("KUAFU_REFERENCE");
  const std::string path = std::string(root && *root ? root : "/root/reference") + "/resources/" + relative;
  std::ifstream probe(path, std::ios::binary);
  if (!probe.good()) throw std::runtime_error("reference asset not found: " + path);
  return path;
}

// The objects shared by the eSpheres and eActive levels of the example app.  refAssets: the scanned head
// is the reference's resources/models/suzanne.dae (Example.hpp:292-306) instead of the stand-in.
void tableTop(Scene* scene, bool withGlassCube, float floorMetallic, bool refAssets = false) {
  auto floor = createYZPlane(true, material(glm::vec3(0.8f), 0.5f, floorMetallic, 0.1f));
  std::shared_ptr<Geometry> glassCube;
  if (withGlassCube) glassCube = createCube(true, material({1.0F, 0.7F, 0.7F}, 0.0f, 0.1f, 0.01f, 1.45f, 1.0f));
  std::shared_ptr<Geometry> balls[5];
  const float rough[5] = {0.07f, 0.2f, 0.4f, 0.6f, 0.8f};
  for (int i = 0; i < 5; i++) balls[i] = createSphere(true, material({0.7F, 0.40F, 0.1F}, 0.0f, 1.0f, rough[i]));
  auto mirror = createYZPlane(true, material(glm::vec3(0.9f), 1.0f, 1.0f, 0.0f));
  auto capsule = createCapsule(2, 1, true, material(glm::vec3(1.0f), 0.0f, 0.0f, 0.0f, 1.4f, 1.0f));

  std::vector<std::shared_ptr<Geometry>> geoms = {floor};
  if (glassCube) geoms.push_back(glassCube);
  for (auto& b : balls) geoms.push_back(b);
  geoms.push_back(mirror);
  geoms.push_back(capsule);
  scene->setGeometries(geoms);

  std::vector<std::shared_ptr<GeometryInstance>> insts;
  insts.push_back(instance(floor, trs({0.0F, 0.0F, -1.F}, 90.F, {0., -1., 0.}, glm::vec3(12.0F))));
  if (glassCube) insts.push_back(instance(glassCube, trs({-1.F, -2.0F, 4.F}, 0.f, {0, 0, 1}, glm::vec3(2.0f))));
  const float xs[5] = {-5.0f, -2.5f, 0.0f, 2.5f, 5.0f};
  for (int i = 0; i < 5; i++) insts.push_back(instance(balls[i], glm::translate(glm::mat4(1.0F), {xs[i], 0.0F, 0.0F})));
  insts.push_back(instance(mirror, trs({7.0F, 0.0F, 0.0F}, 180.F, {0., 1., 0.}, glm::vec3(8.F))));
  glm::mat4 cap = glm::translate(glm::mat4(1.0F), {-4.5F, -4.0F, 0.5F});
  cap = glm::rotate(cap, glm::radians(-45.F), {0., 0., 1.});
  cap = glm::rotate(cap, glm::radians(45.F), {1., 1., 0.});
  insts.push_back(instance(capsule, cap));
  scene->setGeometryInstances(insts);

  // the scanned head model of the example (suzanne.dae) -> procedural stand-in of equal size
  const NiceMaterial headMat = material({0.2F, 0.2F, 0.2F}, 0.0f, 1.0f, 0.05f, 1.45f);
  std::shared_ptr<Geometry> head;
  if (refAssets) {
    auto meshes = loadScene(referenceResource("models/suzanne.dae"), true);
    if (meshes.empty()) throw std::runtime_error("suzanne.dae holds no mesh");
    head = meshes.front();
    head->setMaterial(headMat);
  } else {
    head = createBlob(headMat);
  }
  glm::mat4 t = glm::translate(glm::mat4(1.0F), {-2.0F, 5.0F, 2.F});
  t = glm::rotate(t, glm::radians(-90.f), {0, 0, 1});
  t = glm::rotate(t, glm::radians(-20.f), {1, 0, 0});
  t = glm::scale(t, glm::vec3(3.0f));
  scene->submitGeometry(head);
  scene->submitGeometryInstance(instance(head, t));
}

std::vector<Camera*> loadSpheres(Kuafu& r, const Recipe& rc, bool refAssets = false) {  // config 1
  applyConfig(r, rc, 4, 8, false);
  Scene* scene = r.getScene();
  Camera* cam = mainCamera(r, rc, 800, 600);
  scene->setClearColor({0.64F, 0.60F, 0.52F, 0.3F});
  cam->setPosition({-12.6F, 0.0F, 15.4F});
  cam->setFront({0.67F, 0.0F, -0.8F});
  auto sun = std::make_shared<DirectionalLight>();
  sun->direction = {-2, -1, -1};
  sun->color = {1., 0.9, 0.7};
  sun->strength = 8;
  sun->softness = 0.5;
  scene->setDirectionalLight(sun);
  tableTop(scene, true, 0.0f, refAssets);
  scene->removeEnvironmentMap();
  return {cam};
}

// Square dot pattern in the spirit of a RealSense D415 projector: white dots on black.
std::string irPattern() {
  const uint32_t n = 1024;
  std::vector<uint8_t> px(size_t(n) * n * 4, 0);
  for (size_t i = 0; i < size_t(n) * n; i++) px[4 * i + 3] = 255;
  Rng rng(3);
  for (int d = 0; d < 9000; d++) {
    const int cx = int(rng.uniform() * n), cy = int(rng.uniform() * n);
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const int x = cx + dx, y = cy + dy;
        if (x < 0 || y < 0 || x >= int(n) || y >= int(n)) continue;
        const uint8_t v = (dx == 0 && dy == 0) ? 255 : 140;
        uint8_t* p = &px[4 * (size_t(y) * n + x)];
        p[0] = p[1] = p[2] = std::max(p[0], v);
      }
  }
  return global::registerMemoryTexture("ir-dot-pattern", n, n, px.data());
}

std::vector<Camera*> loadActive(Kuafu& r, const Recipe& rc, bool refAssets = false) {  // config 4
  applyConfig(r, rc, 32, 8, false);
  Scene* scene = r.getScene();
  scene->setClearColor({0.64F, 0.60F, 0.52F, 0.0F});
  auto projector = std::make_shared<ActiveLight>();
  projector->fov = glm::radians(150.F);
  projector->viewMat = glm::lookAt(glm::vec3{-3., -3., 8.}, {0., 0., 0.}, {-1., 0.5, 0});
  projector->color = {1., 1., 1.};
  projector->strength = 1000;
  projector->softness = 0;  // the example's softness 1 is flagged as incorrectly implemented
  // Example.hpp:496: resources/patterns/fakesense_j415.png (3000 x 3000, 36 MB as RGBA8); the dot grid of
  // irPattern() stands in for it where the reference tree is not mounted
  projector->texPath = refAssets ? referenceResource("patterns/fakesense_j415.png") : irPattern();
  scene->addActiveLight(projector);
  tableTop(scene, false, 0.9f, refAssets);
  scene->removeEnvironmentMap();
  // stereo IR pair, 55 mm baseline along camera-right
  const int w = rc.width > 0 ? rc.width : 1280, h = rc.height > 0 ? rc.height : 720;
  const float k = float(w) / 1280.0f;
  const glm::vec3 pos(-12.6F, 1.1F, 15.4F), front(0.67F, 0.0F, -0.8F), up(0, 0, 1);
  const glm::vec3 right = glm::normalize(glm::cross(front, up));
  std::vector<Camera*> cams;
  for (int eye = 0; eye < 2; eye++) {
    Camera* c = scene->createCamera(w, h);
    c->setFullPerspective(float(w), float(h), 920.f * k, 920.f * k, 640.f * k, 360.f * k, 0.f);
    c->setPosition(pos + right * (eye ? 0.0275f : -0.0275f));
    c->setFront(front);
    cams.push_back(c);
  }
  scene->setCamera(cams[0]);
  return cams;
}

std::shared_ptr<Geometry> quadGeometry(glm::vec3 a, glm::vec3 b, glm::vec3 c, glm::vec3 d, NiceMaterial mat) {
  auto g = createYZPlane(true, std::move(mat));  // registers the material, then the corners are replaced
  const glm::vec3 n = glm::normalize(glm::cross(b - a, c - a));
  const glm::vec3 p[4] = {a, b, c, d};
  for (int i = 0; i < 4; i++) {
    g->vertices[i].pos = p[i];
    g->vertices[i].normal = n;
  }
  return g;
}

std::vector<Camera*> loadCornell(Kuafu& r, const Recipe& rc) {  // config 2
  applyConfig(r, rc, 64, 12, true);
  Scene* scene = r.getScene();
  Camera* cam = mainCamera(r, rc, 1024, 1024);
  scene->setClearColor({0.0F, 0.0F, 0.0F, 0.2F});
  cam->setPosition({0.0F, -1.98F, 0.0F});
  cam->setFront({0.0F, 1.0F, 0.0F});
  cam->setUp({0.0F, 0.0F, 1.0F});
  const NiceMaterial white = material(glm::vec3(0.73f), 0.0f, 0.0f, 1.0f);
  const NiceMaterial red = material({0.65f, 0.05f, 0.05f}, 0.0f, 0.0f, 1.0f);
  const NiceMaterial green = material({0.12f, 0.45f, 0.15f}, 0.0f, 0.0f, 1.0f);
  NiceMaterial lamp;
  lamp.emission = {1.0f, 0.85f, 0.6f};
  lamp.emissionStrength = 15.0f;
  std::vector<std::shared_ptr<Geometry>> g;
  // room [-1,1]^3 open towards -y, faces wound to look inwards
  g.push_back(quadGeometry({-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, white));   // floor
  g.push_back(quadGeometry({-1, -1, 1}, {-1, 1, 1}, {1, 1, 1}, {1, -1, 1}, white));       // ceiling
  g.push_back(quadGeometry({-1, 1, -1}, {1, 1, -1}, {1, 1, 1}, {-1, 1, 1}, white));       // back
  g.push_back(quadGeometry({-1, -1, -1}, {-1, 1, -1}, {-1, 1, 1}, {-1, -1, 1}, red));     // left
  g.push_back(quadGeometry({1, -1, -1}, {1, -1, 1}, {1, 1, 1}, {1, 1, -1}, green));       // right
  g.push_back(quadGeometry({-0.5f, -0.5f, 0.995f}, {-0.5f, 0.5f, 0.995f}, {0.5f, 0.5f, 0.995f}, {0.5f, -0.5f, 0.995f}, lamp));
  auto tall = createCube(true, white);
  auto shortBox = createCube(true, white);
  auto glass = createSphere(true, material(glm::vec3(1.0f), 0.0f, 0.0f, 0.0f, 1.5f, 1.0f));
  auto metal = createSphere(true, material({0.9f, 0.75f, 0.4f}, 0.5f, 1.0f, 0.3f));
  g.push_back(tall);
  g.push_back(shortBox);
  g.push_back(glass);
  g.push_back(metal);
  scene->setGeometries(g);
  std::vector<std::shared_ptr<GeometryInstance>> insts;
  for (int i = 0; i < 6; i++) insts.push_back(instance(g[i]));
  insts.push_back(instance(tall, trs({-0.35f, 0.35f, -0.4f}, 20.f, {0, 0, 1}, {0.3f, 0.3f, 0.6f})));
  insts.push_back(instance(shortBox, trs({0.4f, -0.1f, -0.7f}, -18.f, {0, 0, 1}, glm::vec3(0.3f))));
  insts.push_back(instance(glass, trs({0.4f, -0.1f, -0.1f}, 0.f, {0, 0, 1}, glm::vec3(0.3f))));
  insts.push_back(instance(metal, trs({-0.45f, -0.45f, -0.75f}, 0.f, {0, 0, 1}, glm::vec3(0.25f))));
  scene->setGeometryInstances(insts);
  scene->removeEnvironmentMap();
  return {cam};
}

// 512 x 512 value-noise texture (bilinear blend of a 16 x 16 lattice), one per material channel.
std::string noiseTexture(Rng& rng, const std::string& name, bool colour) {
  const uint32_t n = 512, cells = 16;
  float lattice[3][17][17];
  for (int ch = 0; ch < 3; ch++)
    for (uint32_t j = 0; j <= cells; j++)
      for (uint32_t i = 0; i <= cells; i++) lattice[ch][j][i] = rng.uniform();
  std::vector<uint8_t> px(size_t(n) * n * 4);
  for (uint32_t y = 0; y < n; y++)
    for (uint32_t x = 0; x < n; x++) {
      const float fx = float(x) * cells / n, fy = float(y) * cells / n;
      const uint32_t ix = uint32_t(fx), iy = uint32_t(fy);
      const float tx = fx - ix, ty = fy - iy;
      uint8_t* p = &px[4 * (size_t(y) * n + x)];
      for (int ch = 0; ch < 3; ch++) {
        const float (*L)[17] = lattice[colour ? ch : 0];
        const float v = (L[iy][ix] * (1 - tx) + L[iy][ix + 1] * tx) * (1 - ty) + (L[iy + 1][ix] * (1 - tx) + L[iy + 1][ix + 1] * tx) * ty;
        p[ch] = uint8_t(std::min(255.0f, 40.0f + v * 215.0f));
      }
      p[3] = 255;
    }
  return global::registerMemoryTexture(name, n, n, px.data());
}

void skyCube(Scene* scene) {
  const uint32_t n = 256;
  std::vector<uint8_t> faces[6];
  const uint8_t* ptr[6];
  for (int f = 0; f < 6; f++) {
    faces[f].resize(size_t(n) * n * 4);
    for (uint32_t y = 0; y < n; y++)
      for (uint32_t x = 0; x < n; x++) {
        // the miss shader looks the cube up with (-d.y, d.z, -d.x): cube +Y is world up (+z)
        float g;
        if (f == 2) g = 1.0f; else if (f == 3) g = 0.0f; else g = 1.0f - float(y) / float(n - 1);
        uint8_t* p = &faces[f][4 * (size_t(y) * n + x)];
        p[0] = uint8_t(90 + 120 * g);
        p[1] = uint8_t(110 + 120 * g);
        p[2] = uint8_t(140 + 110 * g);
        p[3] = 255;
      }
    ptr[f] = faces[f].data();
  }
  scene->setEnvironmentMapFaces(ptr, n);
}

// "via the Assimp path" (BASELINE config 3): the meshes are written out as Wavefront files and come
// back through loadScene() -- the importer the reference reaches through Assimp -- before the
// PrincipledBSDF materials are assigned with Geometry::setMaterial().  Positions, normals and face
// order survive the text round trip bit for bit (9 significant digits); v passes through FlipUVs twice
// (1 - (1 - v)), which may move it by one ulp.
std::shared_ptr<Geometry> throughObjFile(const std::shared_ptr<Geometry>& g, const std::string& tag) {
  const char* tmp = std::getenv("TMPDIR");
  const std::string path = std::string(tmp && *tmp ? tmp : "/tmp") + "/kuafu_b200_" + tag + ".obj";
  writeObj(*g, path);
  const uint32_t matIndex = g->matIndex.empty() ? 0u : g->matIndex.front();
  auto loaded = loadObj(path, g->dynamic);
  std::remove(path.c_str());
  loaded->setMaterial(global::materials[matIndex]);
  return loaded;
}

std::vector<Camera*> loadMillion(Kuafu& r, const Recipe& rc, bool viaObj = false) {  // config 3
  applyConfig(r, rc, 64, 8, true);
  Scene* scene = r.getScene();
  Camera* cam = mainCamera(r, rc, 1920, 1080);
  cam->setPosition({-16.0F, -9.0F, 9.5F});
  cam->setFront({0.78F, 0.44F, -0.44F});
  scene->setClearColor({0.6F, 0.7F, 0.9F, 1.0F});
  auto sun = std::make_shared<DirectionalLight>();
  sun->direction = {-1.0f, 0.6f, -1.5f};
  sun->color = {1.0f, 0.95f, 0.85f};
  sun->strength = 4;
  sun->softness = 0.2f;
  scene->setDirectionalLight(sun);

  Rng texRng(2);
  std::vector<std::shared_ptr<Geometry>> geoms;
  NiceMaterial floorMat = material(glm::vec3(0.7f), 0.5f, 0.0f, 0.4f);
  floorMat.diffuseTexPath = noiseTexture(texRng, "million-floor", true);
  auto floor = createYZPlane(true, floorMat);
  if (viaObj) floor = throughObjFile(floor, "million_floor");
  geoms.push_back(floor);
  const int nMaterials = 16;
  for (int m = 0; m < nMaterials; m++) {
    NiceMaterial mat = material(glm::vec3(0.8f), 0.5f, (m % 4 == 1) ? 1.0f : 0.0f, 0.3f, 1.45f, (m % 4 == 2) ? 1.0f : 0.0f);
    mat.diffuseTexPath = noiseTexture(texRng, "million-d" + std::to_string(m), true);
    mat.roughnessTexPath = noiseTexture(texRng, "million-r" + std::to_string(m), false);
    if (m % 4 == 3) mat.transmissionTexPath = noiseTexture(texRng, "million-t" + std::to_string(m), false);
    auto sphere = createSphere(true, mat);
    geoms.push_back(viaObj ? throughObjFile(sphere, "million_sphere" + std::to_string(m)) : sphere);
  }
  scene->setGeometries(geoms);

  const int nSpheres = rc.scale > 0 ? rc.scale : 204;
  Rng rng(1);
  std::vector<std::shared_ptr<GeometryInstance>> insts;
  insts.push_back(instance(floor, trs({0.0F, 0.0F, -1.F}, 90.F, {0., -1., 0.}, glm::vec3(30.0F))));
  // jittered 3-D grid: 9 x 8 x 3 cells of 3 units
  int placed = 0;
  for (int k = 0; k < 3 && placed < nSpheres; k++)
    for (int j = 0; j < 8 && placed < nSpheres; j++)
      for (int i = 0; i < 9 && placed < nSpheres; i++) {
        if (placed >= nSpheres) break;
        const glm::vec3 c((i - 4) * 3.0f + rng.range(-0.6f, 0.6f), (j - 3.5f) * 3.0f + rng.range(-0.6f, 0.6f),
                          0.2f + k * 2.8f + rng.range(-0.4f, 0.4f));
        const float s = rng.range(0.7f, 1.25f);
        const glm::vec3 axis(rng.range(-1.f, 1.f), rng.range(-1.f, 1.f), rng.range(0.1f, 1.f));
        insts.push_back(instance(geoms[1 + (placed % nMaterials)], trs(c, rng.range(0.f, 360.f), axis, {s, s * rng.range(0.8f, 1.2f), s})));
        placed++;
      }
  scene->setGeometryInstances(insts);
  skyCube(scene);
  return {cam};
}

// ---- config 5: articulated chains ------------------------------------------------------------
struct Articulated {
  int chains = 64, links = 32;
  std::vector<std::shared_ptr<GeometryInstance>> linkInstances;
};
Articulated g_articulated;

glm::mat4 linkTransform(int chain, int link, int frame, glm::vec3& cursor, float& heading, float& pitch) {
  // deterministic joint angles: each joint swings with its own phase
  const float t = 0.05f * float(frame);
  const float dYaw = 0.35f * std::sin(t * 1.3f + 0.7f * link + 0.37f * chain);
  const float dPitch = 0.25f * std::sin(t * 0.9f + 0.5f * link + 1.1f * chain);
  heading += dYaw;
  pitch = 0.6f * pitch + dPitch;
  const glm::vec3 dir(std::cos(heading) * std::cos(pitch), std::sin(heading) * std::cos(pitch), std::sin(pitch));
  const float len = 1.0f;
  const glm::vec3 centre = cursor + dir * (0.5f * len);
  cursor = cursor + dir * len;
  // orient the link's local x axis along dir
  const glm::vec3 up(0, 0, 1);
  glm::vec3 side = glm::cross(up, dir);
  if (glm::dot(side, side) < 1e-8f) side = glm::vec3(0, 1, 0);
  side = glm::normalize(side);
  const glm::vec3 top = glm::cross(dir, side);
  glm::mat4 m(glm::vec4(dir * (0.5f * len), 0.f), glm::vec4(side * 0.28f, 0.f), glm::vec4(top * 0.28f, 0.f), glm::vec4(centre, 1.f));
  return m;
}

void poseChains(int frame) {
  Articulated& a = g_articulated;
  for (int c = 0; c < a.chains; c++) {
    const int gx = c % 8, gy = c / 8;
    glm::vec3 cursor((gx - 3.5f) * 4.5f, (gy - 3.5f) * 4.5f, 0.4f);
    float heading = 0.4f * c, pitch = 0.5f;
    for (int l = 0; l < a.links; l++) {
      const size_t idx = size_t(c) * a.links + l;
      if (idx >= a.linkInstances.size()) return;
      a.linkInstances[idx]->setTransform(linkTransform(c, l, frame, cursor, heading, pitch));
    }
  }
}

std::vector<Camera*> loadArticulated(Kuafu& r, const Recipe& rc) {  // config 5
  applyConfig(r, rc, 32, 8, true);
  r.getConfig().setGeometryInstanceLimit(4096);
  r.getContext().init();
  Scene* scene = r.getScene();
  scene->setClearColor({0.55F, 0.6F, 0.7F, 1.0F});
  auto sun = std::make_shared<DirectionalLight>();
  sun->direction = {-1.0f, -0.4f, -1.2f};
  sun->color = {1.0f, 0.97f, 0.9f};
  sun->strength = 5;
  sun->softness = 0.1f;
  scene->setDirectionalLight(sun);
  auto floor = createYZPlane(true, material(glm::vec3(0.6f), 0.5f, 0.0f, 0.5f));
  std::shared_ptr<Geometry> link[4];
  const glm::vec3 tint[4] = {{0.8f, 0.3f, 0.2f}, {0.2f, 0.5f, 0.8f}, {0.85f, 0.8f, 0.3f}, {0.7f, 0.7f, 0.7f}};
  for (int i = 0; i < 4; i++) link[i] = createSphere(true, material(tint[i], 0.5f, i == 3 ? 1.0f : 0.0f, 0.35f));
  scene->setGeometries({floor, link[0], link[1], link[2], link[3]});
  Articulated& a = g_articulated;
  a.chains = rc.scale > 0 ? std::min(64, rc.scale) : 64;
  a.links = 32;
  a.linkInstances.clear();
  std::vector<std::shared_ptr<GeometryInstance>> insts;
  insts.push_back(instance(floor, trs({0.0F, 0.0F, -0.2F}, 90.F, {0., -1., 0.}, glm::vec3(40.0F))));
  for (int c = 0; c < a.chains; c++)
    for (int l = 0; l < a.links; l++) {
      auto inst = instance(link[(c + l) % 4]);
      a.linkInstances.push_back(inst);
      insts.push_back(inst);
    }
  poseChains(0);
  scene->setGeometryInstances(insts);
  scene->removeEnvironmentMap();
  const int nCams = rc.scale > 0 ? std::min(64, std::max(1, rc.scale)) : 64;
  const int w = rc.width > 0 ? rc.width : 512, h = rc.height > 0 ? rc.height : 512;
  std::vector<Camera*> cams;
  for (int i = 0; i < nCams; i++) {
    const float ang = 6.2831853f * float(i) / float(nCams);
    const glm::vec3 pos(21.0f * std::cos(ang), 21.0f * std::sin(ang), 8.0f + 3.0f * std::sin(3.0f * ang));
    Camera* c = scene->createCamera(w, h);
    c->setPosition(pos);
    c->setFront(glm::normalize(glm::vec3(0, 0, 1.5f) - pos));
    cams.push_back(c);
  }
  scene->setCamera(cams[0]);
  return cams;
}
}  // namespace

std::shared_ptr<Geometry> createBlob(NiceMaterial mat, uint32_t slices, uint32_t stacks, uint32_t seed) {
  auto g = createSphere(true, std::move(mat));  // registers the material; tessellation is replaced below
  g->vertices.clear();
  g->indices.clear();
  const float pi = glm::pi<float>();
  // seed 0 is the suzanne stand-in; any other seed shifts the phases and amplitudes of the displacement
  // (theta terms keep integer frequencies so the surface closes on itself), so that no two blobs share a shape
  Rng srng(seed * 2654435761u + 1u);
  const float p0 = seed ? srng.range(0.f, 6.28f) : 0.0f, p1 = seed ? srng.range(0.f, 6.28f) : 0.0f;
  const float a0 = seed ? srng.range(0.08f, 0.22f) : 0.18f, a1 = seed ? srng.range(0.03f, 0.09f) : 0.07f;
  auto radiusAt = [=](float phi, float theta) {
    return 1.0f + a0 * std::sin(3.0f * theta + p0) * std::cos(2.0f * phi) + a1 * std::sin(9.0f * phi + 2.0f * theta + p1) +
           0.12f * std::cos(2.0f * theta + 1.0f) * std::cos(phi);
  };
  for (uint32_t i = 1; i < stacks; ++i) {
    const float phi = pi / stacks * i - pi / 2;
    for (uint32_t j = 0; j < slices; ++j) {
      const float theta = pi * 2 / slices * j;
      const float rr = radiusAt(phi, theta);
      Vertex v;
      v.pos = {std::sin(phi) * rr, std::cos(theta) * std::cos(phi) * rr, std::sin(theta) * std::cos(phi) * rr};
      v.texCoord = {float(j) / slices, float(i) / stacks};
      g->vertices.push_back(v);
    }
  }
  auto next = [&](uint32_t i) { return (i + 1) % slices + i / slices * slices; };
  const uint32_t body = (stacks - 2) * slices;
  for (uint32_t i = 0; i < body; ++i) {
    const uint32_t right = next(i), up = i + slices, rightUp = right + slices;
    for (uint32_t k : {i, rightUp, up, i, right, rightUp}) g->indices.push_back(k);
  }
  Vertex south, north;
  south.pos = {-radiusAt(-pi / 2, 0), 0, 0};
  north.pos = {radiusAt(pi / 2, 0), 0, 0};
  g->vertices.push_back(south);
  g->vertices.push_back(north);
  const uint32_t si = uint32_t(g->vertices.size()) - 2, ni = si + 1;
  for (uint32_t i = 0; i < slices; ++i)
    for (uint32_t k : {si, next(i), i}) g->indices.push_back(k);
  for (uint32_t i = body; i < body + slices; ++i)
    for (uint32_t k : {ni, i, next(i)}) g->indices.push_back(k);
  g->matIndex.assign(g->indices.size(), g->matIndex.empty() ? 0u : g->matIndex.front());
  g->path = "procedural:suzanne-substitute";
  g->recalculateNormals();
  return g;
}

// Stress recipes with UNIQUE geometry (no instancing): the layout of config 3, but every object is its
// own displaced blob with its own bottom-level structure, so the traversal's working set is the whole
// triangle count instead of 17 shared meshes -- what a scene of scanned assets looks like, and the case
// where node and triangle fetches actually leave the 126 MB L2.
//   "unique"     204 blobs x 4 900 triangles  =  1.0 M unique triangles (~ 0.15 GB of nodes, triangles, shading records)
//   "unique10m"  1 000 blobs x 10 000 triangles = 10.0 M unique triangles (~ 1.5 GB)
std::vector<Camera*> loadUnique(Kuafu& r, const Recipe& rc, int nDefault, uint32_t slices, uint32_t stacks) {
  applyConfig(r, rc, 16, 8, true);
  Scene* scene = r.getScene();
  const int n = rc.scale > 0 ? rc.scale : nDefault;
  int f = 1;
  while (216 * f * f * f < n) f++;
  Camera* cam = mainCamera(r, rc, 1920, 1080);
  cam->setPosition(glm::vec3(-16.0F, -9.0F, 9.5F) * float(f));
  cam->setFront({0.78F, 0.44F, -0.44F});
  scene->setClearColor({0.6F, 0.7F, 0.9F, 1.0F});
  auto sun = std::make_shared<DirectionalLight>();
  sun->direction = {-1.0f, 0.6f, -1.5f};
  sun->color = {1.0f, 0.95f, 0.85f};
  sun->strength = 4;
  sun->softness = 0.2f;
  scene->setDirectionalLight(sun);
  Rng texRng(2);
  std::vector<std::shared_ptr<Geometry>> geoms;
  NiceMaterial floorMat = material(glm::vec3(0.7f), 0.5f, 0.0f, 0.4f);
  floorMat.diffuseTexPath = noiseTexture(texRng, "unique-floor", true);
  auto floor = createYZPlane(true, floorMat);
  geoms.push_back(floor);
  // sixteen textured materials shared by all blobs (the reference loads the textures of every material
  // entry, scene.cpp:314-413, so one entry per blob would exceed the texture limit)
  const int nMaterials = 16;
  std::vector<uint32_t> mats;
  for (int m = 0; m < nMaterials; m++) {
    NiceMaterial mat = material(glm::vec3(0.8f), 0.5f, (m % 4 == 1) ? 1.0f : 0.0f, 0.3f, 1.45f, (m % 4 == 2) ? 1.0f : 0.0f);
    mat.diffuseTexPath = noiseTexture(texRng, "unique-d" + std::to_string(m), true);
    mat.roughnessTexPath = noiseTexture(texRng, "unique-r" + std::to_string(m), false);
    global::materials.push_back(mat);
    mats.push_back(global::materialIndex++);
  }
  const NiceMaterial plain = material(glm::vec3(0.8f), 0.5f, 0.0f, 0.3f);
  Rng rng(1);
  std::vector<std::shared_ptr<GeometryInstance>> insts;
  insts.push_back(instance(floor, trs({0.0F, 0.0F, -1.F}, 90.F, {0., -1., 0.}, glm::vec3(30.0F * float(f)))));
  int placed = 0;
  for (int k = 0; k < 3 * f && placed < n; k++)
    for (int j = 0; j < 8 * f && placed < n; j++)
      for (int i = 0; i < 9 * f && placed < n; i++) {
        auto blob = createBlob(plain, slices, stacks, uint32_t(placed) + 1u);
        blob->matIndex.assign(blob->matIndex.size(), mats[size_t(placed % nMaterials)]);
        geoms.push_back(blob);
        const glm::vec3 c((i - 4.5f * f + 0.5f) * 3.0f + rng.range(-0.6f, 0.6f), (j - 4.0f * f + 0.5f) * 3.0f + rng.range(-0.6f, 0.6f),
                          0.2f + k * 2.8f + rng.range(-0.4f, 0.4f));
        const float s = rng.range(0.7f, 1.25f);
        const glm::vec3 axis(rng.range(-1.f, 1.f), rng.range(-1.f, 1.f), rng.range(0.1f, 1.f));
        insts.push_back(instance(blob, trs(c, rng.range(0.f, 360.f), axis, {s, s * rng.range(0.8f, 1.2f), s})));
        placed++;
      }
  scene->setGeometries(geoms);
  scene->setGeometryInstances(insts);
  skyCube(scene);
  return {cam};
}

// "file:<path>": every mesh of an asset file (loadScene: .obj / .dae / .stl), one instance each, a
// camera that frames the bounding box, a directional light and a light backdrop.
std::vector<Camera*> loadFile(Kuafu& r, const Recipe& rc, const std::string& path) {
  applyConfig(r, rc, 16, 8, false);
  Scene* scene = r.getScene();
  auto geoms = loadScene(path, false);
  glm::vec3 lo(3.0e38f), hi(-3.0e38f);
  for (auto& g : geoms)
    for (auto& v : g->vertices)
      for (int k = 0; k < 3; k++) {
        lo[k] = std::min(lo[k], v.pos[k]);
        hi[k] = std::max(hi[k], v.pos[k]);
      }
  const glm::vec3 centre = (lo + hi) * 0.5f;
  const float radius = std::max(0.5f * glm::length(hi - lo), 1e-3f);
  Camera* cam = mainCamera(r, rc, 640, 480);
  const glm::vec3 eye = centre + glm::vec3(-1.2f, -1.0f, 0.7f) * radius;
  cam->setPosition(eye);
  cam->setFront(glm::normalize(centre - eye));
  scene->setClearColor({0.7F, 0.75F, 0.8F, 1.0F});
  auto sun = std::make_shared<DirectionalLight>();
  sun->direction = {1.0f, 0.8f, -1.2f};
  sun->color = {1.0f, 1.0f, 1.0f};
  sun->strength = 4;
  sun->softness = 0.1f;
  scene->setDirectionalLight(sun);
  scene->setGeometries(geoms);
  std::vector<std::shared_ptr<GeometryInstance>> insts;
  for (auto& g : geoms) insts.push_back(instance(g, glm::mat4(1.0F)));
  scene->setGeometryInstances(insts);
  scene->removeEnvironmentMap();
  return {cam};
}

std::vector<Camera*> load(Kuafu& renderer, const Recipe& rc) {
  if (rc.name.rfind("file:", 0) == 0) return loadFile(renderer, rc, rc.name.substr(5));
  if (rc.name == "million_obj") return loadMillion(renderer, rc, true);
  if (rc.name == "spheres") return loadSpheres(renderer, rc);
  if (rc.name == "cornell") return loadCornell(renderer, rc);
  if (rc.name == "million") return loadMillion(renderer, rc);
  if (rc.name == "active") return loadActive(renderer, rc);
  if (rc.name == "spheres_ref") return loadSpheres(renderer, rc, true);
  if (rc.name == "active_ref") return loadActive(renderer, rc, true);
  if (rc.name == "articulated") return loadArticulated(renderer, rc);
  if (rc.name == "unique") return loadUnique(renderer, rc, 204, 50, 50);
  if (rc.name == "unique10m") return loadUnique(renderer, rc, 1000, 100, 51);
  throw std::runtime_error("unknown scene recipe: " + rc.name);
}

void animate(Kuafu&, int frame) { poseChains(frame); }
}  // namespace kuafu::scenes
