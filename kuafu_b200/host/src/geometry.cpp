// Geometry model: material registry hooks, normal recalculation, procedural meshes and the minimal
// asset importer (reference src/core/geometry.cpp).  The procedural creators reproduce the
// reference's tessellations (createSphere: 50 x 50 -> 4900 triangles, createCapsule: 32 segments x
// 16 rings) because they are the fixtures of BASELINE configs 1, 3, 4 and 5.
#include "core/geometry.hpp"

#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>

#include "core/context/global.hpp"

namespace kuafu {

bool operator==(const NiceMaterial& a, const NiceMaterial& b) {
  return a.diffuseColor == b.diffuseColor && a.alpha == b.alpha && a.diffuseTexPath == b.diffuseTexPath &&
         a.metallicTexPath == b.metallicTexPath && a.roughnessTexPath == b.roughnessTexPath &&
         a.transmissionTexPath == b.transmissionTexPath && a.metallic == b.metallic &&
         a.specular == b.specular && a.roughness == b.roughness && a.ior == b.ior &&
         a.transmission == b.transmission && a.emission == b.emission &&
         a.emissionStrength == b.emissionStrength;
}

// Appends `mat` to the process-wide registry and returns its index (reference geometry.cpp:314-315).
static uint32_t registerMaterial(NiceMaterial mat) {
  global::materials.push_back(std::move(mat));
  return global::materialIndex++;
}

static std::shared_ptr<Geometry> finishProcedural(std::shared_ptr<Geometry> g, bool dynamic, NiceMaterial mat) {
  g->path = "";
  g->initialized = false;
  g->dynamic = dynamic;
  g->isOpaque = mat.alpha >= 1.0F;
  g->matIndex.assign(g->indices.size(), registerMaterial(std::move(mat)));
  return g;
}

void Geometry::setMaterial(const NiceMaterial& material) {
  isOpaque = material.alpha >= 1;
  const uint32_t index = registerMaterial(material);
  std::fill(matIndex.begin(), matIndex.end(), index);
}

std::shared_ptr<GeometryInstance> instance(const std::shared_ptr<Geometry>& geometry, const glm::mat4& transform) {
  KF_ASSERT(geometry != nullptr, "instance(): geometry must not be null");
  auto result = std::make_shared<GeometryInstance>();
  result->geometry = geometry;
  result->transform = transform;
  return result;
}

void GeometryInstance::setTransform(const glm::mat4& t) { transform = t; }

// Unweighted average of the unit face normals around each vertex (reference geometry.cpp:268-295).
void Geometry::recalculateNormals() {
  for (auto& v : vertices) v.normal = glm::vec3(0);
  for (size_t f = 0; f + 2 < indices.size(); f += 3) {
    Vertex& a = vertices[indices[f]];
    Vertex& b = vertices[indices[f + 1]];
    Vertex& c = vertices[indices[f + 2]];
    const glm::vec3 n = glm::normalize(glm::cross(b.pos - a.pos, c.pos - a.pos));
    if (std::isnan(n.x)) continue;
    a.normal += n;
    b.normal += n;
    c.normal += n;
  }
  for (auto& v : vertices)
    if (!(v.normal == glm::vec3(0))) v.normal = glm::normalize(v.normal);
}

static Vertex makeVertex(glm::vec3 p, glm::vec3 n, glm::vec2 uv) {
  Vertex v;
  v.pos = p;
  v.normal = n;
  v.texCoord = uv;
  return v;
}

std::shared_ptr<Geometry> createYZPlane(bool dynamic, NiceMaterial mat) {
  auto g = std::make_shared<Geometry>();
  const glm::vec3 n(1, 0, 0);
  g->vertices = {makeVertex({0, 1, 1}, n, {1, 0}), makeVertex({0, -1, 1}, n, {0, 0}),
                 makeVertex({0, -1, -1}, n, {1, 1}), makeVertex({0, 1, -1}, n, {0, 1})};
  g->indices = {0, 1, 2, 0, 2, 3};
  return finishProcedural(g, dynamic, std::move(mat));
}

// Axis-aligned cube [-1,1]^3, 24 vertices, outward counter-clockwise faces.  (The reference ships a
// Blender-exported vertex table; this one is generated, so vertex order and UVs differ.)
std::shared_ptr<Geometry> createCube(bool dynamic, NiceMaterial mat) {
  auto g = std::make_shared<Geometry>();
  for (int axis = 0; axis < 3; ++axis) {
    for (int side = 0; side < 2; ++side) {
      glm::vec3 n(0.0f), a(0.0f);
      n[axis] = side ? 1.0f : -1.0f;
      a[(axis + 1) % 3] = 1.0f;
      const glm::vec3 b = glm::cross(n, a);
      const uint32_t base = uint32_t(g->vertices.size());
      const float st[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
      for (auto& c : st)
        g->vertices.push_back(makeVertex(n + a * c[0] + b * c[1], n, {(c[0] + 1) * 0.5f, (c[1] + 1) * 0.5f}));
      for (uint32_t k : {0u, 1u, 2u, 0u, 2u, 3u}) g->indices.push_back(base + k);
    }
  }
  return finishProcedural(g, dynamic, std::move(mat));
}

// UV sphere of the reference (geometry.cpp:375-433): rings of `slices` vertices for stacks 1..49,
// two pole vertices appended last, smooth normals from recalculateNormals().
std::shared_ptr<Geometry> createSphere(bool dynamic, NiceMaterial mat) {
  auto g = std::make_shared<Geometry>();
  const uint32_t stacks = 50, slices = 50;
  const float radius = 1.f, pi = glm::pi<float>();
  for (uint32_t i = 1; i < stacks; ++i) {
    const float phi = pi / stacks * i - pi / 2;
    for (uint32_t j = 0; j < slices; ++j) {
      const float theta = pi * 2 / slices * j;
      Vertex v;
      v.pos = {sinf(phi) * radius, cosf(theta) * cosf(phi) * radius, sinf(theta) * cosf(phi) * radius};
      g->vertices.push_back(v);
    }
  }
  auto nextInRing = [&](uint32_t i) { return (i + 1) % slices + i / slices * slices; };
  const uint32_t body = (stacks - 2) * slices;
  for (uint32_t i = 0; i < body; ++i) {
    const uint32_t right = nextInRing(i), up = i + slices, rightUp = right + slices;
    for (uint32_t k : {i, rightUp, up, i, right, rightUp}) g->indices.push_back(k);
  }
  Vertex south, north;
  south.pos = {-radius, 0, 0};
  north.pos = {radius, 0, 0};
  g->vertices.push_back(south);
  g->vertices.push_back(north);
  const uint32_t southIdx = uint32_t(g->vertices.size()) - 2, northIdx = southIdx + 1;
  for (uint32_t i = 0; i < slices; ++i)
    for (uint32_t k : {southIdx, nextInRing(i), i}) g->indices.push_back(k);
  for (uint32_t i = body; i < body + slices; ++i)
    for (uint32_t k : {northIdx, i, nextInRing(i)}) g->indices.push_back(k);
  auto out = finishProcedural(g, dynamic, std::move(mat));
  out->recalculateNormals();
  return out;
}

// Capsule along X (reference geometry.cpp:437-537): 32 segments, 8 + 8 rings, one cap fan per end.
std::shared_ptr<Geometry> createCapsule(float halfLength, float radius, bool dynamic, NiceMaterial mat) {
  auto g = std::make_shared<Geometry>();
  const int segments = 32, halfRings = 8, rings = 2 * halfRings;
  const float pi = glm::pi<float>();
  for (int s = 0; s < segments; ++s)
    g->vertices.push_back(makeVertex({radius + halfLength, 0.f, 0.f}, {1.f, 0.f, 0.f}, {(0.5f + s) / segments, 1.f}));
  auto ring = [&](int r, float shift, float vCoord) {
    const float theta = pi * r / rings;
    const float x = std::cos(theta), yz = std::sin(theta);
    for (int s = 0; s <= segments; ++s) {
      const float phi = pi * s * 2 / segments;
      const glm::vec3 n(x, yz * std::cos(phi), yz * std::sin(phi));
      g->vertices.push_back(makeVertex(n * radius + glm::vec3(shift, 0, 0), n, {float(s) / segments, vCoord}));
    }
  };
  for (int r = 1; r <= halfRings; ++r) ring(r, halfLength, 1.f - 0.5f * float(r) / rings);
  for (int r = halfRings; r < rings; ++r) ring(r, -halfLength, 0.5f - 0.5f * float(r) / rings);
  for (int s = 0; s < segments; ++s)
    g->vertices.push_back(makeVertex({-radius - halfLength, 0.f, 0.f}, {-1.f, 0.f, 0.f}, {(0.5f + s) / segments, 0.f}));

  auto tri = [&](int a, int b, int c) {
    g->indices.push_back(uint32_t(a));
    g->indices.push_back(uint32_t(b));
    g->indices.push_back(uint32_t(c));
  };
  const int stride = segments + 1;
  for (int s = 0; s < segments; ++s) tri(s, s + segments, s + segments + 1);
  for (int r = 0; r < rings - 1; ++r)
    for (int s = 0; s < segments; ++s) {
      const int a = segments + stride * r + s, b = segments + stride * (r + 1) + s;
      tri(a, b, b + 1);
      tri(a, b + 1, a + 1);
    }
  for (int s = 0; s < segments; ++s) {
    const int a = segments + stride * (rings - 1) + s;
    tri(a, segments + stride * rings + s, a + 1);
  }
  return finishProcedural(g, dynamic, std::move(mat));
}

// ---------------------------------------------------------------------------------------------
// Minimal importer: Wavefront OBJ (+ MTL).  Stands in for the reference's Assimp path
// (geometry.cpp:28-232) with the same material rules: shininess -> roughness, OBJ default
// roughness 1, opacity from `d` / `Tr`, UV flip, material de-duplication against the registry,
// one Geometry per material group.
// ---------------------------------------------------------------------------------------------
namespace {
struct ObjMaterial {
  NiceMaterial mat;
  float shininess = -1.f;
  bool hasRoughness = false;
};

std::string dirOf(const std::string& p) {
  const size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}

std::map<std::string, ObjMaterial> readMtl(const std::string& path) {
  std::map<std::string, ObjMaterial> out;
  std::ifstream in(path);
  std::string line, cur;
  const std::string base = dirOf(path);
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string key;
    if (!(ss >> key) || key[0] == '#') continue;
    if (key == "newmtl") {
      ss >> cur;
      out[cur] = ObjMaterial();
      out[cur].mat.diffuseColor = glm::vec3(0.0f);
      out[cur].mat.emission = glm::vec3(0.0f);
      out[cur].mat.emissionStrength = 1.f;
      out[cur].mat.roughness = 0.f;
      continue;
    }
    if (cur.empty()) continue;
    ObjMaterial& m = out[cur];
    auto readTex = [&](std::string& dst) {
      std::string t;
      ss >> t;
      if (!t.empty()) dst = (t[0] == '/' || t.rfind("mem:", 0) == 0) ? t : base + t;
    };
    if (key == "Kd") ss >> m.mat.diffuseColor.x >> m.mat.diffuseColor.y >> m.mat.diffuseColor.z;
    else if (key == "Ke") ss >> m.mat.emission.x >> m.mat.emission.y >> m.mat.emission.z;
    else if (key == "Ns") ss >> m.shininess;
    else if (key == "Ni") ss >> m.mat.ior;
    else if (key == "d") ss >> m.mat.alpha;
    else if (key == "Tr") { float t = 0; ss >> t; m.mat.alpha = 1.f - t; }
    else if (key == "Pr") { ss >> m.mat.roughness; m.hasRoughness = true; }
    else if (key == "Pm") ss >> m.mat.metallic;
    else if (key == "Tf") { float t = 0; ss >> t; m.mat.transmission = 1.f - t; }
    else if (key == "map_Kd") readTex(m.mat.diffuseTexPath);
    else if (key == "map_Pr") readTex(m.mat.roughnessTexPath);
    else if (key == "map_Pm") readTex(m.mat.metallicTexPath);
  }
  for (auto& kv : out) {
    ObjMaterial& m = kv.second;
    if (m.mat.roughness == 0 && !m.hasRoughness) {
      if (m.shininess > 0) {
        m.mat.roughness = m.shininess <= 5.f ? 1.f : 1.f - std::sqrt(m.shininess - 5.f) * 0.025f;
        if (m.mat.roughness < 0.001f) m.mat.roughness = 0.001f;
      } else {
        m.mat.roughness = 1.f;
      }
    }
  }
  return out;
}

uint32_t globalMaterialIndex(const NiceMaterial& m) {
  for (size_t j = 0; j < global::materials.size(); j++)
    if (m == global::materials[j]) return uint32_t(j);
  return registerMaterial(m);
}
bool hasExtension(const std::string& path, const char* ext) {
  const size_t n = std::strlen(ext);
  if (path.size() < n) return false;
  for (size_t i = 0; i < n; i++)
    if (std::tolower(static_cast<unsigned char>(path[path.size() - n + i])) != ext[i]) return false;
  return true;
}
}  // namespace

// import_dae_stl.cpp
uint32_t importMaterialIndex(const NiceMaterial& m) { return globalMaterialIndex(m); }
std::vector<std::shared_ptr<Geometry>> loadColladaScene(const std::string& path, bool dynamic);
std::vector<std::shared_ptr<Geometry>> loadStlScene(const std::string& path, bool dynamic);
std::vector<std::shared_ptr<Geometry>> loadGltfScene(const std::string& path, bool dynamic);  // import_gltf.cpp

// The reference hands every format to Assimp; this build reads Wavefront .obj (below), COLLADA .dae
// and .stl (import_dae_stl.cpp), glTF 2.0 .gltf / .glb (import_gltf.cpp) and says so for anything
// else (FBX / PLY / 3DS ...).
std::vector<std::shared_ptr<Geometry>> loadScene(std::string_view fname, bool dynamic) {
  std::string path(fname);
  if (!path.empty() && path[0] != '/' && !global::assetsPath.empty()) {
    std::ifstream probe(path);
    if (!probe.good()) path = global::assetsPath + path;
  }
  if (hasExtension(path, ".dae")) return loadColladaScene(path, dynamic);
  if (hasExtension(path, ".stl")) return loadStlScene(path, dynamic);
  if (hasExtension(path, ".gltf") || hasExtension(path, ".glb")) return loadGltfScene(path, dynamic);
  if (!hasExtension(path, ".obj"))
    throw std::runtime_error("Failed to load scene: this build imports .obj, .dae, .stl, .gltf and .glb only, " + path);
  std::ifstream in(path);
  if (!in.good()) throw std::runtime_error("Failed to load scene: cannot open " + path);

  std::vector<glm::vec3> pos, nrm;
  std::vector<glm::vec2> uvs;
  std::map<std::string, ObjMaterial> mtl;
  struct Group {
    std::string material;
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;
    std::map<std::array<int, 3>, uint32_t> lookup;
    bool anyMissingNormal = false;
  };
  std::vector<Group> groups;
  auto groupFor = [&](const std::string& name) -> Group& {
    for (auto& g : groups)
      if (g.material == name) return g;
    groups.push_back(Group());
    groups.back().material = name;
    return groups.back();
  };
  std::string line, current;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string key;
    if (!(ss >> key) || key[0] == '#') continue;
    if (key == "v") { glm::vec3 p; ss >> p.x >> p.y >> p.z; pos.push_back(p); }
    else if (key == "vn") { glm::vec3 n; ss >> n.x >> n.y >> n.z; nrm.push_back(n); }
    else if (key == "vt") { glm::vec2 t; ss >> t.x >> t.y; uvs.push_back({t.x, 1.0f - t.y}); }  // FlipUVs
    else if (key == "mtllib") { std::string f; ss >> f; auto m = readMtl(dirOf(path) + f); mtl.insert(m.begin(), m.end()); }
    else if (key == "usemtl") ss >> current;
    else if (key == "f") {
      Group& g = groupFor(current);
      std::vector<uint32_t> corner;
      std::string tok;
      while (ss >> tok) {
        std::array<int, 3> ref = {0, 0, 0};
        int field = 0;
        std::string num;
        for (size_t c = 0; c <= tok.size(); c++) {
          if (c == tok.size() || tok[c] == '/') {
            if (!num.empty() && field < 3) ref[field] = std::stoi(num);
            num.clear();
            field++;
          } else
            num += tok[c];
        }
        auto resolve = [](int i, size_t n) { return i > 0 ? i - 1 : (i < 0 ? int(n) + i : -1); };
        const std::array<int, 3> key3 = {resolve(ref[0], pos.size()), resolve(ref[1], uvs.size()), resolve(ref[2], nrm.size())};
        if (key3[0] < 0 || size_t(key3[0]) >= pos.size()) throw std::runtime_error("Failed to load scene: bad face index in " + path);
        auto it = g.lookup.find(key3);
        if (it == g.lookup.end()) {
          Vertex v;
          v.pos = pos[key3[0]];
          if (key3[1] >= 0 && size_t(key3[1]) < uvs.size()) v.texCoord = uvs[key3[1]];
          if (key3[2] >= 0 && size_t(key3[2]) < nrm.size()) v.normal = nrm[key3[2]]; else g.anyMissingNormal = true;
          it = g.lookup.emplace(key3, uint32_t(g.vertices.size())).first;
          g.vertices.push_back(v);
        }
        corner.push_back(it->second);
      }
      for (size_t k = 1; k + 1 < corner.size(); k++) {  // triangulate as a fan
        g.indices.push_back(corner[0]);
        g.indices.push_back(corner[k]);
        g.indices.push_back(corner[k + 1]);
      }
    }
  }
  std::vector<std::shared_ptr<Geometry>> out;
  for (auto& grp : groups) {
    if (grp.indices.empty()) {
      KF_WARN("A mesh in the file has no triangles: " + path);
      continue;
    }
    NiceMaterial m;
    auto it = mtl.find(grp.material);
    if (it != mtl.end()) m = it->second.mat;
    else { m.diffuseColor = glm::vec3(0.6f); m.roughness = 1.f; m.emission = glm::vec3(0.0f); m.emissionStrength = 1.f; }
    auto g = std::make_shared<Geometry>();
    g->path = path;
    g->dynamic = dynamic;
    g->vertices = std::move(grp.vertices);
    g->indices = std::move(grp.indices);
    g->initialized = false;
    g->matIndex.assign(g->indices.size() / 3, globalMaterialIndex(m));
    g->isOpaque = global::materials[g->matIndex.front()].alpha >= 1.F;
    if (grp.anyMissingNormal) g->recalculateNormals();  // aiProcess_GenNormals
    out.push_back(std::move(g));
  }
  return out;
}

std::shared_ptr<Geometry> loadObj(std::string_view path, bool dynamic) {
  auto scene = loadScene(path, dynamic);
  KF_ASSERT(scene.size() == 1, "complex scene! use loadScene");
  return scene.front();
}
}  // namespace kuafu
