// glTF 2.0 importer (.gltf + external / base64 buffers, .glb) behind loadScene() (reference
// src/core/geometry.cpp:45-232 through Assimp's glTF2 importer with aiProcess_Triangulate | GenNormals |
// FlipUVs | PreTransformVertices).  One Geometry per mesh primitive of every node of the default
// scene, node transforms baked in, vertices as the accessors hold them (indexed), v -> 1 - v.
// Materials as the reference reads them back from Assimp: COLOR_DIFFUSE = baseColorFactor.rgb,
// OPACITY = baseColorFactor.a unless alphaMode is OPAQUE, METALLIC_FACTOR, ROUGHNESS_FACTOR (no
// shininess rule for glTF, geometry.cpp:108), COLOR_EMISSIVE, REFRACTI / TRANSMISSION_FACTOR /
// SPECULAR_FACTOR from KHR_materials_ior / _transmission / _specular; the base-colour texture as the
// diffuse texture and the metallic-roughness texture as both the metalness and the roughness texture.
// Embedded (bufferView) images are skipped with the reference's warning.  Not covered: sparse
// accessors, skins, morph targets, cameras, KHR_draco / meshopt compression (fail loudly).
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "core/context/global.hpp"
#include "core/geometry.hpp"

namespace kuafu {
uint32_t importMaterialIndex(const NiceMaterial& m);  // geometry.cpp

namespace {
// ---- JSON ---------------------------------------------------------------------------------------
struct Json {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;
  const Json* get(const char* key) const {
    if (kind != Object) return nullptr;
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  size_t size() const { return kind == Array ? arr.size() : 0; }
  const Json& at(size_t i) const {
    static const Json none;
    return (kind == Array && i < arr.size()) ? arr[i] : none;
  }
  double number(const char* key, double dflt) const {
    const Json* j = get(key);
    return (j && j->kind == Number) ? j->num : dflt;
  }
  long index(const char* key) const {  // -1 when absent
    const Json* j = get(key);
    return (j && j->kind == Number) ? long(j->num) : -1;
  }
  std::string text(const char* key, const char* dflt = "") const {
    const Json* j = get(key);
    return (j && j->kind == String) ? j->str : std::string(dflt);
  }
};

struct JsonParser {
  const char* p;
  const char* end;
  std::string file;
  [[noreturn]] void fail(const char* what) { throw std::runtime_error("Failed to load scene: malformed JSON (" + std::string(what) + "), " + file); }
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
  Json value(int depth = 0) {
    if (depth > 128) fail("nesting");
    ws();
    if (p >= end) fail("unexpected end");
    Json j;
    if (*p == '{') {
      j.kind = Json::Object;
      p++;
      ws();
      if (p < end && *p == '}') { p++; return j; }
      for (;;) {
        ws();
        if (p >= end || *p != '"') fail("object key");
        std::string k = string();
        ws();
        if (p >= end || *p != ':') fail("colon");
        p++;
        j.obj.emplace_back(std::move(k), value(depth + 1));
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') { p++; return j; }
        fail("object");
      }
    }
    if (*p == '[') {
      j.kind = Json::Array;
      p++;
      ws();
      if (p < end && *p == ']') { p++; return j; }
      for (;;) {
        j.arr.push_back(value(depth + 1));
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') { p++; return j; }
        fail("array");
      }
    }
    if (*p == '"') { j.kind = Json::String; j.str = string(); return j; }
    if (end - p >= 4 && !std::memcmp(p, "true", 4)) { p += 4; j.kind = Json::Bool; j.b = true; return j; }
    if (end - p >= 5 && !std::memcmp(p, "false", 5)) { p += 5; j.kind = Json::Bool; return j; }
    if (end - p >= 4 && !std::memcmp(p, "null", 4)) { p += 4; return j; }
    char* q = nullptr;
    j.num = std::strtod(p, &q);
    if (q == p) fail("value");
    j.kind = Json::Number;
    p = q;
    return j;
  }
  std::string string() {
    std::string s;
    p++;  // opening quote
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) {
        p++;
        switch (*p) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'b': s += '\b'; break;
          case 'f': s += '\f'; break;
          case 'u': {  // BMP code point -> UTF-8
            if (end - p < 5) fail("escape");
            unsigned cp = unsigned(std::strtoul(std::string(p + 1, 4).c_str(), nullptr, 16));
            p += 4;
            if (cp < 0x80) s += char(cp);
            else if (cp < 0x800) { s += char(0xC0 | (cp >> 6)); s += char(0x80 | (cp & 0x3F)); }
            else { s += char(0xE0 | (cp >> 12)); s += char(0x80 | ((cp >> 6) & 0x3F)); s += char(0x80 | (cp & 0x3F)); }
            break;
          }
          default: s += *p;
        }
        p++;
      } else {
        s += *p++;
      }
    }
    if (p >= end) fail("string");
    p++;
    return s;
  }
};

std::vector<uint8_t> base64(const std::string& in, size_t from) {
  std::vector<uint8_t> out;
  uint32_t acc = 0;
  int bits = 0;
  for (size_t i = from; i < in.size(); i++) {
    const char c = in[i];
    int v;
    if (c >= 'A' && c <= 'Z') v = c - 'A';
    else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
    else if (c >= '0' && c <= '9') v = c - '0' + 52;
    else if (c == '+' || c == '-') v = 62;
    else if (c == '/' || c == '_') v = 63;
    else continue;  // padding / whitespace
    acc = (acc << 6) | uint32_t(v);
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out.push_back(uint8_t((acc >> bits) & 0xFFu));
    }
  }
  return out;
}

bool readBytes(const std::string& path, std::vector<uint8_t>& out) {
  std::ifstream in(path, std::ios::binary);
  if (!in.good()) return false;
  in.seekg(0, std::ios::end);
  const std::streamoff n = in.tellg();
  in.seekg(0);
  out.resize(size_t(n));
  in.read(reinterpret_cast<char*>(out.data()), n);
  return bool(in);
}

std::string dirOfFile(const std::string& p) {
  const size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}

struct Gltf {
  std::string file, dir;
  Json doc;
  std::vector<std::vector<uint8_t>> buffers;
  [[noreturn]] void fail(const std::string& what) const { throw std::runtime_error("Failed to load scene: " + what + ", " + file); }

  // element i, component c of an accessor, as float (normalised integers scaled as the spec says)
  struct View {
    const uint8_t* base = nullptr;
    size_t stride = 0, count = 0;
    int comps = 0, type = 0;
    bool normalized = false;
  };
  View view(long accessor) const {
    const Json& acc = doc.get("accessors") ? doc.get("accessors")->at(size_t(accessor)) : Json();
    if (acc.kind != Json::Object) fail("bad accessor index");
    if (acc.get("sparse")) fail("sparse accessors are not supported");
    View v;
    v.type = int(acc.number("componentType", 0));
    v.count = size_t(acc.number("count", 0));
    v.normalized = acc.get("normalized") && acc.get("normalized")->b;
    const std::string t = acc.text("type");
    v.comps = t == "SCALAR" ? 1 : t == "VEC2" ? 2 : t == "VEC3" ? 3 : t == "VEC4" ? 4 : t == "MAT4" ? 16 : 0;
    const size_t csize = (v.type == 5120 || v.type == 5121) ? 1 : (v.type == 5122 || v.type == 5123) ? 2 : (v.type == 5125 || v.type == 5126) ? 4 : 0;
    if (!v.comps || !csize) fail("unsupported accessor type");
    const long bv = acc.index("bufferView");
    if (bv < 0) fail("accessor without a bufferView");
    const Json& view = doc.get("bufferViews") ? doc.get("bufferViews")->at(size_t(bv)) : Json();
    const long b = view.index("buffer");
    if (b < 0 || size_t(b) >= buffers.size()) fail("bad buffer index");
    const size_t off = size_t(view.number("byteOffset", 0)) + size_t(acc.number("byteOffset", 0));
    v.stride = size_t(view.number("byteStride", 0));
    if (!v.stride) v.stride = csize * size_t(v.comps);
    if (v.count && off + (v.count - 1) * v.stride + csize * size_t(v.comps) > buffers[size_t(b)].size()) fail("accessor exceeds its buffer");
    v.base = buffers[size_t(b)].data() + off;
    return v;
  }
  static float component(const View& v, size_t i, int c) {
    const uint8_t* p = v.base + i * v.stride;
    switch (v.type) {
      case 5126: { float f; std::memcpy(&f, p + 4 * c, 4); return f; }
      case 5121: { const uint8_t x = p[c]; return v.normalized ? float(x) / 255.0f : float(x); }
      case 5123: { uint16_t x; std::memcpy(&x, p + 2 * c, 2); return v.normalized ? float(x) / 65535.0f : float(x); }
      case 5120: { const int8_t x = int8_t(p[c]); return v.normalized ? std::fmax(float(x) / 127.0f, -1.0f) : float(x); }
      case 5122: { int16_t x; std::memcpy(&x, p + 2 * c, 2); return v.normalized ? std::fmax(float(x) / 32767.0f, -1.0f) : float(x); }
      case 5125: { uint32_t x; std::memcpy(&x, p + 4 * c, 4); return float(x); }
    }
    return 0.0f;
  }
  static uint32_t indexAt(const View& v, size_t i) {
    const uint8_t* p = v.base + i * v.stride;
    if (v.type == 5121) return p[0];
    if (v.type == 5123) { uint16_t x; std::memcpy(&x, p, 2); return x; }
    uint32_t x;
    std::memcpy(&x, p, 4);
    return x;
  }
};

glm::mat4 localTransform(const Json& node) {
  if (const Json* m = node.get("matrix"))
    if (m->size() == 16) {
      glm::mat4 t;  // glTF matrices are column-major, like glm
      for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) t[c][r] = float(m->at(size_t(4 * c + r)).num);
      return t;
    }
  glm::mat4 m(1.0f);
  if (const Json* t = node.get("translation"))
    if (t->size() == 3) m = glm::translate(m, glm::vec3(float(t->at(0).num), float(t->at(1).num), float(t->at(2).num)));
  if (const Json* q = node.get("rotation"))
    if (q->size() == 4) {
      const float x = float(q->at(0).num), y = float(q->at(1).num), z = float(q->at(2).num), w = float(q->at(3).num);
      glm::mat4 r(1.0f);
      r[0][0] = 1 - 2 * (y * y + z * z); r[0][1] = 2 * (x * y + z * w);     r[0][2] = 2 * (x * z - y * w);
      r[1][0] = 2 * (x * y - z * w);     r[1][1] = 1 - 2 * (x * x + z * z); r[1][2] = 2 * (y * z + x * w);
      r[2][0] = 2 * (x * z + y * w);     r[2][1] = 2 * (y * z - x * w);     r[2][2] = 1 - 2 * (x * x + y * y);
      m = m * r;
    }
  if (const Json* s = node.get("scale"))
    if (s->size() == 3) m = glm::scale(m, glm::vec3(float(s->at(0).num), float(s->at(1).num), float(s->at(2).num)));
  return m;
}

bool identity(const glm::mat4& m) { return m == glm::mat4(1.0f); }

std::string texturePath(const Gltf& g, const Json* texInfo) {
  if (!texInfo) return "";
  const long ti = texInfo->index("index");
  const Json* textures = g.doc.get("textures");
  if (ti < 0 || !textures) return "";
  const long src = textures->at(size_t(ti)).index("source");
  const Json* images = g.doc.get("images");
  if (src < 0 || !images) return "";
  const Json& img = images->at(size_t(src));
  const std::string uri = img.text("uri");
  if (uri.empty() || uri.rfind("data:", 0) == 0) {
    KF_WARN("embedded texture not supported");
    return "";
  }
  return uri[0] == '/' ? uri : g.dir + uri;
}

NiceMaterial gltfMaterial(const Gltf& g, long index) {
  NiceMaterial m;  // the reference's defaults before it queries Assimp (geometry.cpp:75-84)
  m.diffuseColor = glm::vec3(0.0f);
  m.emission = glm::vec3(0.0f);
  m.emissionStrength = 1.f;
  m.alpha = 1.f;
  m.ior = 1.4f;
  m.specular = .5f;
  m.transmission = 0.f;
  m.metallic = 0.f;
  m.roughness = 0.f;
  const Json* mats = g.doc.get("materials");
  if (index < 0 || !mats) {  // Assimp's default material: grey 0.6
    m.diffuseColor = glm::vec3(0.6f);
    return m;
  }
  const Json& mat = mats->at(size_t(index));
  float base[4] = {1, 1, 1, 1};
  m.metallic = 1.f;
  m.roughness = 1.f;
  if (const Json* pbr = mat.get("pbrMetallicRoughness")) {
    if (const Json* c = pbr->get("baseColorFactor"))
      for (size_t k = 0; k < 4 && k < c->size(); k++) base[k] = float(c->at(k).num);
    m.metallic = float(pbr->number("metallicFactor", 1.0));
    m.roughness = float(pbr->number("roughnessFactor", 1.0));
    m.diffuseTexPath = texturePath(g, pbr->get("baseColorTexture"));
    const std::string mr = texturePath(g, pbr->get("metallicRoughnessTexture"));
    m.metallicTexPath = mr;
    m.roughnessTexPath = mr;
  }
  m.diffuseColor = glm::vec3(base[0], base[1], base[2]);
  if (mat.text("alphaMode", "OPAQUE") != "OPAQUE") m.alpha = base[3];
  if (const Json* e = mat.get("emissiveFactor"))
    if (e->size() == 3) m.emission = glm::vec3(float(e->at(0).num), float(e->at(1).num), float(e->at(2).num));
  if (const Json* ext = mat.get("extensions")) {
    if (const Json* x = ext->get("KHR_materials_ior")) m.ior = float(x->number("ior", 1.5));
    if (const Json* x = ext->get("KHR_materials_transmission")) m.transmission = float(x->number("transmissionFactor", 0.0));
    if (const Json* x = ext->get("KHR_materials_specular")) m.specular = float(x->number("specularFactor", 1.0));
  }
  if (mat.get("normalTexture")) KF_WARN("normals texture not supported");
  return m;
}

void emitPrimitive(const Gltf& g, const Json& prim, const glm::mat4& world, bool dynamic,
                   std::vector<std::shared_ptr<Geometry>>& out) {
  const int mode = int(prim.number("mode", 4));
  if (mode != 4 && mode != 5 && mode != 6) {
    KF_WARN("Mesh not triangulated!");
    return;
  }
  if (const Json* ext = prim.get("extensions"))
    if (ext->get("KHR_draco_mesh_compression")) g.fail("KHR_draco_mesh_compression is not supported");
  const Json* attrs = prim.get("attributes");
  const long posAcc = attrs ? attrs->index("POSITION") : -1;
  if (posAcc < 0) return;
  const Gltf::View pos = g.view(posAcc);
  Gltf::View nrm, uv, col;
  const long nrmAcc = attrs->index("NORMAL"), uvAcc = attrs->index("TEXCOORD_0"), colAcc = attrs->index("COLOR_0");
  if (nrmAcc >= 0) nrm = g.view(nrmAcc);
  if (uvAcc >= 0) uv = g.view(uvAcc);
  if (colAcc >= 0) col = g.view(colAcc);
  auto geom = std::make_shared<Geometry>();
  geom->path = g.file;
  geom->dynamic = dynamic;
  geom->initialized = false;
  const bool bake = !identity(world);
  // inverse transpose of the upper 3x3 (cofactors / determinant), for the normals
  const glm::vec3 a(world[0][0], world[0][1], world[0][2]), b(world[1][0], world[1][1], world[1][2]), c(world[2][0], world[2][1], world[2][2]);
  const glm::vec3 r0 = glm::cross(b, c), r1 = glm::cross(c, a), r2 = glm::cross(a, b);
  const float det = glm::dot(a, r0), idet = det != 0.0f ? 1.0f / det : 0.0f;
  geom->vertices.resize(pos.count);
  for (size_t i = 0; i < pos.count; i++) {
    Vertex& v = geom->vertices[i];
    v.pos = glm::vec3(Gltf::component(pos, i, 0), Gltf::component(pos, i, 1), Gltf::component(pos, i, 2));
    if (nrm.base && i < nrm.count) v.normal = glm::vec3(Gltf::component(nrm, i, 0), Gltf::component(nrm, i, 1), Gltf::component(nrm, i, 2));
    if (uv.base && i < uv.count) v.texCoord = glm::vec2(Gltf::component(uv, i, 0), 1.0f - Gltf::component(uv, i, 1));  // FlipUVs
    if (col.base && i < col.count) v.color = glm::vec3(Gltf::component(col, i, 0), Gltf::component(col, i, 1), Gltf::component(col, i, 2));
    if (bake) {
      const glm::vec4 wp = world * glm::vec4(v.pos, 1.0f);
      v.pos = glm::vec3(wp.x, wp.y, wp.z);
      if (nrm.base) v.normal = glm::normalize((r0 * v.normal.x + r1 * v.normal.y + r2 * v.normal.z) * idet);
    }
  }
  std::vector<uint32_t> seq;
  const long idxAcc = prim.index("indices");
  if (idxAcc >= 0) {
    const Gltf::View iv = g.view(idxAcc);
    seq.resize(iv.count);
    for (size_t i = 0; i < iv.count; i++) seq[i] = Gltf::indexAt(iv, i);
  } else {
    seq.resize(pos.count);
    for (size_t i = 0; i < pos.count; i++) seq[i] = uint32_t(i);
  }
  for (uint32_t i : seq)
    if (i >= pos.count) g.fail("vertex index out of range");
  if (mode == 4) {
    seq.resize(seq.size() / 3 * 3);
    geom->indices = std::move(seq);
  } else {
    for (size_t k = 2; k < seq.size(); k++) {
      if (mode == 5) {  // strip: winding alternates
        const bool odd = (k & 1u) != 0;
        geom->indices.push_back(seq[k - 2 + (odd ? 1 : 0)]);
        geom->indices.push_back(seq[k - 1 - (odd ? 1 : 0)]);
        geom->indices.push_back(seq[k]);
      } else {  // fan
        geom->indices.push_back(seq[0]);
        geom->indices.push_back(seq[k - 1]);
        geom->indices.push_back(seq[k]);
      }
    }
  }
  if (geom->vertices.empty() || geom->indices.empty()) {
    KF_WARN("A mesh in the file has no triangles: " + g.file);
    return;
  }
  if (!nrm.base) geom->recalculateNormals();  // aiProcess_GenNormals
  geom->matIndex.assign(geom->indices.size() / 3, importMaterialIndex(gltfMaterial(g, prim.index("material"))));
  geom->isOpaque = global::materials[geom->matIndex.front()].alpha >= 1.F;
  out.push_back(std::move(geom));
}

void walk(const Gltf& g, size_t nodeIndex, const glm::mat4& parent, bool dynamic, int depth,
          std::vector<std::shared_ptr<Geometry>>& out) {
  if (depth > 64) g.fail("node hierarchy too deep");
  const Json* nodes = g.doc.get("nodes");
  if (!nodes || nodeIndex >= nodes->size()) g.fail("bad node index");
  const Json& node = nodes->at(nodeIndex);
  const glm::mat4 world = parent * localTransform(node);
  const long mesh = node.index("mesh");
  if (mesh >= 0) {
    const Json* meshes = g.doc.get("meshes");
    if (!meshes || size_t(mesh) >= meshes->size()) g.fail("bad mesh index");
    if (const Json* prims = meshes->at(size_t(mesh)).get("primitives"))
      for (size_t k = 0; k < prims->size(); k++) emitPrimitive(g, prims->at(k), world, dynamic, out);
  }
  if (node.get("skin")) KF_WARN("skinned geometry is imported in its bind pose: " + g.file);
  if (const Json* kids = node.get("children"))
    for (size_t k = 0; k < kids->size(); k++) walk(g, size_t(kids->at(k).num), world, dynamic, depth + 1, out);
}
}  // namespace

std::vector<std::shared_ptr<Geometry>> loadGltfScene(const std::string& path, bool dynamic) {
  std::vector<uint8_t> file;
  if (!readBytes(path, file)) throw std::runtime_error("Failed to load scene: cannot open " + path);
  Gltf g;
  g.file = path;
  g.dir = dirOfFile(path);
  std::vector<uint8_t> glbBin;
  bool haveGlbBin = false;
  std::string json;
  if (file.size() >= 12 && !std::memcmp(file.data(), "glTF", 4)) {  // binary container
    uint32_t version = 0, length = 0;
    std::memcpy(&version, file.data() + 4, 4);
    std::memcpy(&length, file.data() + 8, 4);
    if (version != 2) g.fail("only glTF 2.0 is supported");
    size_t off = 12;
    while (off + 8 <= file.size()) {
      uint32_t clen = 0, ctype = 0;
      std::memcpy(&clen, file.data() + off, 4);
      std::memcpy(&ctype, file.data() + off + 4, 4);
      off += 8;
      if (off + clen > file.size()) g.fail("truncated GLB chunk");
      if (ctype == 0x4E4F534Au) json.assign(reinterpret_cast<const char*>(file.data() + off), clen);
      else if (ctype == 0x004E4942u && !haveGlbBin) { glbBin.assign(file.begin() + long(off), file.begin() + long(off + clen)); haveGlbBin = true; }
      off += (size_t(clen) + 3u) & ~size_t(3);
    }
  } else {
    json.assign(reinterpret_cast<const char*>(file.data()), file.size());
  }
  JsonParser jp{json.data(), json.data() + json.size(), path};
  g.doc = jp.value();
  if (g.doc.kind != Json::Object) g.fail("not a glTF document");
  if (const Json* asset = g.doc.get("asset"))
    if (asset->text("version", "2.0").rfind("2.", 0) != 0) g.fail("only glTF 2.0 is supported");
  if (const Json* req = g.doc.get("extensionsRequired"))
    for (size_t k = 0; k < req->size(); k++) {
      const std::string& e = req->at(k).str;
      if (e == "KHR_draco_mesh_compression" || e == "EXT_meshopt_compression") g.fail(e + " is not supported");
    }
  if (const Json* bufs = g.doc.get("buffers"))
    for (size_t k = 0; k < bufs->size(); k++) {
      const std::string uri = bufs->at(k).text("uri");
      std::vector<uint8_t> data;
      if (uri.empty()) {
        if (!(k == 0 && haveGlbBin)) g.fail("buffer without a uri");
        data = glbBin;
      } else if (uri.rfind("data:", 0) == 0) {
        const size_t comma = uri.find(',');
        if (comma == std::string::npos) g.fail("malformed data uri");
        data = base64(uri, comma + 1);
      } else if (!readBytes(uri[0] == '/' ? uri : g.dir + uri, data)) {
        g.fail("cannot open buffer " + uri);
      }
      g.buffers.push_back(std::move(data));
    }
  std::vector<std::shared_ptr<Geometry>> out;
  const Json* scenes = g.doc.get("scenes");
  const Json* nodes = g.doc.get("nodes");
  if (scenes && scenes->size()) {
    long si = g.doc.index("scene");
    if (si < 0 || size_t(si) >= scenes->size()) si = 0;
    if (const Json* roots = scenes->at(size_t(si)).get("nodes"))
      for (size_t k = 0; k < roots->size(); k++) walk(g, size_t(roots->at(k).num), glm::mat4(1.0f), dynamic, 0, out);
  } else if (nodes) {  // no scene: every parentless node is a root
    std::vector<bool> child(nodes->size(), false);
    for (size_t k = 0; k < nodes->size(); k++)
      if (const Json* kids = nodes->at(k).get("children"))
        for (size_t c = 0; c < kids->size(); c++)
          if (size_t(kids->at(c).num) < child.size()) child[size_t(kids->at(c).num)] = true;
    for (size_t k = 0; k < nodes->size(); k++)
      if (!child[k]) walk(g, k, glm::mat4(1.0f), dynamic, 0, out);
  }
  if (out.empty()) throw std::runtime_error("Failed to load scene: no triangle meshes in " + path);
  return out;
}
}  // namespace kuafu
