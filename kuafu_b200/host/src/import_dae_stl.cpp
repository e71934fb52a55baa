// COLLADA (.dae) and STL importers behind loadScene() (reference src/core/geometry.cpp:45-232, where
// Assimp does this work with aiProcess_Triangulate | GenNormals | FlipUVs | PreTransformVertices and
// AI_CONFIG_IMPORT_COLLADA_IGNORE_UP_DIRECTION).  What a mesh file turns into follows that call:
//   * one Geometry per primitive group of every instantiated <geometry>, node transforms baked in
//     (PreTransformVertices), the up axis left alone;
//   * vertices are per face corner (Assimp does not join identical vertices without
//     aiProcess_JoinIdenticalVertices), indices 0..3T-1 -- suzanne.dae: 251 904 triangles,
//     755 712 vertices = the 36.3 MB vertex buffer of SURVEY.md row a13;
//   * v -> 1 - v (FlipUVs); missing normals are generated (GenNormals);
//   * materials: diffuse colour / texture, index of refraction, opacity with the reference's ".dae
//     fully transparent -> opaque" repair, roughness from shininess (geometry.cpp:100-132).
// Covered: <triangles>, <polylist>, <polygons> with <p> children, sources with any stride, node
// trees with <matrix>/<translate>/<rotate>/<scale>, <instance_node>, bind_material.  Not covered
// (fails loudly or is ignored with a warning): skins / morphs, splines, <lines>, embedded textures.
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

#include <stdexcept>

#include "core/context/global.hpp"
#include "core/geometry.hpp"

namespace kuafu {
uint32_t importMaterialIndex(const NiceMaterial& m);  // geometry.cpp: de-duplicating registry lookup

namespace {
// ---- a small XML tree: element names, attributes, children, and the raw text of leaf elements ----
struct Xml {
  std::string_view name;
  std::vector<std::pair<std::string_view, std::string_view>> attrs;
  std::vector<std::unique_ptr<Xml>> kids;
  std::string_view text;  // character data before the first child (all of it for leaf elements)
  std::string_view attr(std::string_view k) const {
    for (auto& a : attrs)
      if (a.first == k) return a.second;
    return {};
  }
  const Xml* child(std::string_view n) const {
    for (auto& k : kids)
      if (k->name == n) return k.get();
    return nullptr;
  }
  template <class F>
  void each(std::string_view n, F f) const {
    for (auto& k : kids)
      if (k->name == n) f(*k);
  }
};

struct XmlParser {
  const char* p;
  const char* end;
  std::string file;
  [[noreturn]] void fail(const char* what) { throw std::runtime_error("Failed to load scene: malformed XML (" + std::string(what) + "), " + file); }
  void skipWs() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
  static bool nameChar(char c) { return c != ' ' && c != '\n' && c != '\r' && c != '\t' && c != '>' && c != '/' && c != '='; }
  // skips declarations, comments and doctype; leaves p at '<' of the next element or at end
  void skipMisc() {
    for (;;) {
      skipWs();
      if (end - p >= 4 && !std::memcmp(p, "<!--", 4)) {
        const char* e = std::strstr(p, "-->");
        if (!e) fail("comment");
        p = e + 3;
      } else if (end - p >= 2 && p[0] == '<' && (p[1] == '?' || p[1] == '!')) {
        while (p < end && *p != '>') p++;
        if (p < end) p++;
      } else {
        return;
      }
    }
  }
  std::unique_ptr<Xml> element() {
    if (p >= end || *p != '<') fail("element start");
    p++;
    auto e = std::make_unique<Xml>();
    const char* s = p;
    while (p < end && nameChar(*p)) p++;
    e->name = std::string_view(s, size_t(p - s));
    for (;;) {
      skipWs();
      if (p >= end) fail("tag");
      if (*p == '/') {
        if (p + 1 >= end || p[1] != '>') fail("empty tag");
        p += 2;
        return e;
      }
      if (*p == '>') { p++; break; }
      const char* k = p;
      while (p < end && nameChar(*p)) p++;
      std::string_view key(k, size_t(p - k));
      skipWs();
      if (p >= end || *p != '=') fail("attribute");
      p++;
      skipWs();
      if (p >= end || (*p != '"' && *p != '\'')) fail("attribute value");
      const char q = *p++;
      const char* v = p;
      while (p < end && *p != q) p++;
      if (p >= end) fail("attribute value");
      e->attrs.emplace_back(key, std::string_view(v, size_t(p - v)));
      p++;
    }
    const char* t = p;
    bool haveText = false;
    for (;;) {
      while (p < end && *p != '<') p++;
      if (p >= end) fail("unterminated element");
      if (!haveText) { e->text = std::string_view(t, size_t(p - t)); haveText = true; }
      if (p + 1 < end && p[1] == '/') {
        p += 2;
        const char* c = p;
        while (p < end && *p != '>') p++;
        std::string_view close(c, size_t(p - c));
        while (!close.empty() && (close.back() == ' ' || close.back() == '\n')) close.remove_suffix(1);
        if (close != e->name) fail("mismatched close tag");
        if (p < end) p++;
        return e;
      }
      if (p + 3 < end && !std::memcmp(p, "<!--", 4)) {
        const char* c = std::strstr(p, "-->");
        if (!c) fail("comment");
        p = c + 3;
        continue;
      }
      e->kids.push_back(element());
    }
  }
};

// whitespace-separated numbers of a text node (the buffer is NUL-terminated, '<' stops strtof)
void parseFloats(std::string_view t, std::vector<float>& out) {
  const char* p = t.data();
  const char* end = p + t.size();
  while (p < end) {
    char* q = nullptr;
    const float v = std::strtof(p, &q);
    if (q == p) break;
    out.push_back(v);
    p = q;
  }
}
void parseUints(std::string_view t, std::vector<uint32_t>& out) {
  const char* p = t.data();
  const char* end = p + t.size();
  while (p < end) {
    char* q = nullptr;
    const unsigned long v = std::strtoul(p, &q, 10);
    if (q == p) break;
    out.push_back(uint32_t(v));
    p = q;
  }
}
std::string_view stripHash(std::string_view s) { return (!s.empty() && s[0] == '#') ? s.substr(1) : s; }

struct Source {
  std::vector<float> data;
  uint32_t stride = 1;
};
struct Input {
  std::string_view semantic, source;
  uint32_t offset = 0, set = 0;
};

glm::mat4 nodeLocalTransform(const Xml& node) {
  glm::mat4 m(1.0f);
  for (auto& k : node.kids) {
    std::vector<float> v;
    if (k->name == "matrix") {
      parseFloats(k->text, v);
      if (v.size() == 16) {
        glm::mat4 t;  // COLLADA matrices are row-major; glm is column-major
        for (int r = 0; r < 4; r++)
          for (int c = 0; c < 4; c++) t[c][r] = v[size_t(4 * r + c)];
        m = m * t;
      }
    } else if (k->name == "translate") {
      parseFloats(k->text, v);
      if (v.size() == 3) m = glm::translate(m, glm::vec3(v[0], v[1], v[2]));
    } else if (k->name == "rotate") {
      parseFloats(k->text, v);
      if (v.size() == 4) m = glm::rotate(m, glm::radians(v[3]), glm::vec3(v[0], v[1], v[2]));
    } else if (k->name == "scale") {
      parseFloats(k->text, v);
      if (v.size() == 3) m = glm::scale(m, glm::vec3(v[0], v[1], v[2]));
    }
  }
  return m;
}

bool isIdentity(const glm::mat4& m) {
  const glm::mat4 id(1.0f);
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++)
      if (m[c][r] != id[c][r]) return false;
  return true;
}

// Inverse transpose of the upper 3x3 of m, as its three columns (cofactors / determinant).
struct NormalMatrix {
  glm::vec3 c0, c1, c2;
  glm::vec3 apply(const glm::vec3& n) const { return c0 * n.x + c1 * n.y + c2 * n.z; }
};
NormalMatrix inverseTranspose3(const glm::mat4& m) {
  const glm::vec3 a(m[0][0], m[0][1], m[0][2]), b(m[1][0], m[1][1], m[1][2]), c(m[2][0], m[2][1], m[2][2]);
  const glm::vec3 r0 = glm::cross(b, c), r1 = glm::cross(c, a), r2 = glm::cross(a, b);
  const float det = glm::dot(a, r0);
  const float id = det != 0.0f ? 1.0f / det : 0.0f;
  return {r0 * id, r1 * id, r2 * id};
}

struct DaeDoc {
  std::string file, dir;
  std::unique_ptr<Xml> root;
  std::map<std::string_view, const Xml*> geometries, materials, effects, images, nodes, scenes;
};

NiceMaterial daeMaterial(const DaeDoc& d, std::string_view materialId) {
  NiceMaterial m;  // reference defaults (geometry.cpp:75-84)
  m.diffuseColor = glm::vec3(0.0f);
  m.emission = glm::vec3(0.0f);
  m.emissionStrength = 1.f;
  m.alpha = 1.f;
  m.ior = 1.4f;
  m.specular = .5f;
  m.transmission = 0.f;
  m.metallic = 0.f;
  m.roughness = 0.f;
  float shininess = -1.f;
  bool haveOpacity = false;
  auto mi = d.materials.find(materialId);
  const Xml* effect = nullptr;
  if (mi != d.materials.end())
    if (const Xml* ie = mi->second->child("instance_effect")) {
      auto ei = d.effects.find(stripHash(ie->attr("url")));
      if (ei != d.effects.end()) effect = ei->second;
    }
  const Xml* profile = effect ? effect->child("profile_COMMON") : nullptr;
  const Xml* technique = profile ? profile->child("technique") : nullptr;
  const Xml* shader = nullptr;
  if (technique)
    for (const char* n : {"phong", "lambert", "blinn", "constant"})
      if (!shader) shader = technique->child(n);
  auto colour = [](const Xml* e, glm::vec4& out) {
    if (!e) return false;
    const Xml* c = e->child("color");
    if (!c) return false;
    std::vector<float> v;
    parseFloats(c->text, v);
    if (v.size() < 3) return false;
    out = glm::vec4(v[0], v[1], v[2], v.size() > 3 ? v[3] : 1.0f);
    return true;
  };
  auto scalar = [](const Xml* e, float& out) {
    if (!e) return false;
    const Xml* f = e->child("float");
    if (!f) return false;
    std::vector<float> v;
    parseFloats(f->text, v);
    if (v.empty()) return false;
    out = v[0];
    return true;
  };
  if (shader) {
    glm::vec4 c;
    if (colour(shader->child("diffuse"), c)) m.diffuseColor = glm::vec3(c.x, c.y, c.z);
    if (colour(shader->child("emission"), c)) m.emission = glm::vec3(c.x, c.y, c.z);
    scalar(shader->child("index_of_refraction"), m.ior);
    scalar(shader->child("shininess"), shininess);
    // diffuse texture: <texture texture="sampler"> -> newparam sampler2D -> surface -> image
    if (const Xml* dif = shader->child("diffuse"))
      if (const Xml* tex = dif->child("texture")) {
        std::string_view ref = tex->attr("texture");
        for (int hop = 0; hop < 2 && profile; hop++)
          profile->each("newparam", [&](const Xml& np) {
            if (np.attr("sid") != ref) return;
            if (const Xml* s = np.child("sampler2D")) {
              if (const Xml* src = s->child("source")) ref = src->text;
              else if (const Xml* ii = s->child("instance_image")) ref = stripHash(ii->attr("url"));
            } else if (const Xml* s2 = np.child("surface")) {
              if (const Xml* init = s2->child("init_from")) ref = init->text;
            }
          });
        auto ii = d.images.find(ref);
        if (ii != d.images.end()) {
          const Xml* init = ii->second->child("init_from");
          std::string_view path = init ? (init->child("ref") ? init->child("ref")->text : init->text) : std::string_view();
          while (!path.empty() && (path.front() == ' ' || path.front() == '\n')) path.remove_prefix(1);
          while (!path.empty() && (path.back() == ' ' || path.back() == '\n')) path.remove_suffix(1);
          if (path.rfind("file://", 0) == 0) path.remove_prefix(7);
          if (!path.empty()) m.diffuseTexPath = path[0] == '/' ? std::string(path) : d.dir + std::string(path);
        }
      }
    // opacity (Assimp ColladaLoader: A_ONE -> transparent.a * transparency, RGB_ZERO -> 1 - luminance * t)
    glm::vec4 tc(1.0f);
    float tf = 1.0f;
    const Xml* transparent = shader->child("transparent");
    const bool haveT = colour(transparent, tc), haveF = scalar(shader->child("transparency"), tf);
    if (haveT || haveF) {
      const bool rgbZero = transparent && transparent->attr("opaque") == "RGB_ZERO";
      m.alpha = rgbZero ? 1.0f - (0.212671f * tc.x + 0.715160f * tc.y + 0.072169f * tc.z) * tf : tc.w * tf;
      haveOpacity = true;
    }
  }
  if (haveOpacity && m.alpha < 1e-5f) {
    KF_WARN("The DAE file " + d.file + " is fully transparent. This is probably due to modeling error. Setting opacity to 1 instead.");
    m.alpha = 1.f;
  }
  if (m.roughness == 0 && shininess > 0) {  // geometry.cpp:112-123
    m.roughness = shininess <= 5.f ? 1.f : 1.f - (std::sqrt(shininess - 5.f) * 0.025f);
    if (m.roughness < 0.001f) m.roughness = 0.001f;
  }
  return m;
}

// One Geometry per primitive group of `geom`, transformed by `world`.
void emitGeometry(const DaeDoc& d, const Xml& geom, const glm::mat4& world,
                  const std::map<std::string_view, std::string_view>& materialBinding, bool dynamic,
                  std::vector<std::shared_ptr<Geometry>>& out) {
  const Xml* mesh = geom.child("mesh");
  if (!mesh) {
    KF_WARN("A geometry in the file is not a mesh (skipped): " + d.file);
    return;
  }
  std::map<std::string_view, Source> sources;
  mesh->each("source", [&](const Xml& s) {
    Source src;
    if (const Xml* fa = s.child("float_array")) parseFloats(fa->text, src.data);
    if (const Xml* tc = s.child("technique_common"))
      if (const Xml* acc = tc->child("accessor")) {
        const std::string stride(acc->attr("stride"));
        if (!stride.empty()) src.stride = uint32_t(std::max(1, std::atoi(stride.c_str())));
      }
    sources.emplace(s.attr("id"), std::move(src));
  });
  std::map<std::string_view, std::string_view> vertexPositions;  // <vertices id> -> position source
  mesh->each("vertices", [&](const Xml& v) {
    v.each("input", [&](const Xml& in) {
      if (in.attr("semantic") == "POSITION") vertexPositions[v.attr("id")] = stripHash(in.attr("source"));
    });
  });
  const bool bake = !isIdentity(world);
  const NormalMatrix normalMat = inverseTranspose3(world);
  for (auto& prim : mesh->kids) {
    const bool tri = prim->name == "triangles", plist = prim->name == "polylist", pgons = prim->name == "polygons";
    if (!tri && !plist && !pgons) {
      if (prim->name == "lines" || prim->name == "linestrips" || prim->name == "tristrips" || prim->name == "trifans")
        KF_WARN("Mesh not triangulated!");
      continue;
    }
    std::vector<Input> inputs;
    uint32_t step = 0;
    prim->each("input", [&](const Xml& in) {
      Input i;
      i.semantic = in.attr("semantic");
      i.source = stripHash(in.attr("source"));
      i.offset = uint32_t(std::atoi(std::string(in.attr("offset")).c_str()));
      i.set = uint32_t(std::atoi(std::string(in.attr("set")).c_str()));
      step = std::max(step, i.offset + 1);
      inputs.push_back(i);
    });
    const Source *pos = nullptr, *nrm = nullptr, *uv = nullptr, *col = nullptr;
    uint32_t posOff = 0, nrmOff = 0, uvOff = 0, colOff = 0, uvSet = ~0u;
    for (auto& i : inputs) {
      auto find = [&](std::string_view id) -> const Source* {
        auto it = sources.find(id);
        return it == sources.end() ? nullptr : &it->second;
      };
      if (i.semantic == "VERTEX") {
        auto vp = vertexPositions.find(i.source);
        if (vp != vertexPositions.end()) { pos = find(vp->second); posOff = i.offset; }
      } else if (i.semantic == "NORMAL") { nrm = find(i.source); nrmOff = i.offset; }
      else if (i.semantic == "TEXCOORD" && i.set < uvSet) { uv = find(i.source); uvOff = i.offset; uvSet = i.set; }
      else if (i.semantic == "COLOR" && !col) { col = find(i.source); colOff = i.offset; }
    }
    if (!pos || step == 0) throw std::runtime_error("Failed to load scene: primitive without positions, " + d.file);
    std::vector<uint32_t> p, vcount;
    if (pgons) {
      prim->each("p", [&](const Xml& pe) {
        const size_t before = p.size();
        parseUints(pe.text, p);
        vcount.push_back(uint32_t((p.size() - before) / step));
      });
    } else {
      if (const Xml* pe = prim->child("p")) parseUints(pe->text, p);
      if (plist)
        if (const Xml* vc = prim->child("vcount")) parseUints(vc->text, vcount);
    }
    auto g = std::make_shared<Geometry>();
    g->path = d.file;
    g->dynamic = dynamic;
    g->initialized = false;
    auto corner = [&](size_t c) {  // c = index of the corner's index tuple in p
      const uint32_t* t = &p[c * step];
      Vertex v;
      const size_t pi = size_t(t[posOff]) * pos->stride;
      if (pi + 2 >= pos->data.size()) throw std::runtime_error("Failed to load scene: bad vertex index in " + d.file);
      v.pos = glm::vec3(pos->data[pi], pos->data[pi + 1], pos->data[pi + 2]);
      if (nrm) {
        const size_t ni = size_t(t[nrmOff]) * nrm->stride;
        if (ni + 2 < nrm->data.size()) v.normal = glm::vec3(nrm->data[ni], nrm->data[ni + 1], nrm->data[ni + 2]);
      }
      if (uv) {
        const size_t ui = size_t(t[uvOff]) * uv->stride;
        if (ui + 1 < uv->data.size()) v.texCoord = glm::vec2(uv->data[ui], 1.0f - uv->data[ui + 1]);  // FlipUVs
      }
      if (col) {
        const size_t ci = size_t(t[colOff]) * col->stride;
        if (ci + 2 < col->data.size()) v.color = glm::vec3(col->data[ci], col->data[ci + 1], col->data[ci + 2]);
      }
      if (bake) {
        const glm::vec4 wp = world * glm::vec4(v.pos, 1.0f);
        v.pos = glm::vec3(wp.x, wp.y, wp.z);
        if (nrm) v.normal = glm::normalize(normalMat.apply(v.normal));
      }
      g->indices.push_back(uint32_t(g->vertices.size()));
      g->vertices.push_back(v);
    };
    const size_t nCorners = p.size() / step;
    if (tri) {
      g->vertices.reserve(nCorners);
      g->indices.reserve(nCorners);
      for (size_t c = 0; c + 2 < nCorners; c += 3) { corner(c); corner(c + 1); corner(c + 2); }
    } else {
      size_t c = 0;
      for (uint32_t n : vcount) {
        if (c + n > nCorners) break;
        for (uint32_t k = 1; k + 1 < n; k++) { corner(c); corner(c + k); corner(c + k + 1); }  // fan
        c += n;
      }
    }
    if (g->indices.empty()) {
      KF_WARN("A mesh in the file has no triangles: " + d.file);
      continue;
    }
    if (!nrm) g->recalculateNormals();  // aiProcess_GenNormals
    std::string_view materialId;
    auto bound = materialBinding.find(prim->attr("material"));
    if (bound != materialBinding.end()) materialId = bound->second;
    else materialId = prim->attr("material");
    g->matIndex.assign(g->indices.size() / 3, importMaterialIndex(daeMaterial(d, materialId)));
    g->isOpaque = global::materials[g->matIndex.front()].alpha >= 1.F;
    out.push_back(std::move(g));
  }
}

void walkNode(const DaeDoc& d, const Xml& node, const glm::mat4& parent, bool dynamic, int depth,
              std::vector<std::shared_ptr<Geometry>>& out) {
  if (depth > 64) throw std::runtime_error("Failed to load scene: node hierarchy too deep, " + d.file);
  const glm::mat4 world = parent * nodeLocalTransform(node);
  node.each("instance_geometry", [&](const Xml& ig) {
    auto gi = d.geometries.find(stripHash(ig.attr("url")));
    if (gi == d.geometries.end()) return;
    std::map<std::string_view, std::string_view> binding;
    if (const Xml* bm = ig.child("bind_material"))
      if (const Xml* tc = bm->child("technique_common"))
        tc->each("instance_material", [&](const Xml& im) { binding[im.attr("symbol")] = stripHash(im.attr("target")); });
    emitGeometry(d, *gi->second, world, binding, dynamic, out);
  });
  if (node.child("instance_controller")) KF_WARN("skinned / morphed geometry is not imported: " + d.file);
  node.each("instance_node", [&](const Xml& in) {
    auto ni = d.nodes.find(stripHash(in.attr("url")));
    if (ni != d.nodes.end()) walkNode(d, *ni->second, world, dynamic, depth + 1, out);
  });
  node.each("node", [&](const Xml& k) { walkNode(d, k, world, dynamic, depth + 1, out); });
}

void indexById(const Xml& lib, std::string_view element, std::map<std::string_view, const Xml*>& out) {
  lib.each(element, [&](const Xml& e) {
    if (!e.attr("id").empty()) out[e.attr("id")] = &e;
  });
}
void indexNodes(const Xml& parent, std::map<std::string_view, const Xml*>& out) {
  parent.each("node", [&](const Xml& n) {
    if (!n.attr("id").empty()) out[n.attr("id")] = &n;
    indexNodes(n, out);
  });
}

std::string dirOfPath(const std::string& p) {
  const size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string() : p.substr(0, k + 1);
}

bool readWholeFile(const std::string& path, std::string& out) {
  std::ifstream in(path, std::ios::binary);
  if (!in.good()) return false;
  in.seekg(0, std::ios::end);
  const std::streamoff n = in.tellg();
  in.seekg(0);
  out.resize(size_t(n));
  in.read(out.data(), n);
  return bool(in);
}
}  // namespace

std::vector<std::shared_ptr<Geometry>> loadColladaScene(const std::string& path, bool dynamic) {
  std::string buf;
  if (!readWholeFile(path, buf)) throw std::runtime_error("Failed to load scene: cannot open " + path);
  XmlParser xp{buf.data(), buf.data() + buf.size(), path};
  xp.skipMisc();
  DaeDoc d;
  d.file = path;
  d.dir = dirOfPath(path);
  d.root = xp.element();
  if (d.root->name != "COLLADA") throw std::runtime_error("Failed to load scene: not a COLLADA document, " + path);
  d.root->each("library_geometries", [&](const Xml& l) { indexById(l, "geometry", d.geometries); });
  d.root->each("library_materials", [&](const Xml& l) { indexById(l, "material", d.materials); });
  d.root->each("library_effects", [&](const Xml& l) { indexById(l, "effect", d.effects); });
  d.root->each("library_images", [&](const Xml& l) { indexById(l, "image", d.images); });
  d.root->each("library_nodes", [&](const Xml& l) { indexNodes(l, d.nodes); });
  d.root->each("library_visual_scenes", [&](const Xml& l) {
    indexById(l, "visual_scene", d.scenes);
    l.each("visual_scene", [&](const Xml& vs) { indexNodes(vs, d.nodes); });
  });
  const Xml* vs = nullptr;
  if (const Xml* sc = d.root->child("scene"))
    if (const Xml* ivs = sc->child("instance_visual_scene")) {
      auto it = d.scenes.find(stripHash(ivs->attr("url")));
      if (it != d.scenes.end()) vs = it->second;
    }
  if (!vs && !d.scenes.empty()) vs = d.scenes.begin()->second;
  std::vector<std::shared_ptr<Geometry>> out;
  if (vs) vs->each("node", [&](const Xml& n) { walkNode(d, n, glm::mat4(1.0f), dynamic, 0, out); });
  if (out.empty()) throw std::runtime_error("Failed to load scene: no triangle meshes in " + path);
  return out;
}

// STL, binary or ASCII: one Geometry, per-corner vertices, facet normals as given (generated when a
// facet normal is zero, as Assimp's STL loader does), default material (Assimp: grey 0.6 diffuse).
std::vector<std::shared_ptr<Geometry>> loadStlScene(const std::string& path, bool dynamic) {
  std::string buf;
  if (!readWholeFile(path, buf)) throw std::runtime_error("Failed to load scene: cannot open " + path);
  auto g = std::make_shared<Geometry>();
  g->path = path;
  g->dynamic = dynamic;
  g->initialized = false;
  bool missingNormal = false;
  auto facet = [&](const glm::vec3& n, const glm::vec3 p[3]) {
    if (n == glm::vec3(0.0f)) missingNormal = true;
    for (int k = 0; k < 3; k++) {
      Vertex v;
      v.pos = p[k];
      v.normal = n;
      g->indices.push_back(uint32_t(g->vertices.size()));
      g->vertices.push_back(v);
    }
  };
  bool binary = false;
  if (buf.size() >= 84) {
    uint32_t n = 0;
    std::memcpy(&n, buf.data() + 80, 4);
    binary = buf.size() == 84 + size_t(n) * 50;
    if (binary)
      for (uint32_t t = 0; t < n; t++) {
        float f[12];
        std::memcpy(f, buf.data() + 84 + size_t(t) * 50, 48);
        const glm::vec3 p[3] = {{f[3], f[4], f[5]}, {f[6], f[7], f[8]}, {f[9], f[10], f[11]}};
        facet(glm::vec3(f[0], f[1], f[2]), p);
      }
  }
  if (!binary) {
    const char* p = buf.c_str();
    glm::vec3 n(0.0f), v[3];
    int nv = 0;
    while ((p = std::strpbrk(p, "fv")) != nullptr) {
      if (!std::strncmp(p, "facet normal", 12)) {
        char* q = nullptr;
        p += 12;
        n.x = std::strtof(p, &q); p = q;
        n.y = std::strtof(p, &q); p = q;
        n.z = std::strtof(p, &q); p = q;
        nv = 0;
      } else if (!std::strncmp(p, "vertex", 6)) {
        char* q = nullptr;
        p += 6;
        glm::vec3 w;
        w.x = std::strtof(p, &q); p = q;
        w.y = std::strtof(p, &q); p = q;
        w.z = std::strtof(p, &q); p = q;
        if (nv < 3) v[nv] = w;
        if (++nv == 3) facet(n, v);
      } else {
        p++;
      }
    }
  }
  if (g->indices.empty()) throw std::runtime_error("Failed to load scene: no triangles in " + path);
  if (missingNormal) g->recalculateNormals();
  NiceMaterial m;
  m.diffuseColor = glm::vec3(0.6f);
  m.emission = glm::vec3(0.0f);
  m.emissionStrength = 1.f;
  m.roughness = 0.f;  // no shininess key and not an .obj: stays 0 (geometry.cpp:108-131)
  g->matIndex.assign(g->indices.size() / 3, importMaterialIndex(m));
  g->isOpaque = true;
  return {g};
}
}  // namespace kuafu
