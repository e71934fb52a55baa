// Flat C view of the facade (include/kuafu_c.h).  Thin: every call forwards to the kuafu.hpp API.
#include "kuafu_c.h"
#include "image_io.hpp"

#include "kf_rt.h"
#include "scenes.hpp"

using namespace kuafu;

struct KfcRenderer {
  std::unique_ptr<Kuafu> renderer;
  std::vector<Camera*> cameras;
  std::vector<CameraUBO> cameraUbos;
  KfrtPushConstants push{};
};

static thread_local std::string g_error;

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

extern "C" {

const char* kfcLastError(void) { return g_error.c_str(); }

KfcRenderer* kfcCreate(int device, int accumulateFrames) {
  KfcRenderer* out = nullptr;
  const int rc = guarded([&] {
    auto config = std::make_shared<Config>();
    config->setPresent(false);
    config->setAccumulatingFrames(accumulateFrames != 0);
    config->setGeometryLimit(1024);
    config->setGeometryInstanceLimit(8192);
    config->setTextureLimit(256);
    config->setMaterialLimit(4096);
    config->setHostOnly(device < 0);
    config->setDeviceOrdinal(device < 0 ? 0 : device);
    auto r = std::make_unique<KfcRenderer>();
    r->renderer = std::make_unique<Kuafu>(config);
    out = r.release();
  });
  return rc ? nullptr : out;
}

void kfcDestroy(KfcRenderer* r) { delete r; }

int kfcLoadScene(KfcRenderer* r, const char* name, int width, int height, int spp, int depth, int scale) {
  return guarded([&] {
    // every recipe starts from an empty material registry so that indices are reproducible
    global::materials.clear();
    global::materialIndex = 0;
    global::frameCount = -1;
    Scene* fresh = r->renderer->createScene();
    Scene* old = r->renderer->getScene();
    r->renderer->setScene(fresh);
    if (old) r->renderer->removeScene(old);
    scenes::Recipe rc;
    rc.name = name;
    rc.width = width;
    rc.height = height;
    rc.spp = spp;
    rc.depth = depth > 0 ? depth : -1;
    rc.scale = scale;
    r->cameras = scenes::load(*r->renderer, rc);
  });
}

int kfcAnimate(KfcRenderer* r, int frame) {
  return guarded([&] { scenes::animate(*r->renderer, frame); });
}

int kfcNumCameras(KfcRenderer* r) { return int(r->cameras.size()); }

int kfcSetCamera(KfcRenderer* r, int camera) {
  return guarded([&] { r->renderer->getScene()->setCamera(r->cameras.at(size_t(camera))); });
}

int kfcRun(KfcRenderer* r) {
  return guarded([&] { r->renderer->run(); });
}

int kfcRunAll(KfcRenderer* r) {
  return guarded([&] { r->renderer->run(r->cameras); });
}

int kfcCameraShard(int nCameras, int rank, int world, int* begin, int* end) {
  return guarded([&] {
    if (nCameras < 0) throw std::runtime_error("kfcCameraShard: negative camera count");
    const auto range = Kuafu::cameraShard(size_t(nCameras), rank, world);
    *begin = int(range.first);
    *end = int(range.second);
  });
}

int kfcRunRange(KfcRenderer* r, int begin, int end) {
  return guarded([&] {
    if (begin < 0 || end < begin || size_t(end) > r->cameras.size()) throw std::runtime_error("kfcRunRange: bad camera range");
    r->renderer->run(std::vector<Camera*>(r->cameras.begin() + begin, r->cameras.begin() + end));
  });
}

int kfcRunShard(KfcRenderer* r, int rank, int world, int interleaved, int* indices, int capacity, int* count) {
  return guarded([&] {
    const auto idx = Kuafu::cameraShardIndices(r->cameras.size(), rank, world, interleaved != 0);
    if (int(idx.size()) > capacity) throw std::runtime_error("kfcRunShard: index buffer too small");
    std::vector<Camera*> cams;
    for (size_t k = 0; k < idx.size(); k++) {
      cams.push_back(r->cameras.at(idx[k]));
      indices[k] = int(idx[k]);
    }
    *count = int(idx.size());
    if (!cams.empty()) r->renderer->run(cams);
  });
}

int kfcSetEnvironmentMap(KfcRenderer* r, const char* path) {
  return guarded([&] { r->renderer->getScene()->setEnvironmentMap(path); });
}

int kfcReadTexture(const char* path, uint32_t* width, uint32_t* height, uint8_t* dst, size_t capacity) {
  return guarded([&] {
    std::vector<uint8_t> rgba;
    if (!io::loadTextureRGBA8(path, *width, *height, rgba)) throw std::runtime_error(std::string("cannot read texture ") + path);
    if (dst && capacity >= rgba.size()) std::memcpy(dst, rgba.data(), rgba.size());
  });
}

int kfcReadKtxCube(const char* path, uint32_t* size, uint8_t* dst, size_t capacity) {
  return guarded([&] {
    std::vector<uint8_t> faces[6];
    if (!io::loadKtxCubeRGBA8(path, *size, faces)) throw std::runtime_error(std::string("cannot read cube map ") + path);
    const size_t fb = size_t(*size) * *size * 4;
    if (dst && capacity >= 6 * fb)
      for (int f = 0; f < 6; f++) std::memcpy(dst + f * fb, faces[f].data(), fb);
  });
}

int kfcSetSampleShard(KfcRenderer* r, uint32_t begin, uint32_t end, int deferResolve) {
  return guarded([&] {
    if (begin == 0 && end == 0) r->renderer->getContext().clearSampleShard();
    else r->renderer->getContext().setSampleShard(begin, end, deferResolve != 0);
  });
}

int kfcResolve(KfcRenderer* r) {
  return guarded([&] { r->renderer->getContext().resolve(); });
}

int kfcDownloadFrame(KfcRenderer* r, int camera, uint8_t* dst, size_t nbytes) {
  return guarded([&] {
    // the caller owns the destination: Camera::downloadLatestFrameInto (the by-value downloadLatestFrame of the
    // reference's signature is kfcDownloadFrameByValue)
    r->cameras.at(size_t(camera))->downloadLatestFrameInto(dst, nbytes);
  });
}

int kfcDownloadFrameByValue(KfcRenderer* r, int camera, uint8_t* dst, size_t nbytes) {
  return guarded([&] {
    const std::vector<uint8_t> f = r->renderer->downloadLatestFrame(r->cameras.at(size_t(camera)));
    if (f.size() != nbytes) throw std::runtime_error("kfcDownloadFrameByValue: destination size mismatch");
    std::memcpy(dst, f.data(), nbytes);
  });
}

int kfcDownloadAux(KfcRenderer* r, int camera, int kind, void* dst, size_t nbytes) {
  return guarded([&] {
    Camera* c = r->cameras.at(size_t(camera));
    Context& ctx = r->renderer->getContext();
    if (!c->mFrames.valid) throw std::runtime_error("kfcDownloadAux: camera has no frame");
    if (kfrtDownloadAux(ctx.getDevice(), c->mFrames.slot, kind, dst, nbytes) != KFRT_OK)
      throw std::runtime_error(std::string("kfrtDownloadAux: ") + kfrtLastError(ctx.getDevice()));
  });
}

uint32_t kfcClockBase(KfcRenderer* r) { return r->renderer->getContext().getClockBase(); }
int kfcSetClockBase(KfcRenderer* r, uint32_t c) {
  r->renderer->getContext().setClockBase(c);
  return 0;
}
int kfcFrameCount(void) { return global::frameCount; }
void* kfcDeviceContext(KfcRenderer* r) { return r->renderer->getContext().getDevice(); }

int kfcPack(KfcRenderer* r) {
  return guarded([&] {
    Context& ctx = r->renderer->getContext();
    ctx.pack();
    Scene* s = r->renderer->getScene();
    r->cameraUbos.clear();
    for (Camera* c : r->cameras) {
      c->updateViewMatrix();
      r->cameraUbos.push_back(s->packCamera(c));
    }
    r->push = ctx.makePushConstants(ctx.predictFrameCount());
  });
}

int kfcWireCounts(KfcRenderer* r, uint32_t out[8]) {
  return guarded([&] {
    Scene* s = r->renderer->getScene();
    const WireScene& w = s->wire();
    out[0] = uint32_t(s->getGeometries().size());
    out[1] = uint32_t(w.materials.size());
    out[2] = uint32_t(w.textures.size());
    out[3] = uint32_t(w.instances.size());
    out[4] = uint32_t(r->cameras.size());
    out[5] = r->push.useEnvironmentMap ? w.envSize : 0u;
    out[6] = r->cameras.empty() ? 0u : uint32_t(r->cameras[0]->getWidth());
    out[7] = r->cameras.empty() ? 0u : uint32_t(r->cameras[0]->getHeight());
  });
}

int kfcWireGeometry(KfcRenderer* r, uint32_t index, const void** vertices, uint32_t* nVertices,
                    const uint32_t** indices, uint32_t* nIndices, const uint32_t** matIndex, uint32_t* nMatIndex,
                    int* opaque, int* hideRender) {
  return guarded([&] {
    const auto& g = r->renderer->getScene()->getGeometries().at(index);
    *vertices = g->vertices.data();
    *nVertices = uint32_t(g->vertices.size());
    *indices = g->indices.data();
    *nIndices = uint32_t(g->indices.size());
    *matIndex = g->matIndex.data();
    *nMatIndex = uint32_t(g->matIndex.size());
    *opaque = g->isOpaque ? 1 : 0;
    *hideRender = g->hideRender ? 1 : 0;
  });
}

int kfcWireBuffer(KfcRenderer* r, int kind, uint32_t index, const void** ptr, size_t* nbytes) {
  return guarded([&] {
    const WireScene& w = r->renderer->getScene()->wire();
    switch (kind) {
      case KFC_WIRE_MATERIALS: *ptr = w.materials.data(); *nbytes = w.materials.size() * sizeof(NiceMaterialSSBO); break;
      case KFC_WIRE_INSTANCES: *ptr = w.instances.data(); *nbytes = w.instances.size() * sizeof(GeometryInstanceSSBO); break;
      case KFC_WIRE_DIRECTIONAL: *ptr = &w.directional; *nbytes = sizeof(w.directional); break;
      case KFC_WIRE_POINTS: *ptr = &w.points; *nbytes = sizeof(w.points); break;
      case KFC_WIRE_ACTIVES: *ptr = &w.actives; *nbytes = sizeof(w.actives); break;
      case KFC_WIRE_CAMERA: *ptr = &r->cameraUbos.at(index); *nbytes = sizeof(CameraUBO); break;
      case KFC_WIRE_PUSH: *ptr = &r->push; *nbytes = sizeof(r->push); break;
      case KFC_WIRE_TEXTURE: *ptr = w.textures.at(index).rgba.data(); *nbytes = w.textures.at(index).rgba.size(); break;
      case KFC_WIRE_ENV_FACE:
        if (index >= 6) throw std::runtime_error("cube face index out of range");
        *ptr = w.envFaces[index].data();
        *nbytes = w.envFaces[index].size();
        break;
      default: throw std::runtime_error("unknown wire buffer kind");
    }
  });
}

int kfcTextureDims(KfcRenderer* r, uint32_t index, uint32_t* width, uint32_t* height) {
  return guarded([&] {
    const auto& t = r->renderer->getScene()->wire().textures.at(index);
    *width = t.width;
    *height = t.height;
  });
}

}  // extern "C"
