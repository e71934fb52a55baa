/*
 * kuafu_c.h -- flat C view of the C++ host facade (libkuafu.so), for tooling that cannot include
 * kuafu.hpp: the Python tests and bench.py use it to (1) build the BASELINE scenes through the real
 * facade code path, (2) drive Kuafu::run()/downloadLatestFrame(), and (3) read back the packed wire
 * buffers the facade hands to the kf_rt.h C ABI, so that the CPU oracle receives identical bits.
 * It adds no rendering logic of its own.
 */
#ifndef KUAFU_C_H
#define KUAFU_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define KFC_API __declspec(dllexport)
#else
#define KFC_API __attribute__((visibility("default")))
#endif

typedef struct KfcRenderer KfcRenderer;

/* wire buffer kinds for kfcWireBuffer */
enum {
  KFC_WIRE_MATERIALS = 0,   /* KfrtMaterial[n] */
  KFC_WIRE_INSTANCES = 1,   /* KfrtInstance[n] */
  KFC_WIRE_DIRECTIONAL = 2, /* KfrtDirectionalLight */
  KFC_WIRE_POINTS = 3,      /* KfrtPointLights */
  KFC_WIRE_ACTIVES = 4,     /* KfrtActiveLights */
  KFC_WIRE_CAMERA = 5,      /* KfrtCamera of recipe camera `index` */
  KFC_WIRE_PUSH = 6,        /* KfrtPushConstants the next frame would use */
  KFC_WIRE_TEXTURE = 7,     /* RGBA8 texels of texture `index` */
  KFC_WIRE_ENV_FACE = 8     /* RGBA8 texels of cube face `index` */
};

/* device < 0: host-only renderer (scene building and packing work, rendering fails loudly). */
KFC_API KfcRenderer* kfcCreate(int device, int accumulateFrames);
KFC_API void kfcDestroy(KfcRenderer* r);
KFC_API const char* kfcLastError(void);

/* Scene recipes of kuafu_b200/host/include/scenes.hpp; 0 keeps a config's own value. */
KFC_API int kfcLoadScene(KfcRenderer* r, const char* name, int width, int height, int spp, int depth, int scale);
KFC_API int kfcAnimate(KfcRenderer* r, int frame);
KFC_API int kfcNumCameras(KfcRenderer* r);
KFC_API int kfcSetCamera(KfcRenderer* r, int camera);

/* Kuafu::run() on the current camera / on all recipe cameras in one launch. */
KFC_API int kfcRun(KfcRenderer* r);
KFC_API int kfcRunAll(KfcRenderer* r);
/* Camera-batch split across processes: Kuafu::cameraShard() and Kuafu::run() on the recipe
 * cameras [begin, end) in one launch. */
KFC_API int kfcCameraShard(int nCameras, int rank, int world, int* begin, int* end);
KFC_API int kfcRunRange(KfcRenderer* r, int begin, int end);
/* Kuafu::cameraShardIndices() + Kuafu::run() on those recipe cameras in one launch; the indices rendered are
 * returned in `indices` (capacity entries), their number in *count.  After it, kfcDownloadFrame(camera) works
 * for the cameras rendered. */
KFC_API int kfcRunShard(KfcRenderer* r, int rank, int world, int interleaved, int* indices, int capacity, int* count);
/* Scene::setEnvironmentMap(path) (a .ktx cube map) on the current scene. */
KFC_API int kfcSetEnvironmentMap(KfcRenderer* r, const char* path);
/* The facade's texture / cube-map file readers (image_io.hpp), for the reader tests: dimensions always,
 * texels when dst holds at least width * height * 4 (x 6 for a cube) bytes. */
KFC_API int kfcReadTexture(const char* path, uint32_t* width, uint32_t* height, uint8_t* dst, size_t capacity);
KFC_API int kfcReadKtxCube(const char* path, uint32_t* size, uint8_t* dst, size_t capacity);
/* spp sharding across processes: trace samples [begin,end) only; with deferResolve the caller sums
 * the KFRT_AUX_SUM32F device buffers of all ranks and then calls kfcResolve(). begin == end == 0
 * restores whole frames. */
KFC_API int kfcSetSampleShard(KfcRenderer* r, uint32_t begin, uint32_t end, int deferResolve);
KFC_API int kfcResolve(KfcRenderer* r);
/* Camera::downloadLatestFrameInto (the frame straight into dst) / Kuafu::downloadLatestFrame (the reference's
 * by-value signature: a std::vector, then copied into dst). */
KFC_API int kfcDownloadFrame(KfcRenderer* r, int camera, uint8_t* dst, size_t nbytes);
KFC_API int kfcDownloadFrameByValue(KfcRenderer* r, int camera, uint8_t* dst, size_t nbytes);
/* kind: KFRT_AUX_* of kf_rt.h */
KFC_API int kfcDownloadAux(KfcRenderer* r, int camera, int kind, void* dst, size_t nbytes);
KFC_API uint32_t kfcClockBase(KfcRenderer* r);
KFC_API int kfcSetClockBase(KfcRenderer* r, uint32_t clockBase);
KFC_API int kfcFrameCount(void);
KFC_API void* kfcDeviceContext(KfcRenderer* r); /* the KfrtContext* under the facade */

/* Packs the current scene into wire format on the host (no device work). */
KFC_API int kfcPack(KfcRenderer* r);
/* out[0..7] = geometries, materials, textures, instances, cameras, envSize, width, height */
KFC_API int kfcWireCounts(KfcRenderer* r, uint32_t out[8]);
KFC_API int kfcWireGeometry(KfcRenderer* r, uint32_t index, const void** vertices, uint32_t* nVertices,
                            const uint32_t** indices, uint32_t* nIndices, const uint32_t** matIndex,
                            uint32_t* nMatIndex, int* opaque, int* hideRender);
KFC_API int kfcWireBuffer(KfcRenderer* r, int kind, uint32_t index, const void** ptr, size_t* nbytes);
KFC_API int kfcTextureDims(KfcRenderer* r, uint32_t index, uint32_t* width, uint32_t* height);

#ifdef __cplusplus
}
#endif
#endif /* KUAFU_C_H */
