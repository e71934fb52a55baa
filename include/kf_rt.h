/*
 * kf_rt.h -- C ABI of the B200-native path-tracing core that sits under the kuafu.hpp host API.
 *
 * This header is the drop-in boundary.  Everything the reference does through
 * Vulkan-RT on this path -- `RayTracer::createBottomLevelAS/buildTlas/updateTlas/trace`
 * (reference src/core/rt/rt.cpp:142-161,372-495,637-666), `Scene::upload*`
 * (reference src/core/scene.cpp:231-292,311-467) and `Camera::downloadLatestFrame`
 * (reference src/core/camera.cpp:188-207) -- is reached through the entry points below instead.
 * It lives beside the reference's `cuda_dl` loader pattern (reference include/cuda_dl.hpp:10-16):
 * plain C, opaque handle, integer status codes, no C++/torch types in any signature.
 *
 * Threading: one KfrtContext is used from one host thread at a time (the reference host is
 * single-threaded).  All device work of a context is issued on one CUDA stream
 * (internal by default, replaceable with kfrtSetStream).  No exception crosses this ABI; the C++
 * facade converts non-zero codes into std::runtime_error like the reference's KF_CRITICAL
 * (reference include/stdafx.hpp:61-64).
 *
 * There is no CPU fallback: every compute entry point fails with KFRT_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef KF_RT_H
#define KF_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define KFRT_API __declspec(dllexport)
#else
#define KFRT_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------------------------------
 * Wire structs.  Bit-for-bit the buffers the reference uploads to its shaders.
 * ---------------------------------------------------------------------------------------------- */

/* reference include/core/context/vertex.hpp:28-33 (48 B; shader view PathTrace.rchit:52-64) */
typedef struct KfrtVertex {
  float pos[3];
  float normal[3];
  float color[3];
  float texCoord[2];
  float padding0;
} KfrtVertex;

/* reference include/core/geometry.hpp:70-87 (80 B; shader view base/Geometry.glsl:1-19) */
typedef struct KfrtMaterial {
  float diffuse[4];  /* rgb + padding */
  float emission[4]; /* rgb + emission strength */
  float alpha;
  float metallic;
  float specular;
  float roughness;
  float ior;
  float transmission;
  int32_t diffuseTexIdx;
  int32_t metallicTexIdx;
  int32_t roughnessTexIdx;
  int32_t transmissionTexIdx;
  int32_t padding0;
  int32_t padding1;
} KfrtMaterial;

/* reference include/core/geometry.hpp:91-99 (80 B); transform is a column-major glm::mat4 */
typedef struct KfrtInstance {
  float transform[16];
  uint32_t geometryIndex;
  uint32_t padding0;
  uint32_t padding1;
  uint32_t padding2;
} KfrtInstance;

/* reference include/core/camera.hpp:181-192 (320 B; shader view base/Camera.glsl:1-13).
 * position.w = aperture, front.w = focus distance (reference src/core/scene.cpp:246-247). */
typedef struct KfrtCamera {
  float view[16];
  float projection[16];
  float viewInverse[16];
  float projectionInverse[16];
  float position[4];
  float front[4];
  float padding1[4];
  float padding2[4];
} KfrtCamera;

#define KFRT_MAX_POINT_LIGHTS 32 /* reference include/core/context/global.hpp:38 */
#define KFRT_MAX_ACTIVE_LIGHTS 8 /* reference include/core/context/global.hpp:39 */

/* reference include/core/light.hpp:41-44 (32 B) */
typedef struct KfrtDirectionalLight {
  float direction[4]; /* normalized direction + softness */
  float rgbs[4];      /* rgb + strength */
} KfrtDirectionalLight;

/* reference include/core/light.hpp:46-49 (1024 B) */
typedef struct KfrtPointLights {
  float posr[KFRT_MAX_POINT_LIGHTS][4]; /* position + radius */
  float rgbs[KFRT_MAX_POINT_LIGHTS][4]; /* rgb + strength (strength > 0 == in use) */
} KfrtPointLights;

/* reference include/core/light.hpp:51-58 (1536 B) */
typedef struct KfrtActiveLights {
  float viewMat[KFRT_MAX_ACTIVE_LIGHTS][16];
  float projMat[KFRT_MAX_ACTIVE_LIGHTS][16];
  float front[KFRT_MAX_ACTIVE_LIGHTS][4];    /* front + in-use flag (w > 0) */
  float rgbs[KFRT_MAX_ACTIVE_LIGHTS][4];     /* rgb + strength */
  float position[KFRT_MAX_ACTIVE_LIGHTS][4]; /* position */
  float sftp[KFRT_MAX_ACTIVE_LIGHTS][4];     /* softness, fov, texture id, padding */
} KfrtActiveLights;

/* reference include/core/rt/rt.hpp:31-42 (48 B; shader view base/PushConstants.glsl:1-16) */
typedef struct KfrtPushConstants {
  float clearColor[4];
  int32_t frameCount;
  uint32_t sampleRatePerPixel;
  uint32_t maxPathDepth;
  uint32_t useEnvironmentMap;
  uint32_t russianRoulette;
  uint32_t russianRouletteMinBounces;
  uint32_t nextEventEstimation;           /* pushed but never read by the reference shaders */
  uint32_t nextEventEstimationMinBounces; /* pushed but never read by the reference shaders */
} KfrtPushConstants;

/* Work counters of the last kfrtRender call (a "ray" is one traceRayEXT
 * equivalent, reference PathTrace.rgen:97 and PathTrace.rchit:186). */
typedef struct KfrtCounters {
  uint64_t paths;          /* camera paths started (pixels x samples x cameras) */
  uint64_t extensionRays;  /* closest-hit rays traced */
  uint64_t shadowRays;     /* occlusion rays traced */
  uint64_t extensionHits;  /* extension rays that hit geometry */
  uint64_t nodeVisits;     /* wide BVH nodes fetched (only when detail counters are on) */
  uint64_t triangleTests;  /* triangles fetched and tested (detail) */
  uint64_t instanceVisits; /* TLAS leaf entries (ray transformed into a BLAS) (detail) */
  uint64_t textureFetches; /* bilinear texture lookups (detail) */
  uint64_t kernelLaunches; /* CUDA kernels launched by the last kfrtRender/kfrtResolve */
  uint64_t shadowNodeVisits;     /* the occlusion-ray share of nodeVisits (detail) */
  uint64_t shadowTriangleTests;  /* the occlusion-ray share of triangleTests (detail) */
  uint64_t shadowInstanceVisits; /* the occlusion-ray share of instanceVisits (detail) */
  uint64_t tlasNodeVisits;       /* the top-level share of nodeVisits, both ray kinds (detail) */
  uint64_t instanceEntries;      /* instanceVisits whose world-box re-test passed: BLAS traversals (detail) */
  uint64_t shadowRaysSkipped;    /* light samples whose occlusion ray could not change colour or RNG state and was not traced (not in shadowRays) */
  uint64_t reserved[1];
} KfrtCounters;

/* BVH statistics for the roofline accounting in DESIGN.md. */
typedef struct KfrtBvhStats {
  uint32_t blasCount;
  uint32_t instanceCount;
  uint64_t triangleCount;      /* sum over BLAS (un-instanced) */
  uint64_t instancedTriangles; /* sum over instances */
  uint64_t blasNodeCount;      /* wide nodes, all BLAS */
  uint64_t tlasNodeCount;
  uint32_t nodeBytes;          /* bytes per wide node */
  uint32_t triangleBytes;      /* bytes per stored triangle */
  uint32_t instanceBytes;      /* bytes per instance record read on TLAS leaf entry */
  uint32_t tlasRebuilds;       /* kfrtRefitTlas calls that rebuilt the top level instead (quality watch) */
  uint64_t subtreeNodeCount;   /* wide nodes of top level + world-space instance subtrees (kfrtSetInstanceSubtrees); 0 while the two-level structure is walked (as of the last kfrtRender) */
  uint64_t subtreeTriangles;   /* triangle records in them (== instancedTriangles of the visible instances) */
  uint32_t subtreeDepth;       /* levels: top level + deepest subtree */
  uint32_t subtreeBuilds;      /* times they have been built in this context */
} KfrtBvhStats;

typedef struct KfrtContext KfrtContext;

enum {
  KFRT_OK = 0,
  KFRT_ERR_INVALID = 1,   /* bad argument / call order */
  KFRT_ERR_CUDA = 2,      /* CUDA runtime failure (incl. no usable device) */
  KFRT_ERR_LIMIT = 3,     /* a Config limit (geometry/instances/textures/materials) exceeded */
  KFRT_ERR_NOT_BUILT = 4, /* render without BLAS/TLAS */
  KFRT_ERR_NCCL = 5
};

/* kfrtDownloadAux / kfrtGetDeviceBuffer kinds.  HIT_* and DEPTH/SEGMENTATION are the additive
 * outputs the reference lists as TODO (reference README.md:64-68); they describe sample 0,
 * depth 0 of the last kfrtRender whose sampleBegin was 0. */
enum {
  KFRT_AUX_RGBA32F = 0,      /* float4 running-mean image (reference PathTrace.rgen:153-163) */
  KFRT_AUX_ALBEDO32F = 1,    /* float4 first-hit albedo (PathTrace.rgen:165) */
  KFRT_AUX_NORMAL32F = 2,    /* float4 first-hit normal (PathTrace.rgen:166) */
  KFRT_AUX_HIT_IDS = 3,      /* int32x2 (instance index, primitive index); (-1,-1) on miss */
  KFRT_AUX_HIT_T = 4,        /* float  ray parameter t of the primary hit; 0 on miss */
  KFRT_AUX_DEPTH = 5,        /* float  t * dot(direction, camera front); 0 on miss */
  KFRT_AUX_SEGMENTATION = 6, /* int32  instance index; -1 on miss */
  KFRT_AUX_SUM32F = 7,       /* float4 un-normalised sample sum of the last kfrtRender */
  KFRT_AUX_BGRA8 = 8         /* uint8x4 encoded frame (same bytes as kfrtDownloadBGRA8) */
};

/* ------------------------------------------------------------------------------------------------
 * Lifetime
 * ---------------------------------------------------------------------------------------------- */
KFRT_API int kfrtCreate(int deviceOrdinal, KfrtContext** out);
KFRT_API int kfrtDestroy(KfrtContext* ctx);
/* Last error text of `ctx` (or of the failed kfrtCreate when ctx == NULL). Never NULL. */
KFRT_API const char* kfrtLastError(const KfrtContext* ctx);
KFRT_API const char* kfrtVersion(void);
/* Issue all further device work on `cudaStream` (a cudaStream_t); NULL restores the internal one. */
KFRT_API int kfrtSetStream(KfrtContext* ctx, void* cudaStream);
KFRT_API int kfrtSynchronize(KfrtContext* ctx);

/* Config limits (reference include/core/config.hpp:157-163; exceeded -> KFRT_ERR_LIMIT, the
 * facade turns that into the std::runtime_error of reference src/core/scene.cpp:60-64,115-119). */
KFRT_API int kfrtSetLimits(KfrtContext* ctx, uint32_t maxGeometry, uint32_t maxInstances,
                           uint32_t maxTextures, uint32_t maxMaterials);

/* ------------------------------------------------------------------------------------------------
 * Scene upload  (replaces Scene::uploadGeometries, reference src/core/scene.cpp:311-442)
 * ---------------------------------------------------------------------------------------------- */
/* nIndices = 3 * primitive count.  matIndex holds >= primitive count entries (the reference
 * sizes it to nIndices, src/core/geometry.cpp:314).  opaque == 0 runs the stochastic any-hit test
 * on extension rays (reference src/core/rt/rt.cpp:101-102); hideRender != 0 makes the geometry
 * untouchable by rays (the reference swaps in a dummy BLAS, rt.cpp:153-157). */
KFRT_API int kfrtUploadGeometry(KfrtContext* ctx, uint32_t geometryIndex, const KfrtVertex* vertices,
                                uint32_t nVertices, const uint32_t* indices, uint32_t nIndices,
                                const uint32_t* matIndex, uint32_t nMatIndex, int opaque,
                                int hideRender);
KFRT_API int kfrtClearGeometries(KfrtContext* ctx);
KFRT_API int kfrtUploadMaterials(KfrtContext* ctx, const KfrtMaterial* materials, uint32_t n);
/* RGBA8 texels, row 0 first; sampled bilinear / repeat / mip 0 with sRGB decode of rgb
 * (reference vkCore.hpp:580-637,1761-1795). */
KFRT_API int kfrtUploadTexture(KfrtContext* ctx, uint32_t textureIndex, const uint8_t* rgba8,
                               uint32_t width, uint32_t height);
/* Six square RGBA8 faces in Vulkan cube order (+X,-X,+Y,-Y,+Z,-Z), sRGB decoded on sample
 * (reference src/core/scene.cpp:294-309, PathTrace.rmiss:14-30). */
KFRT_API int kfrtSetEnvironmentCube(KfrtContext* ctx, const uint8_t* const faces[6], uint32_t size);
KFRT_API int kfrtClearEnvironment(KfrtContext* ctx);
KFRT_API int kfrtSetLights(KfrtContext* ctx, const KfrtDirectionalLight* directional,
                           const KfrtPointLights* points, const KfrtActiveLights* actives);

/* ------------------------------------------------------------------------------------------------
 * Acceleration structures  (replaces reference src/core/rt/rt.cpp:142-370 and :372-495)
 * ---------------------------------------------------------------------------------------------- */
/* (Re)build the bottom-level structure of every geometry uploaded since the last call. */
KFRT_API int kfrtBuildBlas(KfrtContext* ctx);
KFRT_API int kfrtSetInstances(KfrtContext* ctx, const KfrtInstance* instances, uint32_t n);
/* Full top-level build over the current instances. */
KFRT_API int kfrtBuildTlas(KfrtContext* ctx);
/* Per-frame path: new column-major 4x4 transforms for the n current instances; topology of the
 * top-level tree is kept, boxes and inverse transforms are recomputed on the device.  The call watches
 * the quality of the refitted tree (area sum of its nodes, read back one call late) and runs the full
 * build instead when it has degraded past 1.1x its value at build time. */
KFRT_API int kfrtRefitTlas(KfrtContext* ctx, const float* transforms, uint32_t n);
/* World-space instance subtrees (no counterpart in the reference, which leaves instancing to the driver
 * behind vkCmdBuildAccelerationStructuresKHR, reference src/core/rt/rt.cpp:116-140): when the instanced
 * triangles of the scene number at most maxTriangles (and its visible instances at most 1 024), every
 * instance can get its own subtree in world space under the top level, and kfrtRender then walks that
 * single-space hierarchy instead of the two-level structure -- same hit buffers bit for bit, see kf_wsi.cuh.
 * mode 0: never.  mode 1 (default): scenes that have not been refitted since kfrtBuildTlas; the first
 * kfrtRender after the build times one sample of its frame with either structure and keeps the faster
 * (KfrtBvhStats.subtreeNodeCount != 0 afterwards: the subtrees won).  mode 2: whenever the scene fits,
 * rebuilt by the first kfrtRender after every kfrtBuildTlas / kfrtRefitTlas.  maxTriangles 0 keeps the
 * current budget (default 4 Mi).  Environment overrides at kfrtCreate: KFRT_INSTANCE_SUBTREES,
 * KFRT_INSTANCE_SUBTREES_MAX_TRIS. */
KFRT_API int kfrtSetInstanceSubtrees(KfrtContext* ctx, int mode, uint64_t maxTriangles);
/* Light samples whose occlusion ray cannot change the image -- a contribution of exactly zero that draws no
 * random number (PathTrace.rchit:107-108: a transmissive surface seen from inside), or any sample of a path
 * whose weight the BSDF has just taken to zero (rgen:119 ends it before its next draw) -- are answered without
 * a ray and counted in KfrtCounters.shadowRaysSkipped (on: 1, the default).  0 traces every light sample the
 * reference traces: same first-hit buffers bit for bit, radiance within the rounding of the shading arithmetic
 * (the switch selects another instantiation of the shade kernel), shadowRays larger by shadowRaysSkipped. */
KFRT_API int kfrtSetLightSampleCulling(KfrtContext* ctx, int on);
/* kfrtBuildBlas finds out which geometries are convex (every vertex on or behind the plane of every triangle;
 * closed convex meshes with outward winding, planar meshes, convex caps).  A bounce or occlusion ray that leaves
 * such a geometry to the front side of the triangle it starts on, by more than a grazing margin, cannot meet it
 * again, and the traversal stages do not enter the instance it starts on (on: 1, the default; environment
 * override at kfrtCreate: KFRT_SKIP_OWN_INSTANCE).  0: every ray walks every instance its path meets, as in
 * the reference; the buffers agree except where rounding makes a ray along its own convex surface hit it. */
KFRT_API int kfrtSetOwnInstanceSkip(KfrtContext* ctx, int on);
KFRT_API int kfrtGetBvhStats(KfrtContext* ctx, KfrtBvhStats* out);

/* ------------------------------------------------------------------------------------------------
 * Render  (replaces RayTracer::trace -> traceRaysKHR(w,h,1), reference src/core/rt/rt.cpp:637-666)
 * ---------------------------------------------------------------------------------------------- */
/* Trace samples [sampleBegin, sampleEnd) of every pixel of nCameras cameras of width x height and
 * leave their un-normalised sum in the SUM32F buffer.  pc->sampleRatePerPixel is the TOTAL sample
 * count of the frame (the divisor used by kfrtResolve); a single-GPU frame passes
 * sampleBegin = 0, sampleEnd = sampleRatePerPixel.
 *
 * clockBase is the deterministic surrogate of the reference's clockARB() seeds
 * (PathTrace.rgen:23,32; declared deviation D1): the pixel-jitter stream is seeded with
 * tea(pixel, clockBase) and sample i with tea(pixel, clockBase + 1 + i). */
KFRT_API int kfrtRender(KfrtContext* ctx, const KfrtCamera* cameras, uint32_t nCameras,
                        uint32_t width, uint32_t height, const KfrtPushConstants* pc,
                        uint32_t sampleBegin, uint32_t sampleEnd, uint32_t clockBase);
/* finalColor = SUM / sampleRatePerPixel; frameCount <= 0 ? store : mix(old, new, 1/(frameCount+1))
 * (PathTrace.rgen:143-163), then encode clamp -> linear-to-sRGB -> BGRA8, alpha 255
 * (PostProcessing.frag:11-18 + B8G8R8A8Srgb attachment, include/core/config.hpp:144). */
KFRT_API int kfrtResolve(KfrtContext* ctx);
/* spp-sharded frames: sum the SUM32F buffers of all ranks of `ncclComm` (an ncclComm_t) onto
 * `root` (root < 0: all-reduce).  Call between kfrtRender and kfrtResolve. */
KFRT_API int kfrtReduceNccl(KfrtContext* ctx, void* ncclComm, int root);

/* == Camera::downloadLatestFrame (reference src/core/camera.cpp:188-207): width*height*4 bytes,
 * BGRA, sRGB-encoded, alpha 255; synchronises the context's stream. */
KFRT_API int kfrtDownloadBGRA8(KfrtContext* ctx, uint32_t camera, uint8_t* dst, size_t nbytes);
KFRT_API int kfrtDownloadAux(KfrtContext* ctx, uint32_t camera, int kind, void* dst, size_t nbytes);
/* The same frame without the copy: a pointer into the context's pinned staging buffer, which kfrtResolve
 * starts filling asynchronously the moment the frame is encoded.  Waits for that copy only; the bytes
 * stay valid until the next kfrtResolve / kfrtRender of a different size.  (A host that returns the frame
 * by value, like Camera::downloadLatestFrame, builds its vector from this in one pass.) */
KFRT_API int kfrtMapBGRA8(KfrtContext* ctx, uint32_t camera, const uint8_t** bytes, size_t* nbytes);
/* Device address and size (all cameras, camera-major) of an output buffer, for host plumbing that
 * wants to run a collective or a copy on it without a host round trip. */
KFRT_API int kfrtGetDeviceBuffer(KfrtContext* ctx, int kind, void** devicePtr, size_t* nbytes);

/* Per-stage device time of the last kfrtRender, for the roofline accounting of single kernels.
 * With timers on, kfrtRender records a CUDA event on the context's stream before every stage launch;
 * kfrtGetStageTimes waits for the last one and sums the event-to-event durations per stage. */
enum {
  KFRT_STAGE_RAYGEN = 0,          /* k_wf_raygen */
  KFRT_STAGE_TRACE_CLOSEST = 1,   /* k_wf_trace<closest hit>: extension rays */
  KFRT_STAGE_SHADE = 2,           /* k_wf_shade */
  KFRT_STAGE_TRACE_OCCLUSION = 3, /* k_wf_trace<first hit>: shadow rays */
  KFRT_STAGE_SHADOW_RESOLVE = 4,  /* k_wf_shadow_resolve */
  KFRT_STAGE_FINISH = 5,          /* k_wf_finish */
  KFRT_STAGE_OTHER = 6,
  KFRT_STAGE_END = 7
};
typedef struct KfrtStageTimes {
  double ms[8];          /* total device milliseconds per stage */
  uint64_t launches[8];  /* kernel launches per stage */
} KfrtStageTimes;
KFRT_API int kfrtSetStageTimers(KfrtContext* ctx, int on);
KFRT_API int kfrtGetStageTimes(KfrtContext* ctx, KfrtStageTimes* out);

/* detail != 0 additionally counts node/triangle/instance/texture fetches (slower kernels). */
KFRT_API int kfrtSetDetailCounters(KfrtContext* ctx, int detail);
KFRT_API int kfrtGetCounters(KfrtContext* ctx, KfrtCounters* out);

#ifdef __cplusplus
}
#endif
#endif /* KF_RT_H */
