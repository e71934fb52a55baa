// kf_rt.cu -- implementation of the kf_rt.h C ABI: context, uploads, acceleration-structure build
// sequencing, the path-tracing kernels and the output stage.  sm_100a only; no CPU fallback.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "kf_blas_batch.cuh"
#include "kf_bvh_build.cuh"
#include "kf_common.cuh"
#include "kf_shade.cuh"
#include "kf_traverse.cuh"
#include "kf_wavefront.cuh"
#include "kf_wsi.cuh"

using namespace kf;

static_assert(sizeof(KfrtVertex) == 48, "Vertex wire layout");
static_assert(sizeof(KfrtMaterial) == 80, "NiceMaterialSSBO wire layout");
static_assert(sizeof(KfrtInstance) == 80, "GeometryInstanceSSBO wire layout");
static_assert(sizeof(KfrtCamera) == 320, "CameraUBO wire layout");
static_assert(sizeof(KfrtDirectionalLight) == 32, "DirectionalLightUBO wire layout");
static_assert(sizeof(KfrtPointLights) == 1024, "PointLightsUBO wire layout");
static_assert(sizeof(KfrtActiveLights) == 1536, "ActiveLightsUBO wire layout");
static_assert(sizeof(KfrtPushConstants) == 48, "RtPushConstants wire layout");

// ------------------------------------------------------------------------------------------------
// small host utilities
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_createError = "";

// Grow-only device buffer.  Storage comes from the device's stream-ordered pool (kept by
// kfrtCreate's release threshold), so that a context created after another one in the same process
// re-uses its memory instead of going back to cudaMalloc (measured: 8 - 140 ms for the scratch of a
// 250 k-triangle build, against 4 ms for the build itself).  Growing or releasing a buffer
// synchronises the device, exactly as the cudaFree it replaces did.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    release();
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T), cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
    if (e == cudaSuccess) cap = n;
    else p = nullptr;
    return e;
  }
  void release() {
    if (p) {
      cudaDeviceSynchronize();
      cudaFreeAsync(p, cudaStreamPerThread);
    }
    p = nullptr;
    cap = 0;
  }
};

// Device storage shared by the bottom-level structures of one build batch (nodes, triangles and shading
// records of all its geometries in one allocation); returned to the stream-ordered pool when the last
// geometry that lives in it is rebuilt or freed.
struct BlasBlock {
  void* p = nullptr;
  cudaStream_t stream = nullptr;
  ~BlasBlock() {
    if (p) cudaFreeAsync(p, stream);
  }
};

struct GeomHost {
  bool present = false, dirty = false, opaque = true, hide = false;
  uint32_t nVerts = 0, nIdx = 0, nMat = 0;
  KfrtVertex* verts = nullptr;
  uint32_t* idx = nullptr;
  uint32_t* matIndex = nullptr;
  Node8* nodes = nullptr;       // root; the record in front of it (nodesAlloc) holds the bounding sphere
  Node8* nodesAlloc = nullptr;  // nodesAlloc, tris and shade point into `block`
  Tri48* tris = nullptr;
  ShadeTri* shade = nullptr;
  float4* faceNormal = nullptr;  // object-space e1 x e2 per primitive (in `block`)
  bool convex = false;           // k_batch_convex: rays leaving its surface outwards skip the instance they start on
  std::shared_ptr<BlasBlock> block;
  uint32_t nNodes = 0;
  uint32_t depth = 0;  // levels of its wide tree
  float box[6] = {0, 0, 0, 0, 0, 0};
  void freeAll() {
    cudaFree(verts); cudaFree(idx); cudaFree(matIndex);
    block.reset();
    verts = nullptr; idx = nullptr; matIndex = nullptr; nodes = nullptr; nodesAlloc = nullptr; tris = nullptr; shade = nullptr;
    faceNormal = nullptr; convex = false;
    present = false;
    nNodes = 0;
  }
};

struct TexHost {
  uchar4* texels = nullptr;
  uint32_t w = 0, h = 0;
};

// Scratch + retained state of one BVH build (the TLAS keeps it for refits).
struct BuildState {
  uint32_t n = 0;
  uint32_t nWide = 0;
  uint32_t depth = 1;  // levels of the wide tree (the traversal stack holds at most one entry per level)
  DevBuf<float> primBox;
  DevBuf<int> sceneBox;
  DevBuf<uint64_t> keysA, keysB;
  DevBuf<uint32_t> valsA, valsB, hist, flags, outPrim, counters, slotOfInst;
  DevBuf<int2> children, range;
  DevBuf<int> parent, wideBinary, wideMembers;
  DevBuf<float> nodeBox;
  DevBuf<Node8> outNodes;
  // scratch of the batched bottom-level build (kf_blas_batch.cuh)
  DevBuf<BatchGeom> batchGeoms;
  DevBuf<uint32_t> primGeom, batchCounters, batchConvex;
  DevBuf<int> batchBoxes;
  // scratch of the top-level SAH build (k_tlas_sah)
  DevBuf<uint32_t> sahSegOf, sahBinCount, sahPre, sahSegPre;
  DevBuf<int4> sahSegs, sahDecision;
  DevBuf<float2> sahDecisionF;
  DevBuf<int> sahBounds, sahBinBox;
  DevBuf<unsigned long long> sahBestKey;
  uint32_t* sortedVals = nullptr;  // valsA or valsB after the sort
  uint64_t* sortedKeys = nullptr;  // keysA or keysB after the sort
  void release() {
    primBox.release(); sceneBox.release(); keysA.release(); keysB.release(); valsA.release();
    valsB.release(); hist.release(); flags.release(); outPrim.release(); counters.release();
    children.release(); range.release(); parent.release(); wideBinary.release();
    wideMembers.release(); nodeBox.release(); outNodes.release(); slotOfInst.release();
    batchGeoms.release(); primGeom.release(); batchCounters.release(); batchBoxes.release(); batchConvex.release();
    sahSegOf.release(); sahBinCount.release(); sahPre.release(); sahSegPre.release(); sahSegs.release();
    sahDecision.release(); sahDecisionF.release(); sahBounds.release(); sahBinBox.release();
    sahBestKey.release();
  }
};

}  // namespace

struct KfrtContext {
  int device = 0;
  cudaStream_t ownStream = nullptr, stream = nullptr;
  std::string err;
  uint32_t maxGeometry = 128, maxInstances = 256, maxTextures = 128, maxMaterials = 256;

  std::vector<GeomHost> geoms;
  DevBuf<GeomRec> geomTable;
  DevBuf<BlasInfo> blasInfo;
  bool tablesDirty = true;

  DevBuf<KfrtMaterial> mats;
  uint32_t nMats = 0;
  std::vector<TexHost> texs;
  DevBuf<TexRec> texTable;
  bool texDirty = true;
  DevBuf<uchar4> env;
  uint32_t envSize = 0;
  DevBuf<KfrtDirectionalLight> dl;
  DevBuf<KfrtPointLights> pl;
  DevBuf<KfrtActiveLights> al;
  DevBuf<float> alProjView;  // proj * view of every projector slot (PathTrace.rchit:300: `proj * view * p`)
  uint32_t nLightSlots = 0;
  unsigned long long lightMask = 0;
  DevBuf<float> srgbToLinear, srgbThreshold;

  std::vector<KfrtInstance> instHost;
  DevBuf<KfrtInstance> instDev;
  DevBuf<InstRec> instRec;
  DevBuf<int> instBoxInt;  // per-instance world boxes as ordered ints (k_instance_box)
  DevBuf<Node8> tlasNodes;
  uint32_t nTlasNodes = 0;
  bool blasBuilt = false, tlasBuilt = false;
  BuildState blasBuild, tlasBuild;
  // quality watch of the refitted top level (see kfrtRefitTlas)
  DevBuf<float> tlasArea;          // device scalar: area sum of the binary nodes after the last build / refit
  // what the host wants to know of a top-level build or refit comes back asynchronously, in pinned memory
  struct TlasReadback {
    float areaAtBuild, areaAfterRefit;
    uint32_t nWide, depth;
  };
  TlasReadback* tlasRb = nullptr;
  cudaEvent_t tlasBuildReady = nullptr, tlasAreaReady = nullptr;
  bool tlasBuildPending = false, tlasAreaPending = false;
  float tlasAreaAtBuild = 0.0f;
  uint32_t tlasBlasDepth = 0;  // deepest bottom level among the instances of the last build
  bool tlasSahOptIn = false;   // k_tlas_sah may use its large dynamic shared memory on this device
  uint64_t tlasRebuilds = 0;

  // world-space instance subtrees (kf_wsi.cuh): built lazily by the first kfrtRender after the instance
  // set or a transform has changed, when the instanced triangles fit the budget
  bool cullLightSamples = true;  // kfrtSetLightSampleCulling
  bool skipOwnInstance = true;   // kfrtSetOwnInstanceSkip
  int wsiMode = 1;  // 0 never; 1 static scenes within the budget, if a timed probe says they are faster; 2 whenever they fit
  uint64_t wsiMaxTris = uint64_t(4) << 20;
  bool wsiValid = false, wsiDirty = true;
  bool wsiMoving = false;  // kfrtRefitTlas since the last kfrtBuildTlas: mode 1 leaves such scenes on the two-level walk
  int wsiChoice = 0;       // mode 1: 0 not probed yet, 1 the two-level structure won, 2 the instance subtrees won
  float wsiProbeMs[2] = {0.0f, 0.0f};  // what the probe measured: two-level, instance subtrees
  DevBuf<Node8> wsiNodes;
  DevBuf<Tri48> wsiTris;
  DevBuf<KfrtVertex> wsiVerts;
  DevBuf<WsiInst> wsiVis;
  DevBuf<uint32_t> wsiRank;
  uint64_t wsiNodeCount = 0, wsiTriCount = 0, wsiBuilds = 0;
  uint32_t wsiDepth = 0;
  int gridTraceWsi[4] = {0, 0, 0, 0};

  // outputs
  uint32_t nCams = 0, width = 0, height = 0;
  DevBuf<KfrtCamera> cams;
  DevBuf<float4> sum, rgba, albedo, normal;
  DevBuf<int2> hitIds;
  DevBuf<float> hitT, depth;
  DevBuf<uchar4> bgra;
  DevBuf<unsigned long long> counters;
  // The encoded frame starts its way to the host as soon as it is resolved: kfrtResolve queues an
  // asynchronous copy of the BGRA8 buffer into pinned staging memory on a second stream, so that
  // kfrtDownloadBGRA8 (== downloadLatestFrame) only waits for that copy and the next frame's kernels are
  // not held up behind a pageable device -> host transfer.
  cudaStream_t copyStream = nullptr;
  cudaEvent_t resolvedEvent = nullptr, stagedEvent = nullptr;
  uint8_t* bgraStage = nullptr;
  size_t bgraStageCap = 0;
  bool bgraStaged = false;
  KfrtPushConstants lastPc{};
  bool rendered = false;
  int detail = 0;
  uint64_t launches = 0;
  // wavefront scheduler state
  int numSMs = 148;
  size_t batchSlotTarget = size_t(64) << 20;  // path slots per batch (144 B each).  Measured on config 3 at the end of round 2 (ms per 64-spp frame): 16 Mi 213.1, 32 Mi 208.1, 64 Mi 205.7, 128 Mi 204.5 -- for 10 GB more path state and 0.45 s more in the first frame (its allocation)
  bool traceLog = false;  // KFRT_TRACE_LOG=1: per-launch ray count and time of every traversal stage on stderr
  size_t wfSlots = 0;
  bool wfMulti = false;
  DevBuf<float4> wfRayO, wfRayD, wfHitA, wfStateW, wfStateC, wfShadowL, wfShadowC, wfCtx;
  DevBuf<int> wfHitB;
  DevBuf<uint32_t> wfQueue0, wfQueue1, wfShadowQ0, wfShadowQ1, wfCounts;
  int gridTrace[4] = {0, 0, 0, 0}, gridShade[4] = {0, 0, 0, 0}, gridShadow[4] = {0, 0, 0, 0};
  KfrtCounters lastCounters{};
  // per-stage device timers (kfrtSetStageTimers): one event before every launch + one at the end
  bool stageTimers = false;
  std::vector<cudaEvent_t> stageEvents;
  std::vector<int> stageOf;  // stage launched after event i
  size_t stageUsed = 0;
};

// Marks the start of a launch of `stage` on the context's stream (no-op unless timers are on).
static void stageMark(KfrtContext* ctx, int stage) {
  if (!ctx->stageTimers) return;
  if (ctx->stageUsed == ctx->stageEvents.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    ctx->stageEvents.push_back(e);
    ctx->stageOf.push_back(0);
  }
  ctx->stageOf[ctx->stageUsed] = stage;
  cudaEventRecord(ctx->stageEvents[ctx->stageUsed], ctx->stream);
  ctx->stageUsed++;
}

#define KF_FAIL(ctx, code, msg)  \
  do {                           \
    (ctx)->err = (msg);          \
    return (code);               \
  } while (0)

#define KF_CUDA(ctx, expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                         \
      return KFRT_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

// Every entry point makes the context's device current first: a context may be used from any one
// thread at a time, and that thread may have another device current (two contexts on two GPUs in one
// thread, or a context handed to a new thread); allocations, launches and copies below must land on
// the device the context was created on.
#define KF_CHECK_CTX(ctx)                                                                      \
  do {                                                                                         \
    if (!(ctx)) return KFRT_ERR_INVALID;                                                       \
    cudaError_t _e = cudaSetDevice((ctx)->device);                                             \
    if (_e != cudaSuccess) {                                                                   \
      (ctx)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(_e);                    \
      return KFRT_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

static inline unsigned gridFor(size_t n, unsigned block) { return unsigned((n + block - 1) / block); }

// ------------------------------------------------------------------------------------------------
// BVH build sequencing
// ------------------------------------------------------------------------------------------------
// keyBits: the keys have no bit set at or above it, so the passes over higher digits are skipped.
static int radixSort(KfrtContext* ctx, BuildState& st, uint32_t n, int keyBits = 64) {
  const uint32_t numBlocks = gridFor(n, KF_SORT_TILE);
  KF_CUDA(ctx, st.hist.ensure(size_t(256) * numBlocks));
  uint64_t *kin = st.keysA.p, *kout = st.keysB.p;
  uint32_t *vin = st.valsA.p, *vout = st.valsB.p;
  const int passes = std::min(8, std::max(1, (keyBits + 7) / 8));
  for (int pass = 0; pass < passes; pass++) {
    const int shift = pass * 8;
    k_sort_hist<<<numBlocks, KF_SORT_THREADS, 0, ctx->stream>>>(kin, n, shift, st.hist.p, numBlocks);
    k_sort_scan<<<1, 1024, 0, ctx->stream>>>(st.hist.p, 256u * numBlocks);
    k_sort_scatter<<<numBlocks, KF_SORT_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, st.hist.p,
                                                                   numBlocks);
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  // an even number of passes leaves the data in A, an odd one in B
  st.sortedVals = vin;
  st.sortedKeys = kin;
  KF_CUDA(ctx, cudaGetLastError());
  return KFRT_OK;
}

// Binary hierarchy over the instance boxes by top-down binned SAH: one launch of k_tlas_sah
// (kf_bvh_build.cuh), no host round trip, no stream synchronisation.  It is built when the instance
// set changes and when the refit watch trips (per-frame motion is the refit).
#define KF_TLAS_SAH_MAX 65536u
#define KF_TLAS_REBUILD_RATIO 1.1f
static int sahTopLevelHierarchy(KfrtContext* ctx, BuildState& st, uint32_t n) {
  const size_t maxSeg = n / 2 + 1, maxSlots = n / 3 + 2;
  KF_CUDA(ctx, st.sahSegOf.ensure(size_t(2) * n));
  KF_CUDA(ctx, st.sahSegs.ensure(2 * maxSeg));
  KF_CUDA(ctx, st.sahBounds.ensure(6 * maxSlots));
  KF_CUDA(ctx, st.sahBinCount.ensure(3 * KF_TLAS_BINS * maxSlots));
  KF_CUDA(ctx, st.sahBinBox.ensure(size_t(18) * KF_TLAS_BINS * maxSlots));
  KF_CUDA(ctx, st.sahBestKey.ensure(maxSlots));
  KF_CUDA(ctx, st.sahDecision.ensure(maxSeg));
  KF_CUDA(ctx, st.sahDecisionF.ensure(maxSeg));
  KF_CUDA(ctx, st.sahPre.ensure(size_t(n) + 1));
  KF_CUDA(ctx, st.sahSegPre.ensure(2 * (maxSeg + 1)));
  TlasSahArgs a;
  a.n = n;
  a.primBox = st.primBox.p;
  a.vals = st.valsA.p;
  a.valsTmp = st.valsB.p;
  a.children = st.children.p;
  a.range = st.range.p;
  a.parent = st.parent.p;
  a.segOf = st.sahSegOf.p;
  a.segs = st.sahSegs.p;
  a.cbounds = st.sahBounds.p;
  a.binCount = st.sahBinCount.p;
  a.binBox = st.sahBinBox.p;
  a.bestKey = st.sahBestKey.p;
  a.decision = st.sahDecision.p;
  a.decisionF = st.sahDecisionF.p;
  a.pre = st.sahPre.p;
  a.segPre = st.sahSegPre.p;
  size_t shared = 0;
  a.inShared = 0;
  if (n <= KF_TLAS_SHARED_MAX) {
    shared = tlasSahSharedBytes(n);
    if (!ctx->tlasSahOptIn) {  // opt in to more than 48 KB of dynamic shared memory, once per context's device
      KF_CUDA(ctx, cudaFuncSetAttribute(k_tlas_sah, cudaFuncAttributeMaxDynamicSharedMemorySize, int(tlasSahSharedBytes(KF_TLAS_SHARED_MAX))));
      ctx->tlasSahOptIn = true;
    }
    a.inShared = 1;
  }
  k_tlas_sah<<<1, KF_TLAS_SAH_THREADS, shared, ctx->stream>>>(a);
  KF_CUDA(ctx, cudaGetLastError());
  st.sortedVals = st.valsA.p;
  return KFRT_OK;
}

// Scratch of a top-level build over n instances: up to n real nodes plus n instance slots in the wide array.
static int reserveTopLevel(KfrtContext* ctx, BuildState& st, uint32_t n) {
  const size_t maxNodes = size_t(2) * std::max<uint32_t>(n, 1) + 1;
  KF_CUDA(ctx, st.primBox.ensure(size_t(6) * std::max<uint32_t>(n, 1)));
  KF_CUDA(ctx, st.sceneBox.ensure(6));
  KF_CUDA(ctx, st.outPrim.ensure(n));
  KF_CUDA(ctx, st.wideMembers.ensure(size_t(8) * maxNodes));
  KF_CUDA(ctx, st.wideBinary.ensure(maxNodes));
  KF_CUDA(ctx, st.counters.ensure(5));
  KF_CUDA(ctx, st.nodeBox.ensure(size_t(6) * std::max<uint32_t>(n, 1)));
  KF_CUDA(ctx, st.slotOfInst.ensure(std::max<uint32_t>(n, 1)));
  if (n == 1) return KFRT_OK;
  KF_CUDA(ctx, st.hist.ensure(size_t(256) * gridFor(n, KF_SORT_TILE)));
  KF_CUDA(ctx, st.keysA.ensure(n));
  KF_CUDA(ctx, st.keysB.ensure(n));
  KF_CUDA(ctx, st.valsA.ensure(n));
  KF_CUDA(ctx, st.valsB.ensure(n));
  KF_CUDA(ctx, st.children.ensure(n));
  KF_CUDA(ctx, st.range.ensure(n));
  KF_CUDA(ctx, st.parent.ensure(size_t(2) * n));
  KF_CUDA(ctx, st.flags.ensure(n));
  return KFRT_OK;
}

static void orderedBoxToFloat(const int* ib, float* out) {
  for (int k = 0; k < 6; k++) {
    int i = ib[k];
    i = i >= 0 ? i : i ^ 0x7fffffff;
    std::memcpy(out + k, &i, 4);
  }
}

// A batch holds at most KF_BATCH_MAX_GEOMS geometries (the geometry index takes the top 10 key bits) and,
// unless a single geometry is larger, KF_BATCH_MAX_TRIS triangles (what bounds the build scratch: about
// 0.25 KB per triangle).
#define KF_BATCH_MAX_TRIS (size_t(4) << 20)

// Builds the bottom-level structures of geoms[first, first + count) of `list` in one pass (see
// kf_blas_batch.cuh).  One stream synchronisation per batch: node counts, depths and boxes come back
// together before the final storage is sized.
static int buildBlasBatch(KfrtContext* ctx, GeomHost* const* list, uint32_t count) {
  BuildState& st = ctx->blasBuild;
  cudaStream_t stream = ctx->stream;
  std::vector<BatchGeom> table(count);
  uint32_t nTris = 0, nodeSlots = 0, maxTris = 0;
  for (uint32_t i = 0; i < count; i++) {
    const GeomHost& g = *list[i];
    BatchGeom& b = table[i];
    b.verts = g.verts;
    b.idx = g.idx;
    b.matIndex = g.matIndex;
    b.nVerts = g.nVerts;
    b.nTris = g.nIdx / 3;
    b.triOffset = nTris;
    b.nodeOffset = nodeSlots;
    b.nodesAlloc = nullptr;
    b.tris = nullptr;
    b.shade = nullptr;
    b.faceNormal = nullptr;
    nTris += b.nTris;
    nodeSlots += b.nTris + 1;
    maxTris = std::max(maxTris, b.nTris);
  }
  const uint32_t n = nTris;
  KF_CUDA(ctx, st.batchGeoms.ensure(count));
  KF_CUDA(ctx, st.batchCounters.ensure(size_t(KF_BATCH_COUNTERS) * count));
  KF_CUDA(ctx, st.batchBoxes.ensure(size_t(6) * count));
  KF_CUDA(ctx, st.primGeom.ensure(n));
  KF_CUDA(ctx, st.primBox.ensure(size_t(6) * n));
  KF_CUDA(ctx, st.hist.ensure(size_t(256) * gridFor(n, KF_SORT_TILE)));
  KF_CUDA(ctx, st.keysA.ensure(n));
  KF_CUDA(ctx, st.keysB.ensure(n));
  KF_CUDA(ctx, st.valsA.ensure(n));
  KF_CUDA(ctx, st.valsB.ensure(n));
  KF_CUDA(ctx, st.children.ensure(n));
  KF_CUDA(ctx, st.range.ensure(n));
  KF_CUDA(ctx, st.parent.ensure(size_t(2) * n));
  KF_CUDA(ctx, st.flags.ensure(n));
  KF_CUDA(ctx, st.nodeBox.ensure(size_t(6) * n));
  KF_CUDA(ctx, st.outNodes.ensure(nodeSlots));
  KF_CUDA(ctx, st.wideBinary.ensure(nodeSlots));
  KF_CUDA(ctx, st.outPrim.ensure(n));
  KF_CUDA(ctx, cudaMemcpyAsync(st.batchGeoms.p, table.data(), sizeof(BatchGeom) * count, cudaMemcpyHostToDevice, stream));
  k_batch_init<<<gridFor(6 * size_t(count), 256), 256, 0, stream>>>(st.batchBoxes.p, count);
  k_batch_tri_boxes<<<gridFor(n, 256), 256, 0, stream>>>(st.batchGeoms.p, count, n, st.primBox.p, st.primGeom.p, st.batchBoxes.p);
  // which geometries are convex (their answer comes back with the node counts below)
  std::vector<uint32_t> convex(count, 1u);
  KF_CUDA(ctx, st.batchConvex.ensure(count));
  {
    uint32_t maxExamined = 1;
    uint64_t spent = 0;
    for (uint32_t i = 0; i < count; i++) {
      const uint64_t work = uint64_t(table[i].nTris) * table[i].nVerts;
      if (work <= KF_CONVEX_MAX_WORK && spent + work <= KF_CONVEX_BATCH_WORK) {
        spent += work;
        maxExamined = std::max(maxExamined, table[i].nTris);
      } else {
        convex[i] = 0u;  // not examined
      }
    }
    KF_CUDA(ctx, cudaMemcpyAsync(st.batchConvex.p, convex.data(), sizeof(uint32_t) * count, cudaMemcpyHostToDevice, stream));
    const unsigned cx = std::max(1u, std::min<unsigned>(gridFor(maxExamined, 128), std::max(1u, unsigned(ctx->numSMs) * 16u / count)));
    k_batch_convex<<<dim3(cx, count), 128, 0, stream>>>(st.batchGeoms.p, st.batchBoxes.p, st.batchConvex.p);
  }
  if (n > 1) {
    k_batch_morton<<<gridFor(n, 256), 256, 0, stream>>>(st.primBox.p, st.primGeom.p, n, st.batchBoxes.p, st.keysA.p, st.valsA.p);
    int geomBits = 0;
    while ((1u << geomBits) < count) geomBits++;
    int rc = radixSort(ctx, st, n, 3 * KF_BATCH_MORTON_BITS + geomBits);
    if (rc) return rc;
    k_lbvh_hierarchy<<<gridFor(n - 1, 256), 256, 0, stream>>>(st.sortedKeys, int(n), st.children.p, st.range.p, st.parent.p);
    KF_CUDA(ctx, cudaMemsetAsync(st.flags.p, 0, sizeof(uint32_t) * n, stream));
    k_lbvh_bounds<<<gridFor(n, 256), 256, 0, stream>>>(int(n), st.children.p, st.parent.p, st.primBox.p, st.sortedVals,
                                                       st.nodeBox.p, st.flags.p);
  } else {
    const uint32_t zero = 0;
    KF_CUDA(ctx, cudaMemcpyAsync(st.valsA.p, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
    st.sortedVals = st.valsA.p;
  }
  k_batch_roots<<<gridFor(count, 128), 128, 0, stream>>>(st.batchGeoms.p, count, st.range.p, n, st.batchCounters.p, st.wideBinary.p);
  BatchCollapseArgs ca;
  ca.geoms = st.batchGeoms.p;
  ca.nGeoms = count;
  ca.base.n = int(n);
  ca.base.children = st.children.p;
  ca.base.range = st.range.p;
  ca.base.nodeBox = st.nodeBox.p;
  ca.base.primBox = st.primBox.p;
  ca.base.vals = st.sortedVals;
  ca.base.outNodes = st.outNodes.p;
  ca.base.outPrim = st.outPrim.p;
  ca.base.wideBinary = st.wideBinary.p;
  ca.base.wideMembers = nullptr;
  ca.base.counters = st.batchCounters.p;
  ca.base.slotOfInst = nullptr;
  // a level of the largest geometry has at most ~maxTris / 2 nodes; a few blocks per geometry stride over it
  const unsigned bx = std::max(1u, std::min<unsigned>(gridFor(maxTris / 4 + 1, 64), std::max(1u, unsigned(ctx->numSMs) * 16u / count)));
  std::vector<uint32_t> counters(size_t(KF_BATCH_COUNTERS) * count);
  std::vector<int> boxes(size_t(6) * count);
  // a balanced 8-wide tree over n / 2 leaves has log8(n / 2) levels; LBVH trees are a little deeper
  int levels = 4;
  for (uint32_t m = maxTris; m > 16; m >>= 3) levels++;
  for (bool done = false; !done;) {
    for (int level = 0; level < levels; level++) {
      k_batch_collapse_level<<<dim3(bx, count), 64, 0, stream>>>(ca);
      k_batch_next_level<<<gridFor(count, 128), 128, 0, stream>>>(st.batchCounters.p, count);
    }
    KF_CUDA(ctx, cudaMemcpyAsync(counters.data(), st.batchCounters.p, sizeof(uint32_t) * counters.size(), cudaMemcpyDeviceToHost, stream));
    KF_CUDA(ctx, cudaMemcpyAsync(boxes.data(), st.batchBoxes.p, sizeof(int) * boxes.size(), cudaMemcpyDeviceToHost, stream));
    KF_CUDA(ctx, cudaMemcpyAsync(convex.data(), st.batchConvex.p, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, stream));
    KF_CUDA(ctx, cudaStreamSynchronize(stream));
    done = true;
    for (uint32_t i = 0; i < count; i++) {
      const uint32_t* c = &counters[size_t(KF_BATCH_COUNTERS) * i];
      if (c[0] > table[i].nTris + 1) KF_FAIL(ctx, KFRT_ERR_CUDA, "internal: wide node count exceeded its bound");
      if (c[2] < c[3]) done = false;
    }
    levels = 4;
  }
  // final storage: one allocation for the batch, 256-byte aligned pieces
  auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
  size_t bytes = 0;
  std::vector<size_t> offNodes(count), offTris(count), offShade(count), offNormal(count);
  for (uint32_t i = 0; i < count; i++) {
    const uint32_t nWide = counters[size_t(KF_BATCH_COUNTERS) * i];
    offNodes[i] = bytes; bytes = align(bytes + sizeof(Node8) * (size_t(nWide) + 1));
    offTris[i] = bytes; bytes = align(bytes + sizeof(Tri48) * table[i].nTris);
    offShade[i] = bytes; bytes = align(bytes + sizeof(ShadeTri) * table[i].nTris);
    offNormal[i] = bytes; bytes = align(bytes + sizeof(float4) * table[i].nTris);
  }
  auto block = std::make_shared<BlasBlock>();
  block->stream = stream;
  KF_CUDA(ctx, cudaMallocAsync(&block->p, bytes, stream));
  char* base = static_cast<char*>(block->p);
  for (uint32_t i = 0; i < count; i++) {
    GeomHost& g = *list[i];
    g.block = block;
    g.nodesAlloc = reinterpret_cast<Node8*>(base + offNodes[i]);
    g.nodes = g.nodesAlloc + 1;
    g.tris = reinterpret_cast<Tri48*>(base + offTris[i]);
    g.shade = reinterpret_cast<ShadeTri*>(base + offShade[i]);
    g.faceNormal = reinterpret_cast<float4*>(base + offNormal[i]);
    g.convex = convex[i] != 0u;
    g.nNodes = counters[size_t(KF_BATCH_COUNTERS) * i];
    g.depth = counters[size_t(KF_BATCH_COUNTERS) * i + 4];
    orderedBoxToFloat(&boxes[size_t(6) * i], g.box);
    table[i].nodesAlloc = g.nodesAlloc;
    table[i].tris = g.tris;
    table[i].shade = g.shade;
    table[i].faceNormal = g.faceNormal;
    KF_CUDA(ctx, cudaMemsetAsync(g.nodesAlloc, 0, sizeof(Node8), stream));  // the header record
  }
  KF_CUDA(ctx, cudaMemcpyAsync(st.batchGeoms.p, table.data(), sizeof(BatchGeom) * count, cudaMemcpyHostToDevice, stream));
  k_batch_copy_nodes<<<dim3(bx, count), 128, 0, stream>>>(st.batchGeoms.p, st.outNodes.p, st.batchCounters.p);
  uint32_t maxVerts = 1;
  for (uint32_t i = 0; i < count; i++) maxVerts = std::max(maxVerts, table[i].nVerts);
  const unsigned sx = std::max(1u, std::min<unsigned>(gridFor(maxVerts, 1024), std::max(1u, unsigned(ctx->numSMs) * 8u / count)));
  k_batch_spheres<<<dim3(sx, count), 256, 0, stream>>>(st.batchGeoms.p, st.batchBoxes.p);
  k_batch_write_tris<<<gridFor(n, 256), 256, 0, stream>>>(st.batchGeoms.p, st.primGeom.p, st.outPrim.p, n);
  k_batch_write_shade_tris<<<gridFor(n, 256), 256, 0, stream>>>(st.batchGeoms.p, st.primGeom.p, n);
  KF_CUDA(ctx, cudaGetLastError());
  // the table on the host goes out of scope: the copy above must have read it
  KF_CUDA(ctx, cudaStreamSynchronize(stream));
  return KFRT_OK;
}

static int uploadTables(KfrtContext* ctx) {
  const size_t n = ctx->geoms.size();
  std::vector<GeomRec> recs(std::max<size_t>(n, 1));
  std::vector<BlasInfo> infos(std::max<size_t>(n, 1));
  for (size_t i = 0; i < n; i++) {
    const GeomHost& g = ctx->geoms[i];
    recs[i].verts = g.verts;
    recs[i].idx = g.idx;
    recs[i].matIndex = g.matIndex;
    recs[i].nTris = g.nIdx / 3;
    recs[i].flags = (g.opaque ? 1u : 0u) | (g.hide ? 2u : 0u);
    infos[i].nodes = g.nodes;
    infos[i].tris = g.tris;
    infos[i].verts = g.verts;
    infos[i].idx = g.idx;
    infos[i].matIndex = g.matIndex;
    infos[i].shade = g.shade;
    infos[i].faceNormal = g.faceNormal;
    infos[i].nVerts = g.nVerts;
    std::memcpy(infos[i].box, g.box, sizeof(g.box));
    infos[i].flags = ((g.present && g.nodes != nullptr) ? 1u : 0u) | (g.opaque ? 0u : 2u) | (g.convex ? 4u : 0u);
  }
  KF_CUDA(ctx, ctx->geomTable.ensure(recs.size()));
  KF_CUDA(ctx, ctx->blasInfo.ensure(infos.size()));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->geomTable.p, recs.data(), sizeof(GeomRec) * recs.size(),
                               cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->blasInfo.p, infos.data(), sizeof(BlasInfo) * infos.size(),
                               cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->tablesDirty = false;
  return KFRT_OK;
}

static int uploadTexTable(KfrtContext* ctx) {
  std::vector<TexRec> recs(std::max<size_t>(ctx->texs.size(), 1));
  for (size_t i = 0; i < ctx->texs.size(); i++) {
    recs[i].texels = ctx->texs[i].texels;
    recs[i].w = ctx->texs[i].w;
    recs[i].h = ctx->texs[i].h;
  }
  if (ctx->texs.empty()) recs[0] = TexRec{nullptr, 0, 0};
  KF_CUDA(ctx, ctx->texTable.ensure(recs.size()));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->texTable.p, recs.data(), sizeof(TexRec) * recs.size(),
                               cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->texDirty = false;
  return KFRT_OK;
}

// ------------------------------------------------------------------------------------------------
// Render kernels
// ------------------------------------------------------------------------------------------------
struct RenderArgs {
  SceneDev sc;
  const KfrtCamera* cams;
  uint32_t nCams, w, h;
  KfrtPushConstants pc;
  uint32_t s0, s1, clockBase;
  float4* sum;
  float4* albedo;
  float4* normal;
  int2* hitIds;
  float* hitT;
  float* depth;
  unsigned long long* counters;  // [0] paths [1] ext [2] shadow [3] hits [4] nodes [5] tris [6] insts [7] tex
};

// Accumulate + encode (reference PathTrace.rgen:143-163, PostProcessing.frag:11-18 into a
// B8G8R8A8Srgb attachment).  The 8-bit sRGB code is found by binary search over the 255 linear
// thresholds so that the bytes are bit-exact with the oracle (no pow on the device).
KF_D unsigned char encodeSrgb8(const float* __restrict__ thr, float c) {
  if (!(c > 0.0f)) return 0;
  int lo = 0, hi = 255;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (c >= __ldg(thr + mid)) lo = mid + 1; else hi = mid;
  }
  return (unsigned char)lo;
}
__global__ void k_resolve(const float4* __restrict__ sum, float4* __restrict__ rgba, uchar4* __restrict__ bgra,
                          size_t nPixels, uint32_t spp, int frameCount, const float* __restrict__ thr) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nPixels) return;
  const float4 s = sum[i];
  float fc[3] = {cdiv(s.x, float(spp)), cdiv(s.y, float(spp)), cdiv(s.z, float(spp))};
  if (frameCount > 0) {
    const float4 old = rgba[i];
    const float al = cdiv(1.0f, float(frameCount + 1));
    const float om = csub(1.0f, al);
    fc[0] = cadd(cmul(old.x, om), cmul(fc[0], al));
    fc[1] = cadd(cmul(old.y, om), cmul(fc[1], al));
    fc[2] = cadd(cmul(old.z, om), cmul(fc[2], al));
  }
  rgba[i] = make_float4(fc[0], fc[1], fc[2], 1.0f);
  bgra[i] = make_uchar4(encodeSrgb8(thr, fc[2]), encodeSrgb8(thr, fc[1]), encodeSrgb8(thr, fc[0]), 255);
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* kfrtVersion(void) { return "kuafu_b200 kfrt 0.1 (sm_100a)"; }

const char* kfrtLastError(const KfrtContext* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int kfrtCreate(int deviceOrdinal, KfrtContext** out) {
  if (!out) return KFRT_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_createError = std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                    " (kfrt has no CPU fallback)";
    return KFRT_ERR_CUDA;
  }
  if (deviceOrdinal < 0 || deviceOrdinal >= count) {
    g_createError = "device ordinal out of range";
    return KFRT_ERR_INVALID;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, deviceOrdinal);
  if (e != cudaSuccess) {
    g_createError = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
    return KFRT_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_createError = "kfrt is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                    std::to_string(prop.minor);
    return KFRT_ERR_CUDA;
  }
  e = cudaSetDevice(deviceOrdinal);
  if (e != cudaSuccess) {
    g_createError = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return KFRT_ERR_CUDA;
  }
  KfrtContext* ctx = new KfrtContext();
  ctx->device = deviceOrdinal;
  e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    g_createError = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
    delete ctx;
    return KFRT_ERR_CUDA;
  }
  ctx->stream = ctx->ownStream;
  ctx->numSMs = prop.multiProcessorCount;
  {  // BLAS storage comes from the stream-ordered pool; keep freed blocks for the next build
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, deviceOrdinal) == cudaSuccess && pool) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (const char* e = std::getenv("KFRT_TRACE_LOG")) ctx->traceLog = std::atoi(e) != 0;
  if (const char* e = std::getenv("KFRT_SKIP_OWN_INSTANCE")) ctx->skipOwnInstance = std::atoi(e) != 0;
  if (const char* e = std::getenv("KFRT_INSTANCE_SUBTREES")) ctx->wsiMode = std::max(0, std::min(2, std::atoi(e)));
  if (const char* e = std::getenv("KFRT_INSTANCE_SUBTREES_MAX_TRIS")) {
    const long long v = std::atoll(e);
    if (v > 0) ctx->wsiMaxTris = uint64_t(v);
  }
  if (const char* e = std::getenv("KFRT_BATCH_SLOTS")) {
    const long long v = std::atoll(e);
    if (v > 0) ctx->batchSlotTarget = size_t(v);
  }
  // sRGB decode table and encode thresholds (same double-precision formulas as the oracle)
  float lut[256], thr[255];
  for (int i = 0; i < 256; i++) {
    const double c = i / 255.0;
    lut[i] = float(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
  }
  for (int k = 0; k < 255; k++) {
    const double c = (k + 0.5) / 255.0;
    thr[k] = float(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
  }
  bool ok = ctx->srgbToLinear.ensure(256) == cudaSuccess && ctx->srgbThreshold.ensure(255) == cudaSuccess &&
            ctx->dl.ensure(1) == cudaSuccess && ctx->pl.ensure(1) == cudaSuccess && ctx->al.ensure(1) == cudaSuccess && ctx->alProjView.ensure(16 * KFRT_MAX_ACTIVE_LIGHTS) == cudaSuccess &&
            ctx->counters.ensure(16) == cudaSuccess;
  if (ok) {
    ok = cudaMemcpy(ctx->srgbToLinear.p, lut, sizeof(lut), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(ctx->srgbThreshold.p, thr, sizeof(thr), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemset(ctx->dl.p, 0, sizeof(KfrtDirectionalLight)) == cudaSuccess &&
         cudaMemset(ctx->pl.p, 0, sizeof(KfrtPointLights)) == cudaSuccess &&
         cudaMemset(ctx->al.p, 0, sizeof(KfrtActiveLights)) == cudaSuccess;
  }
  if (!ok) {
    g_createError = std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    kfrtDestroy(ctx);
    return KFRT_ERR_CUDA;
  }
  *out = ctx;
  return KFRT_OK;
}

int kfrtDestroy(KfrtContext* ctx) {
  KF_CHECK_CTX(ctx);
  cudaStreamSynchronize(ctx->stream);
  for (auto& g : ctx->geoms) g.freeAll();
  for (auto& t : ctx->texs) cudaFree(t.texels);
  ctx->geomTable.release(); ctx->blasInfo.release(); ctx->mats.release(); ctx->texTable.release();
  ctx->env.release(); ctx->dl.release(); ctx->pl.release(); ctx->al.release(); ctx->alProjView.release();
  ctx->srgbToLinear.release(); ctx->srgbThreshold.release(); ctx->instDev.release();
  ctx->instRec.release(); ctx->instBoxInt.release(); ctx->tlasNodes.release(); ctx->tlasArea.release();
  ctx->wsiNodes.release(); ctx->wsiTris.release(); ctx->wsiVerts.release(); ctx->wsiVis.release(); ctx->wsiRank.release();
  if (ctx->tlasRb) cudaFreeHost(ctx->tlasRb);
  if (ctx->tlasAreaReady) cudaEventDestroy(ctx->tlasAreaReady);
  if (ctx->tlasBuildReady) cudaEventDestroy(ctx->tlasBuildReady);
  ctx->blasBuild.release(); ctx->tlasBuild.release(); ctx->cams.release(); ctx->sum.release();
  ctx->rgba.release(); ctx->albedo.release(); ctx->normal.release(); ctx->hitIds.release();
  ctx->hitT.release(); ctx->depth.release(); ctx->bgra.release(); ctx->counters.release();
  ctx->wfRayO.release(); ctx->wfRayD.release(); ctx->wfHitA.release(); ctx->wfStateW.release();
  ctx->wfStateC.release(); ctx->wfShadowL.release(); ctx->wfShadowC.release(); ctx->wfCtx.release();
  ctx->wfHitB.release(); ctx->wfQueue0.release(); ctx->wfQueue1.release(); ctx->wfShadowQ0.release();
  ctx->wfShadowQ1.release(); ctx->wfCounts.release();
  for (cudaEvent_t e : ctx->stageEvents) cudaEventDestroy(e);
  if (ctx->copyStream) {
    cudaStreamSynchronize(ctx->copyStream);
    cudaStreamDestroy(ctx->copyStream);
  }
  if (ctx->resolvedEvent) cudaEventDestroy(ctx->resolvedEvent);
  if (ctx->stagedEvent) cudaEventDestroy(ctx->stagedEvent);
  if (ctx->bgraStage) cudaFreeHost(ctx->bgraStage);
  if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
  delete ctx;
  return KFRT_OK;
}

int kfrtSetStream(KfrtContext* ctx, void* cudaStream) {
  KF_CHECK_CTX(ctx);
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cudaStream ? static_cast<cudaStream_t>(cudaStream) : ctx->ownStream;
  return KFRT_OK;
}

int kfrtSynchronize(KfrtContext* ctx) {
  KF_CHECK_CTX(ctx);
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KFRT_OK;
}

int kfrtSetLimits(KfrtContext* ctx, uint32_t maxGeometry, uint32_t maxInstances, uint32_t maxTextures,
                  uint32_t maxMaterials) {
  KF_CHECK_CTX(ctx);
  if (!maxGeometry || !maxInstances || !maxTextures || !maxMaterials)
    KF_FAIL(ctx, KFRT_ERR_INVALID, "limits must be non-zero");
  ctx->maxGeometry = maxGeometry;
  ctx->maxInstances = maxInstances;
  ctx->maxTextures = maxTextures;
  ctx->maxMaterials = maxMaterials;
  return KFRT_OK;
}

int kfrtUploadGeometry(KfrtContext* ctx, uint32_t geometryIndex, const KfrtVertex* vertices, uint32_t nVertices,
                       const uint32_t* indices, uint32_t nIndices, const uint32_t* matIndex, uint32_t nMatIndex,
                       int opaque, int hideRender) {
  KF_CHECK_CTX(ctx);
  if (geometryIndex >= ctx->maxGeometry) KF_FAIL(ctx, KFRT_ERR_LIMIT, "geometry index exceeds the geometry limit");
  if (nIndices % 3 != 0) KF_FAIL(ctx, KFRT_ERR_INVALID, "index count must be a multiple of 3");
  if (nIndices && (!vertices || !indices || !matIndex)) KF_FAIL(ctx, KFRT_ERR_INVALID, "null geometry buffer");
  if (nMatIndex < nIndices / 3) KF_FAIL(ctx, KFRT_ERR_INVALID, "matIndex shorter than the primitive count");
  for (uint32_t i = 0; i < nIndices; i++)
    if (indices[i] >= nVertices) KF_FAIL(ctx, KFRT_ERR_INVALID, "vertex index out of range");
  if (ctx->geoms.size() <= geometryIndex) ctx->geoms.resize(geometryIndex + 1);
  GeomHost& g = ctx->geoms[geometryIndex];
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  g.freeAll();
  g.nVerts = nVertices;
  g.nIdx = nIndices;
  g.nMat = nMatIndex;
  g.opaque = opaque != 0;
  g.hide = hideRender != 0;
  if (nIndices) {
    KF_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&g.verts), sizeof(KfrtVertex) * nVertices));
    KF_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&g.idx), sizeof(uint32_t) * nIndices));
    KF_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&g.matIndex), sizeof(uint32_t) * nMatIndex));
    KF_CUDA(ctx, cudaMemcpyAsync(g.verts, vertices, sizeof(KfrtVertex) * nVertices, cudaMemcpyHostToDevice, ctx->stream));
    KF_CUDA(ctx, cudaMemcpyAsync(g.idx, indices, sizeof(uint32_t) * nIndices, cudaMemcpyHostToDevice, ctx->stream));
    KF_CUDA(ctx, cudaMemcpyAsync(g.matIndex, matIndex, sizeof(uint32_t) * nMatIndex, cudaMemcpyHostToDevice, ctx->stream));
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  g.present = true;
  g.dirty = true;
  ctx->tablesDirty = true;
  ctx->tlasBuilt = false;
  return KFRT_OK;
}

int kfrtClearGeometries(KfrtContext* ctx) {
  KF_CHECK_CTX(ctx);
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& g : ctx->geoms) g.freeAll();
  ctx->geoms.clear();
  ctx->instHost.clear();  // they index the geometries that are gone; kfrtSetInstances brings new ones
  ctx->tablesDirty = true;
  ctx->blasBuilt = false;
  ctx->tlasBuilt = false;
  return KFRT_OK;
}

int kfrtUploadMaterials(KfrtContext* ctx, const KfrtMaterial* materials, uint32_t n) {
  KF_CHECK_CTX(ctx);
  if (n > ctx->maxMaterials) KF_FAIL(ctx, KFRT_ERR_LIMIT, "material count exceeds the material limit");
  if (n && !materials) KF_FAIL(ctx, KFRT_ERR_INVALID, "null materials");
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  KF_CUDA(ctx, ctx->mats.ensure(std::max<uint32_t>(n, 1)));
  if (n) KF_CUDA(ctx, cudaMemcpyAsync(ctx->mats.p, materials, sizeof(KfrtMaterial) * n, cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->nMats = n;
  return KFRT_OK;
}

int kfrtUploadTexture(KfrtContext* ctx, uint32_t textureIndex, const uint8_t* rgba8, uint32_t width, uint32_t height) {
  KF_CHECK_CTX(ctx);
  if (textureIndex >= ctx->maxTextures) KF_FAIL(ctx, KFRT_ERR_LIMIT, "texture index exceeds the texture limit");
  if (!rgba8 || !width || !height) KF_FAIL(ctx, KFRT_ERR_INVALID, "empty texture");
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->texs.size() <= textureIndex) ctx->texs.resize(textureIndex + 1);
  TexHost& t = ctx->texs[textureIndex];
  cudaFree(t.texels);
  t.texels = nullptr;
  KF_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&t.texels), size_t(width) * height * 4));
  KF_CUDA(ctx, cudaMemcpyAsync(t.texels, rgba8, size_t(width) * height * 4, cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  t.w = width;
  t.h = height;
  ctx->texDirty = true;
  return KFRT_OK;
}

int kfrtSetEnvironmentCube(KfrtContext* ctx, const uint8_t* const faces[6], uint32_t size) {
  KF_CHECK_CTX(ctx);
  if (!faces || !size) KF_FAIL(ctx, KFRT_ERR_INVALID, "empty cube map");
  for (int f = 0; f < 6; f++)
    if (!faces[f]) KF_FAIL(ctx, KFRT_ERR_INVALID, "null cube face");
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const size_t faceTexels = size_t(size) * size;
  KF_CUDA(ctx, ctx->env.ensure(6 * faceTexels));
  for (int f = 0; f < 6; f++)
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->env.p + f * faceTexels, faces[f], faceTexels * 4, cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->envSize = size;
  return KFRT_OK;
}

int kfrtClearEnvironment(KfrtContext* ctx) {
  KF_CHECK_CTX(ctx);
  ctx->envSize = 0;
  return KFRT_OK;
}

int kfrtSetLights(KfrtContext* ctx, const KfrtDirectionalLight* directional, const KfrtPointLights* points,
                  const KfrtActiveLights* actives) {
  KF_CHECK_CTX(ctx);
  KfrtDirectionalLight d{};
  KfrtPointLights p{};
  KfrtActiveLights a{};
  if (directional) d = *directional;
  if (points) p = *points;
  if (actives) a = *actives;
  uint32_t slots = 0;
  unsigned long long mask = 0;
  if (d.rgbs[0] * d.rgbs[3] != 0 || d.rgbs[1] * d.rgbs[3] != 0 || d.rgbs[2] * d.rgbs[3] != 0) mask |= 1ull;
  for (int i = 0; i < KFRT_MAX_POINT_LIGHTS; i++)
    if (p.rgbs[i][3] > 0) mask |= 1ull << (1 + i);
  for (int i = 0; i < KFRT_MAX_ACTIVE_LIGHTS; i++)
    if (a.front[i][3] > 0) mask |= 1ull << (33 + i);
  for (unsigned long long m = mask; m; m &= m - 1) slots++;
  ctx->lightMask = mask;
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->dl.p, &d, sizeof(d), cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->pl.p, &p, sizeof(p), cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->al.p, &a, sizeof(a), cudaMemcpyHostToDevice, ctx->stream));
  // GLSL evaluates `proj * view * vec4(worldPos, 1)` left to right: the matrix product comes first.
  // Column j of proj * view is proj * (column j of view), each sum taken left to right; volatile keeps
  // the host compiler from contracting the products into FMAs.
  float pv[KFRT_MAX_ACTIVE_LIGHTS][16];
  for (int i = 0; i < KFRT_MAX_ACTIVE_LIGHTS; i++)
    for (int j = 0; j < 4; j++)
      for (int r = 0; r < 4; r++) {
        volatile float acc = a.projMat[i][r] * a.viewMat[i][4 * j + 0];
        volatile float t1 = a.projMat[i][4 + r] * a.viewMat[i][4 * j + 1];
        acc = acc + t1;
        volatile float t2 = a.projMat[i][8 + r] * a.viewMat[i][4 * j + 2];
        acc = acc + t2;
        volatile float t3 = a.projMat[i][12 + r] * a.viewMat[i][4 * j + 3];
        acc = acc + t3;
        pv[i][4 * j + r] = acc;
      }
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->alProjView.p, pv, sizeof(pv), cudaMemcpyHostToDevice, ctx->stream));
  // (the sources are pageable stack memory: cudaMemcpyAsync has staged them before it returns)
  ctx->nLightSlots = slots;
  return KFRT_OK;
}

int kfrtBuildBlas(KfrtContext* ctx) {
  KF_CHECK_CTX(ctx);
  // all geometries uploaded since the last build, in batches (kf_blas_batch.cuh)
  std::vector<GeomHost*> todo;
  for (auto& g : ctx->geoms) {
    if (!g.present || !g.dirty) continue;
    g.block.reset();  // the old structure goes back to the pool behind the work already enqueued
    g.nodes = nullptr;
    g.nodesAlloc = nullptr;
    g.tris = nullptr;
    g.shade = nullptr;
    g.faceNormal = nullptr;
    g.convex = false;
    g.nNodes = 0;
    g.depth = 0;
    g.dirty = false;
    if (g.nIdx / 3 == 0 || g.hide) continue;  // hidden geometries get no bottom level (reference rt.cpp:153-157)
    todo.push_back(&g);
  }
  for (size_t first = 0; first < todo.size();) {
    size_t last = first, tris = 0;
    while (last < todo.size() && last - first < KF_BATCH_MAX_GEOMS &&
           (last == first || tris + todo[last]->nIdx / 3 <= KF_BATCH_MAX_TRIS)) {
      tris += todo[last]->nIdx / 3;
      last++;
    }
    int rc = buildBlasBatch(ctx, todo.data() + first, uint32_t(last - first));
    if (rc) return rc;
    first = last;
  }
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->blasBuilt = true;
  ctx->tablesDirty = true;
  ctx->tlasBuilt = false;
  return KFRT_OK;
}

int kfrtSetInstances(KfrtContext* ctx, const KfrtInstance* instances, uint32_t n) {
  KF_CHECK_CTX(ctx);
  if (n > ctx->maxInstances) KF_FAIL(ctx, KFRT_ERR_LIMIT, "instance count exceeds the instance limit");
  if (n && !instances) KF_FAIL(ctx, KFRT_ERR_INVALID, "null instances");
  for (uint32_t i = 0; i < n; i++)
    if (instances[i].geometryIndex >= ctx->geoms.size() || !ctx->geoms[instances[i].geometryIndex].present)
      KF_FAIL(ctx, KFRT_ERR_INVALID,
              "geometry index is out of bounds (hint for SAPIEN users: are you creating two active renders?)");
  ctx->instHost.assign(instances, instances + n);
  ctx->tlasBuilt = false;
  return KFRT_OK;
}

// World boxes of all instances (input of the top-level build / refit).
static int instanceBoxes(KfrtContext* ctx, bool withSceneBox) {
  const uint32_t n = uint32_t(ctx->instHost.size());
  BuildState& st = ctx->tlasBuild;
  KF_CUDA(ctx, ctx->instDev.ensure(std::max<uint32_t>(n, 1)));
  KF_CUDA(ctx, ctx->instRec.ensure(std::max<uint32_t>(n, 1)));
  KF_CUDA(ctx, st.primBox.ensure(size_t(6) * std::max<uint32_t>(n, 1)));
  KF_CUDA(ctx, st.sceneBox.ensure(6));
  if (n == 0) return KFRT_OK;
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->instDev.p, ctx->instHost.data(), sizeof(KfrtInstance) * n,
                               cudaMemcpyHostToDevice, ctx->stream));
  if (withSceneBox) k_init_scene_box<<<1, 32, 0, ctx->stream>>>(st.sceneBox.p);
  uint32_t maxVerts = 1;
  for (const auto& in : ctx->instHost) maxVerts = std::max(maxVerts, ctx->geoms[in.geometryIndex].nVerts);
  KF_CUDA(ctx, ctx->instBoxInt.ensure(size_t(6) * n));
  k_instance_box_init<<<gridFor(6 * size_t(n), 256), 256, 0, ctx->stream>>>(ctx->instBoxInt.p, n);
  k_instance_box<<<dim3(n, gridFor(maxVerts, KF_BOX_CHUNK)), 128, 0, ctx->stream>>>(
      ctx->instDev.p, n, ctx->blasInfo.p, uint32_t(ctx->geoms.size()), ctx->instBoxInt.p);
  k_instance_box_finish<<<gridFor(n, 128), 128, 0, ctx->stream>>>(ctx->instBoxInt.p, n, st.primBox.p,
                                                                   withSceneBox ? st.sceneBox.p : nullptr);
  KF_CUDA(ctx, cudaGetLastError());
  return KFRT_OK;
}

// Inverse transforms: the per-instance shading records and the InstNode slots of the top-level array.
static int instanceRecords(KfrtContext* ctx) {
  const uint32_t n = uint32_t(ctx->instHost.size());
  BuildState& st = ctx->tlasBuild;
  if (n == 0) return KFRT_OK;
  k_instance_setup<<<gridFor(n, 128), 128, 0, ctx->stream>>>(ctx->instDev.p, n, ctx->blasInfo.p,
                                                             uint32_t(ctx->geoms.size()), ctx->instRec.p,
                                                             st.primBox.p, st.slotOfInst.p, ctx->tlasNodes.p);
  KF_CUDA(ctx, cudaGetLastError());
  return KFRT_OK;
}

// What the last top-level build reported (node count, depth, area sum), once its readback has landed.
static int consumeTlasBuild(KfrtContext* ctx) {
  ctx->tlasBuildPending = false;
  BuildState& st = ctx->tlasBuild;
  st.nWide = ctx->tlasRb->nWide;
  st.depth = ctx->tlasRb->depth;
  ctx->nTlasNodes = st.nWide;
  ctx->tlasAreaAtBuild = ctx->tlasRb->areaAtBuild;
  // the traversal stack holds at most one entry per level of the two trees a ray is in
  if (st.depth + ctx->tlasBlasDepth > uint32_t(KF_STACK_SHARED + KF_STACK)) {
    ctx->tlasBuilt = false;
    KF_FAIL(ctx, KFRT_ERR_LIMIT, "acceleration structure deeper than the traversal stack (KF_STACK)");
  }
  return KFRT_OK;
}

// Top-level build from ctx->instHost: instance boxes, binary hierarchy (binned SAH), 8-wide collapse with
// every instance as a child slot, instance records, and the area sum the later refits are compared with.
// Everything is enqueued on the stream; what the host wants back (node count, depth, area) arrives in
// pinned memory behind an event.  wait: block until it has (kfrtBuildTlas); the rebuild the refit watch
// triggers on the per-frame path does not wait.
static int buildTopLevel(KfrtContext* ctx, bool wait) {
  const uint32_t n = uint32_t(ctx->instHost.size());
  // geometries may have been uploaded again (fewer of them, or one left empty) since kfrtSetInstances
  for (const auto& in : ctx->instHost)
    if (in.geometryIndex >= ctx->geoms.size() || !ctx->geoms[in.geometryIndex].present)
      KF_FAIL(ctx, KFRT_ERR_INVALID, "an instance refers to a geometry that no longer exists; call kfrtSetInstances again");
  int rc = instanceBoxes(ctx, true);
  if (rc) return rc;
  ctx->tlasBuildPending = ctx->tlasAreaPending = false;
  ctx->tlasAreaAtBuild = 0.0f;
  ctx->wsiValid = false;  // the instance subtrees are rebuilt by the next kfrtRender (buildWsi)
  ctx->wsiDirty = true;
  if (n == 0) {
    ctx->nTlasNodes = 0;
    ctx->tlasBuilt = true;
    return KFRT_OK;
  }
  BuildState& st = ctx->tlasBuild;
  st.n = n;
  rc = reserveTopLevel(ctx, st, n);
  if (rc) return rc;
  const size_t maxNodes = size_t(2) * n + 1;
  KF_CUDA(ctx, ctx->tlasNodes.ensure(maxNodes));
  KF_CUDA(ctx, ctx->tlasArea.ensure(1));
  if (!ctx->tlasRb) KF_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->tlasRb), sizeof(*ctx->tlasRb)));
  if (!ctx->tlasAreaReady) KF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->tlasAreaReady, cudaEventDisableTiming));
  if (!ctx->tlasBuildReady) KF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->tlasBuildReady, cudaEventDisableTiming));
  // the wide nodes are written where the traversal reads them; the InstNode slots among them come
  // from k_instance_setup (instanceRecords) and get defined bytes until then
  KF_CUDA(ctx, cudaMemsetAsync(ctx->tlasNodes.p, 0, sizeof(Node8) * maxNodes, ctx->stream));
  KF_CUDA(ctx, cudaMemsetAsync(ctx->tlasArea.p, 0, sizeof(float), ctx->stream));
  if (n == 1) {
    k_single_instance_root<<<1, 32, 0, ctx->stream>>>(st.primBox.p, ctx->tlasNodes.p, st.wideBinary.p,
                                                      st.wideMembers.p, st.slotOfInst.p, st.nodeBox.p);
    const uint32_t done[5] = {2u, 0u, 2u, 2u, 1u};
    KF_CUDA(ctx, cudaMemcpyAsync(st.counters.p, done, sizeof(done), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    if (n <= KF_TLAS_SAH_MAX) {
      rc = sahTopLevelHierarchy(ctx, st, n);
      if (rc) return rc;
    } else {
      k_morton<<<gridFor(n, 256), 256, 0, ctx->stream>>>(st.primBox.p, n, st.sceneBox.p, st.keysA.p, st.valsA.p);
      rc = radixSort(ctx, st, n);
      if (rc) return rc;
      k_lbvh_hierarchy<<<gridFor(n - 1, 256), 256, 0, ctx->stream>>>(st.sortedKeys, int(n), st.children.p,
                                                                     st.range.p, st.parent.p);
    }
    KF_CUDA(ctx, cudaMemsetAsync(st.flags.p, 0, sizeof(uint32_t) * n, ctx->stream));
    k_lbvh_bounds<<<gridFor(n, 256), 256, 0, ctx->stream>>>(int(n), st.children.p, st.parent.p, st.primBox.p,
                                                            st.sortedVals, st.nodeBox.p, st.flags.p);
    const uint32_t init[5] = {1u, 0u, 0u, 1u, 0u};
    const int zero = 0;
    KF_CUDA(ctx, cudaMemcpyAsync(st.counters.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    KF_CUDA(ctx, cudaMemcpyAsync(st.wideBinary.p, &zero, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CollapseArgs a;
    a.n = int(n);
    a.children = st.children.p;
    a.range = st.range.p;
    a.nodeBox = st.nodeBox.p;
    a.primBox = st.primBox.p;
    a.vals = st.sortedVals;
    a.outNodes = ctx->tlasNodes.p;
    a.outPrim = st.outPrim.p;
    a.wideBinary = st.wideBinary.p;
    a.wideMembers = st.wideMembers.p;
    a.counters = st.counters.p;
    a.slotOfInst = st.slotOfInst.p;
    k_collapse_all<true><<<1, KF_COLLAPSE_ALL_THREADS, 0, ctx->stream>>>(a);
    k_area_sum<<<std::min<unsigned>(gridFor(n - 1, 256), 64u), 256, 0, ctx->stream>>>(st.nodeBox.p, n - 1, ctx->tlasArea.p);
  }
  rc = instanceRecords(ctx);
  if (rc) return rc;
  KF_CUDA(ctx, cudaMemcpyAsync(&ctx->tlasRb->areaAtBuild, ctx->tlasArea.p, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  KF_CUDA(ctx, cudaMemcpyAsync(&ctx->tlasRb->nWide, st.counters.p + 0, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  KF_CUDA(ctx, cudaMemcpyAsync(&ctx->tlasRb->depth, st.counters.p + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  KF_CUDA(ctx, cudaEventRecord(ctx->tlasBuildReady, ctx->stream));
  KF_CUDA(ctx, cudaGetLastError());
  ctx->tlasBlasDepth = 0;
  for (const auto& in : ctx->instHost) ctx->tlasBlasDepth = std::max(ctx->tlasBlasDepth, ctx->geoms[in.geometryIndex].depth);
  ctx->tlasBuildPending = true;
  ctx->nTlasNodes = uint32_t(maxNodes);  // its bound, until the readback lands
  st.nWide = uint32_t(maxNodes);
  ctx->tlasBuilt = true;
  if (wait) {
    KF_CUDA(ctx, cudaEventSynchronize(ctx->tlasBuildReady));
    return consumeTlasBuild(ctx);
  }
  return KFRT_OK;
}

// World-space instance subtrees (kf_wsi.cuh) from the current instances, transforms and top level: the
// vertices of every visible instance in world space, the batched bottom-level builder over them (one
// pseudo-geometry per instance), then the assembly -- top-level nodes, subtree roots in the instance
// slots, the other subtree nodes behind, object-space triangle records in leaf order.  One stream
// synchronisation (the collapse reports its node counts), as in buildBlasBatch.
static int buildWsi(KfrtContext* ctx) {
  ctx->wsiValid = false;
  ctx->wsiDirty = false;
  if (ctx->wsiMode == 0 || ctx->instHost.empty()) return KFRT_OK;
  if (ctx->wsiMode == 1 && (ctx->wsiMoving || ctx->wsiChoice == 1)) return KFRT_OK;
  const uint32_t nInst = uint32_t(ctx->instHost.size());
  std::vector<WsiInst> vis;
  std::vector<uint32_t> rank(nInst, 0xffffffffu);
  std::vector<BatchGeom> table;
  uint64_t nTris64 = 0, nVerts64 = 0;
  for (uint32_t i = 0; i < nInst; i++) {
    const uint32_t gi = ctx->instHost[i].geometryIndex;
    if (gi >= ctx->geoms.size()) return KFRT_OK;
    const GeomHost& g = ctx->geoms[gi];
    if (!g.present || !g.nodes || g.nIdx / 3 == 0) continue;
    nTris64 += g.nIdx / 3;
    nVerts64 += g.nVerts;
    if (vis.size() >= KF_BATCH_MAX_GEOMS || nTris64 > ctx->wsiMaxTris || nVerts64 >= (uint64_t(1) << 31)) return KFRT_OK;
    WsiInst v{};
    v.instance = i;
    v.vertOffset = uint32_t(nVerts64 - g.nVerts);
    v.flags = g.opaque ? 0u : 2u;
    rank[i] = uint32_t(vis.size());
    vis.push_back(v);
  }
  if (vis.empty()) return KFRT_OK;
  if (ctx->tlasBuildPending) {  // the top level's node count is needed below
    KF_CUDA(ctx, cudaEventSynchronize(ctx->tlasBuildReady));
    int rc = consumeTlasBuild(ctx);
    if (rc) return rc;
  }
  const uint32_t count = uint32_t(vis.size());
  const uint32_t n = uint32_t(nTris64);
  BuildState& st = ctx->blasBuild;
  cudaStream_t stream = ctx->stream;
  KF_CUDA(ctx, ctx->wsiVerts.ensure(size_t(nVerts64)));
  KF_CUDA(ctx, ctx->wsiVis.ensure(count));
  KF_CUDA(ctx, ctx->wsiRank.ensure(nInst));
  table.resize(count);
  uint32_t triAt = 0, nodeSlots = 0, maxTris = 0, maxVerts = 1;
  for (uint32_t k = 0; k < count; k++) {
    const GeomHost& g = ctx->geoms[ctx->instHost[vis[k].instance].geometryIndex];
    BatchGeom& b = table[k];
    b.verts = ctx->wsiVerts.p + vis[k].vertOffset;
    b.idx = g.idx;
    b.matIndex = g.matIndex;
    b.nVerts = g.nVerts;
    b.nTris = g.nIdx / 3;
    b.triOffset = triAt;
    b.nodeOffset = nodeSlots;
    b.nodesAlloc = nullptr;
    b.tris = nullptr;
    b.shade = nullptr;
    triAt += b.nTris;
    nodeSlots += b.nTris + 1;
    maxTris = std::max(maxTris, b.nTris);
    maxVerts = std::max(maxVerts, b.nVerts);
  }
  KF_CUDA(ctx, st.batchGeoms.ensure(count));
  KF_CUDA(ctx, st.batchCounters.ensure(size_t(KF_BATCH_COUNTERS) * count));
  KF_CUDA(ctx, st.batchBoxes.ensure(size_t(6) * count));
  KF_CUDA(ctx, st.primGeom.ensure(n));
  KF_CUDA(ctx, st.primBox.ensure(size_t(6) * n));
  KF_CUDA(ctx, st.hist.ensure(size_t(256) * gridFor(n, KF_SORT_TILE)));
  KF_CUDA(ctx, st.keysA.ensure(n));
  KF_CUDA(ctx, st.keysB.ensure(n));
  KF_CUDA(ctx, st.valsA.ensure(n));
  KF_CUDA(ctx, st.valsB.ensure(n));
  KF_CUDA(ctx, st.children.ensure(n));
  KF_CUDA(ctx, st.range.ensure(n));
  KF_CUDA(ctx, st.parent.ensure(size_t(2) * n));
  KF_CUDA(ctx, st.flags.ensure(n));
  KF_CUDA(ctx, st.nodeBox.ensure(size_t(6) * n));
  KF_CUDA(ctx, st.outNodes.ensure(nodeSlots));
  KF_CUDA(ctx, st.wideBinary.ensure(nodeSlots));
  KF_CUDA(ctx, st.outPrim.ensure(n));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->wsiVis.p, vis.data(), sizeof(WsiInst) * count, cudaMemcpyHostToDevice, stream));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->wsiRank.p, rank.data(), sizeof(uint32_t) * nInst, cudaMemcpyHostToDevice, stream));
  KF_CUDA(ctx, cudaMemcpyAsync(st.batchGeoms.p, table.data(), sizeof(BatchGeom) * count, cudaMemcpyHostToDevice, stream));
  const unsigned vx = std::max(1u, std::min<unsigned>(gridFor(maxVerts, 256), std::max(1u, unsigned(ctx->numSMs) * 8u / count)));
  k_wsi_world_verts<<<dim3(vx, count), 256, 0, stream>>>(ctx->wsiVis.p, ctx->instDev.p, ctx->blasInfo.p, ctx->wsiVerts.p);
  k_batch_init<<<gridFor(6 * size_t(count), 256), 256, 0, stream>>>(st.batchBoxes.p, count);
  k_batch_tri_boxes<<<gridFor(n, 256), 256, 0, stream>>>(st.batchGeoms.p, count, n, st.primBox.p, st.primGeom.p, st.batchBoxes.p);
  if (n > 1) {
    k_batch_morton<<<gridFor(n, 256), 256, 0, stream>>>(st.primBox.p, st.primGeom.p, n, st.batchBoxes.p, st.keysA.p, st.valsA.p);
    int geomBits = 0;
    while ((1u << geomBits) < count) geomBits++;
    int rc = radixSort(ctx, st, n, 3 * KF_BATCH_MORTON_BITS + geomBits);
    if (rc) return rc;
    k_lbvh_hierarchy<<<gridFor(n - 1, 256), 256, 0, stream>>>(st.sortedKeys, int(n), st.children.p, st.range.p, st.parent.p);
    KF_CUDA(ctx, cudaMemsetAsync(st.flags.p, 0, sizeof(uint32_t) * n, stream));
    k_lbvh_bounds<<<gridFor(n, 256), 256, 0, stream>>>(int(n), st.children.p, st.parent.p, st.primBox.p, st.sortedVals,
                                                       st.nodeBox.p, st.flags.p);
  } else {
    const uint32_t zero = 0;
    KF_CUDA(ctx, cudaMemcpyAsync(st.valsA.p, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
    st.sortedVals = st.valsA.p;
  }
  k_batch_roots<<<gridFor(count, 128), 128, 0, stream>>>(st.batchGeoms.p, count, st.range.p, n, st.batchCounters.p, st.wideBinary.p);
  BatchCollapseArgs ca;
  ca.geoms = st.batchGeoms.p;
  ca.nGeoms = count;
  ca.base.n = int(n);
  ca.base.children = st.children.p;
  ca.base.range = st.range.p;
  ca.base.nodeBox = st.nodeBox.p;
  ca.base.primBox = st.primBox.p;
  ca.base.vals = st.sortedVals;
  ca.base.outNodes = st.outNodes.p;
  ca.base.outPrim = st.outPrim.p;
  ca.base.wideBinary = st.wideBinary.p;
  ca.base.wideMembers = nullptr;
  ca.base.counters = st.batchCounters.p;
  ca.base.slotOfInst = nullptr;
  const unsigned bx = std::max(1u, std::min<unsigned>(gridFor(maxTris / 4 + 1, 64), std::max(1u, unsigned(ctx->numSMs) * 16u / count)));
  std::vector<uint32_t> counters(size_t(KF_BATCH_COUNTERS) * count);
  int levels = 4;
  for (uint32_t m = maxTris; m > 16; m >>= 3) levels++;
  for (bool done = false; !done;) {
    for (int level = 0; level < levels; level++) {
      k_batch_collapse_level<<<dim3(bx, count), 64, 0, stream>>>(ca);
      k_batch_next_level<<<gridFor(count, 128), 128, 0, stream>>>(st.batchCounters.p, count);
    }
    KF_CUDA(ctx, cudaMemcpyAsync(counters.data(), st.batchCounters.p, sizeof(uint32_t) * counters.size(), cudaMemcpyDeviceToHost, stream));
    KF_CUDA(ctx, cudaStreamSynchronize(stream));
    done = true;
    for (uint32_t k = 0; k < count; k++) {
      const uint32_t* c = &counters[size_t(KF_BATCH_COUNTERS) * k];
      if (c[0] > table[k].nTris + 1) KF_FAIL(ctx, KFRT_ERR_CUDA, "internal: wide node count exceeded its bound");
      if (c[2] < c[3]) done = false;
    }
    levels = 4;
  }
  // assembly: [top level incl. the instance slots][nodes 1.. of every subtree]
  const uint32_t nTop = ctx->tlasBuild.nWide;
  uint64_t nodeAt = nTop;
  uint32_t depth = 0;
  for (uint32_t k = 0; k < count; k++) {
    const uint32_t nWide = counters[size_t(KF_BATCH_COUNTERS) * k];
    vis[k].nodeStart = uint32_t(nodeAt);
    nodeAt += nWide - 1;
    depth = std::max(depth, counters[size_t(KF_BATCH_COUNTERS) * k + 4]);
  }
  if (nodeAt >= (uint64_t(1) << 32)) return KFRT_OK;
  if (ctx->tlasBuild.depth + depth > uint32_t(KF_STACK_SHARED + KF_STACK))
    KF_FAIL(ctx, KFRT_ERR_LIMIT, "acceleration structure deeper than the traversal stack (KF_STACK)");
  KF_CUDA(ctx, ctx->wsiNodes.ensure(size_t(nodeAt)));
  KF_CUDA(ctx, ctx->wsiTris.ensure(n));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->wsiVis.p, vis.data(), sizeof(WsiInst) * count, cudaMemcpyHostToDevice, stream));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->wsiNodes.p, ctx->tlasNodes.p, sizeof(Node8) * nTop, cudaMemcpyDeviceToDevice, stream));
  k_wsi_empty_slots<<<gridFor(nInst, 128), 128, 0, stream>>>(ctx->wsiRank.p, nInst, ctx->tlasBuild.slotOfInst.p, ctx->wsiNodes.p);
  k_wsi_place_nodes<<<dim3(bx, count), 128, 0, stream>>>(ctx->wsiVis.p, st.batchGeoms.p, st.batchCounters.p, st.outNodes.p,
                                                         ctx->tlasBuild.slotOfInst.p, ctx->wsiNodes.p);
  k_wsi_write_tris<<<gridFor(n, 256), 256, 0, stream>>>(ctx->wsiVis.p, st.batchGeoms.p, ctx->instDev.p, ctx->blasInfo.p,
                                                        st.primGeom.p, st.outPrim.p, n, ctx->wsiTris.p);
  KF_CUDA(ctx, cudaGetLastError());
  // the tables on the host go out of scope: the copies above must have read them
  KF_CUDA(ctx, cudaStreamSynchronize(stream));
  ctx->wsiNodeCount = nodeAt;
  ctx->wsiTriCount = n;
  ctx->wsiDepth = ctx->tlasBuild.depth + depth;
  ctx->wsiBuilds++;
  ctx->wsiValid = true;
  return KFRT_OK;
}

int kfrtBuildTlas(KfrtContext* ctx) {
  KF_CHECK_CTX(ctx);
  if (!ctx->blasBuilt) KF_FAIL(ctx, KFRT_ERR_NOT_BUILT, "kfrtBuildTlas before kfrtBuildBlas");
  for (auto& g : ctx->geoms)
    if (g.present && g.dirty) KF_FAIL(ctx, KFRT_ERR_NOT_BUILT, "geometry uploaded after the last kfrtBuildBlas");
  if (ctx->tablesDirty) {
    int rc = uploadTables(ctx);
    if (rc) return rc;
  }
  ctx->wsiMoving = false;
  ctx->wsiChoice = 0;
  return buildTopLevel(ctx, true);
}

int kfrtRefitTlas(KfrtContext* ctx, const float* transforms, uint32_t n) {
  KF_CHECK_CTX(ctx);
  if (!ctx->tlasBuilt) KF_FAIL(ctx, KFRT_ERR_NOT_BUILT, "kfrtRefitTlas before kfrtBuildTlas");
  if (n != ctx->instHost.size()) KF_FAIL(ctx, KFRT_ERR_INVALID, "transform count differs from the instance count");
  if (n == 0) return KFRT_OK;
  if (!transforms) KF_FAIL(ctx, KFRT_ERR_INVALID, "null transforms");
  for (uint32_t i = 0; i < n; i++) std::memcpy(ctx->instHost[i].transform, transforms + 16 * i, 64);
  ctx->wsiValid = false;
  ctx->wsiDirty = true;
  ctx->wsiMoving = true;
  // A refit keeps the hierarchy of the last build; when the instances have moved far from where they
  // were then, its boxes overlap and every ray pays (articulated scene: 6.3 top-level node visits per
  // ray after a build, 38 after 40 frames of refits).  The area sum of the last refit (read back
  // asynchronously, so one frame late and without a stall) decides: past KF_TLAS_REBUILD_RATIO times
  // the value at build time the top level is built again from the new transforms -- enqueued like the
  // refit it replaces, without waiting for it.
  if (ctx->tlasBuildPending && cudaEventQuery(ctx->tlasBuildReady) == cudaSuccess) {
    int rc = consumeTlasBuild(ctx);
    if (rc) return rc;
  }
  if (ctx->tlasAreaPending && cudaEventQuery(ctx->tlasAreaReady) == cudaSuccess) {
    ctx->tlasAreaPending = false;
    if (ctx->tlasAreaAtBuild > 0.0f && ctx->tlasRb->areaAfterRefit > KF_TLAS_REBUILD_RATIO * ctx->tlasAreaAtBuild) {
      ctx->tlasRebuilds++;
      return buildTopLevel(ctx, false);
    }
  }
  int rc = instanceBoxes(ctx, false);
  if (rc) return rc;
  BuildState& st = ctx->tlasBuild;
  if (n > 1) {
    KF_CUDA(ctx, cudaMemsetAsync(st.flags.p, 0, sizeof(uint32_t) * n, ctx->stream));
    k_lbvh_bounds<<<gridFor(n, 256), 256, 0, ctx->stream>>>(int(n), st.children.p, st.parent.p, st.primBox.p,
                                                            st.sortedVals, st.nodeBox.p, st.flags.p);
    k_requantise<<<gridFor(st.nWide, 128), 128, 0, ctx->stream>>>(st.counters.p, st.wideMembers.p, st.wideBinary.p,
                                                                  st.nodeBox.p, st.primBox.p, st.sortedVals,
                                                                  ctx->tlasNodes.p, 0);
  } else {
    k_requantise<<<1, 32, 0, ctx->stream>>>(st.counters.p, st.wideMembers.p, nullptr, st.nodeBox.p, st.primBox.p,
                                            st.sortedVals, ctx->tlasNodes.p, 1);
  }
  rc = instanceRecords(ctx);
  if (rc) return rc;
  if (n > 1 && !ctx->tlasAreaPending && ctx->tlasRb) {
    KF_CUDA(ctx, cudaMemsetAsync(ctx->tlasArea.p, 0, sizeof(float), ctx->stream));
    k_area_sum<<<std::min<unsigned>(gridFor(n - 1, 256), 64u), 256, 0, ctx->stream>>>(st.nodeBox.p, n - 1, ctx->tlasArea.p);
    KF_CUDA(ctx, cudaMemcpyAsync(&ctx->tlasRb->areaAfterRefit, ctx->tlasArea.p, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    KF_CUDA(ctx, cudaEventRecord(ctx->tlasAreaReady, ctx->stream));
    ctx->tlasAreaPending = true;
  }
  KF_CUDA(ctx, cudaGetLastError());
  return KFRT_OK;
}

int kfrtSetLightSampleCulling(KfrtContext* ctx, int on) {
  KF_CHECK_CTX(ctx);
  ctx->cullLightSamples = on != 0;
  return KFRT_OK;
}

int kfrtSetOwnInstanceSkip(KfrtContext* ctx, int on) {
  KF_CHECK_CTX(ctx);
  ctx->skipOwnInstance = on != 0;
  return KFRT_OK;
}

int kfrtSetInstanceSubtrees(KfrtContext* ctx, int mode, uint64_t maxTriangles) {
  KF_CHECK_CTX(ctx);
  if (mode < 0 || mode > 2) KF_FAIL(ctx, KFRT_ERR_INVALID, "instance-subtree mode must be 0, 1 or 2");
  ctx->wsiMode = mode;
  ctx->wsiChoice = 0;
  if (maxTriangles) ctx->wsiMaxTris = maxTriangles;
  ctx->wsiValid = false;
  ctx->wsiDirty = true;
  return KFRT_OK;
}

int kfrtGetBvhStats(KfrtContext* ctx, KfrtBvhStats* out) {
  KF_CHECK_CTX(ctx);
  if (!out) KF_FAIL(ctx, KFRT_ERR_INVALID, "null stats");
  if (ctx->tlasBuildPending) {  // a rebuild on the refit path: its node count is still on its way
    KF_CUDA(ctx, cudaEventSynchronize(ctx->tlasBuildReady));
    int rc = consumeTlasBuild(ctx);
    if (rc) return rc;
  }
  std::memset(out, 0, sizeof(*out));
  for (auto& g : ctx->geoms) {
    if (!g.present) continue;
    out->blasCount++;
    out->triangleCount += g.nIdx / 3;
    out->blasNodeCount += g.nNodes;
  }
  out->instanceCount = uint32_t(ctx->instHost.size());
  for (auto& in : ctx->instHost)
    if (in.geometryIndex < ctx->geoms.size()) out->instancedTriangles += ctx->geoms[in.geometryIndex].nIdx / 3;
  out->tlasNodeCount = ctx->nTlasNodes;
  out->nodeBytes = sizeof(Node8);
  out->triangleBytes = sizeof(Tri48);
  out->instanceBytes = ctx->wsiValid ? 48u : uint32_t(sizeof(InstNode));  // subtrees: the three rows of the inverse, per change of instance
  out->tlasRebuilds = uint32_t(ctx->tlasRebuilds);
  out->subtreeNodeCount = ctx->wsiValid ? ctx->wsiNodeCount : 0;
  out->subtreeTriangles = ctx->wsiValid ? ctx->wsiTriCount : 0;
  out->subtreeDepth = ctx->wsiValid ? ctx->wsiDepth : 0;
  out->subtreeBuilds = uint32_t(ctx->wsiBuilds);
  return KFRT_OK;
}


}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Wavefront scheduler (host side): sizes the path-state buffers, then sequences the stage kernels.
// No host synchronisation inside a frame: queue sizes stay on the device and every stage runs as a
// persistent grid (a multiple of the SM count) that strides over its queue.
// ------------------------------------------------------------------------------------------------
template <typename K>
static int persistentGrid(KfrtContext* ctx, K kernel, int blockSize) {
  int perSm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, blockSize, 0) != cudaSuccess || perSm < 1) perSm = 1;
  return ctx->numSMs * perSm;
}

// One traversal stage launch: closest hit (any == false) or occlusion.
static void launchTrace(KfrtContext* ctx, const TraceArgs& te, bool any, bool detail) {
  const int v = (any ? 2 : 0) + (detail ? 1 : 0);
  if (te.sc.wsiNodes) {
    switch (v) {
      case 0: k_wf_trace_wsi<false, false><<<ctx->gridTraceWsi[0], 128, 0, ctx->stream>>>(te); break;
      case 1: k_wf_trace_wsi<false, true><<<ctx->gridTraceWsi[1], 128, 0, ctx->stream>>>(te); break;
      case 2: k_wf_trace_wsi<true, false><<<ctx->gridTraceWsi[2], 128, 0, ctx->stream>>>(te); break;
      default: k_wf_trace_wsi<true, true><<<ctx->gridTraceWsi[3], 128, 0, ctx->stream>>>(te); break;
    }
    return;
  }
  switch (v) {
    case 0: k_wf_trace<false, false><<<ctx->gridTrace[0], 128, 0, ctx->stream>>>(te); break;
    case 1: k_wf_trace<false, true><<<ctx->gridTrace[1], 128, 0, ctx->stream>>>(te); break;
    case 2: k_wf_trace<true, false><<<ctx->gridTrace[2], 128, 0, ctx->stream>>>(te); break;
    default: k_wf_trace<true, true><<<ctx->gridTrace[3], 128, 0, ctx->stream>>>(te); break;
  }
}

// Shade and shadow-resolve launches: variant = MULTI << 1 | DETAIL (the grid table's index); the culling of light
// samples and the own-instance skip are template parameters of the kernels (kf_wavefront.cuh).
template <bool MULTI, bool DETAIL>
static void launchShadeVariant(KfrtContext* ctx, int grid, const WfArgs& a, int q, uint32_t depth) {
  cudaStream_t st = ctx->stream;
  const int sel = (ctx->cullLightSamples ? 2 : 0) | (ctx->skipOwnInstance ? 1 : 0);
  switch (sel) {
    case 0: k_wf_shade<MULTI, DETAIL, false, false><<<grid, KF_SHADE_THREADS, 0, st>>>(a, q, depth); break;
    case 1: k_wf_shade<MULTI, DETAIL, false, true><<<grid, KF_SHADE_THREADS, 0, st>>>(a, q, depth); break;
    case 2: k_wf_shade<MULTI, DETAIL, true, false><<<grid, KF_SHADE_THREADS, 0, st>>>(a, q, depth); break;
    default: k_wf_shade<MULTI, DETAIL, true, true><<<grid, KF_SHADE_THREADS, 0, st>>>(a, q, depth); break;
  }
}
static void launchShade(KfrtContext* ctx, int variant, const WfArgs& a, int q, uint32_t depth) {
  switch (variant) {
    case 0: launchShadeVariant<false, false>(ctx, ctx->gridShade[0], a, q, depth); break;
    case 1: launchShadeVariant<false, true>(ctx, ctx->gridShade[1], a, q, depth); break;
    case 2: launchShadeVariant<true, false>(ctx, ctx->gridShade[2], a, q, depth); break;
    default: launchShadeVariant<true, true>(ctx, ctx->gridShade[3], a, q, depth); break;
  }
}
template <bool MULTI, bool DETAIL>
static void launchShadowResolveVariant(KfrtContext* ctx, int grid, const WfArgs& a, int sq, int qNext, uint32_t depth) {
  if (ctx->cullLightSamples) k_wf_shadow_resolve<MULTI, DETAIL, true><<<grid, 128, 0, ctx->stream>>>(a, sq, qNext, depth);
  else k_wf_shadow_resolve<MULTI, DETAIL, false><<<grid, 128, 0, ctx->stream>>>(a, sq, qNext, depth);
}
static void launchShadowResolve(KfrtContext* ctx, int variant, const WfArgs& a, int sq, int qNext, uint32_t depth) {
  switch (variant) {
    case 0: launchShadowResolveVariant<false, false>(ctx, ctx->gridShadow[0], a, sq, qNext, depth); break;
    case 1: launchShadowResolveVariant<false, true>(ctx, ctx->gridShadow[1], a, sq, qNext, depth); break;
    case 2: launchShadowResolveVariant<true, false>(ctx, ctx->gridShadow[2], a, sq, qNext, depth); break;
    default: launchShadowResolveVariant<true, true>(ctx, ctx->gridShadow[3], a, sq, qNext, depth); break;
  }
}

// Path-state buffers of the wavefront scheduler for `slots` path slots (they only ever grow).
static int ensurePathState(KfrtContext* ctx, size_t slots, bool multi) {
  if (slots >= (size_t(1) << 31)) KF_FAIL(ctx, KFRT_ERR_INVALID, "too many pixels for one batch");
  KF_CUDA(ctx, ctx->wfRayO.ensure(slots));
  KF_CUDA(ctx, ctx->wfRayD.ensure(slots));
  KF_CUDA(ctx, ctx->wfHitA.ensure(slots));
  KF_CUDA(ctx, ctx->wfHitB.ensure(slots));
  KF_CUDA(ctx, ctx->wfStateW.ensure(slots));
  KF_CUDA(ctx, ctx->wfStateC.ensure(slots));
  KF_CUDA(ctx, ctx->wfShadowL.ensure(slots));
  KF_CUDA(ctx, ctx->wfShadowC.ensure(slots));
  if (multi) KF_CUDA(ctx, ctx->wfCtx.ensure(slots * 6));
  KF_CUDA(ctx, ctx->wfQueue0.ensure(slots));
  KF_CUDA(ctx, ctx->wfQueue1.ensure(slots));
  KF_CUDA(ctx, ctx->wfShadowQ0.ensure(slots));
  if (multi) KF_CUDA(ctx, ctx->wfShadowQ1.ensure(slots));
  KF_CUDA(ctx, ctx->wfCounts.ensure(8));
  return KFRT_OK;
}

static int renderWavefront(KfrtContext* ctx, const RenderArgs& ra) {
  const uint32_t tilesX = (ra.w + 7) / 8, tilesY = (ra.h + 3) / 4;
  const size_t slotsPerSample = size_t(ra.nCams) * tilesX * tilesY * 32;
  const uint32_t nSamples = ra.s1 - ra.s0;
  ctx->launches = 0;
  if (nSamples == 0) {
    KF_CUDA(ctx, cudaMemsetAsync(ra.sum, 0, sizeof(float4) * size_t(ra.nCams) * ra.w * ra.h, ctx->stream));
    return KFRT_OK;
  }
  uint32_t batch = uint32_t(std::max<size_t>(1, ctx->batchSlotTarget / slotsPerSample));
  batch = std::min(batch, nSamples);
  const bool multi = ctx->nLightSlots > 1;
  {
    int rc = ensurePathState(ctx, slotsPerSample * batch, multi);
    if (rc) return rc;
  }
  if (!ctx->gridTrace[0]) {
    ctx->gridTrace[0] = persistentGrid(ctx, k_wf_trace<false, false>, 128);
    ctx->gridTrace[1] = persistentGrid(ctx, k_wf_trace<false, true>, 128);
    ctx->gridTrace[2] = persistentGrid(ctx, k_wf_trace<true, false>, 128);
    ctx->gridTrace[3] = persistentGrid(ctx, k_wf_trace<true, true>, 128);
    ctx->gridTraceWsi[0] = persistentGrid(ctx, k_wf_trace_wsi<false, false>, 128);
    ctx->gridTraceWsi[1] = persistentGrid(ctx, k_wf_trace_wsi<false, true>, 128);
    ctx->gridTraceWsi[2] = persistentGrid(ctx, k_wf_trace_wsi<true, false>, 128);
    ctx->gridTraceWsi[3] = persistentGrid(ctx, k_wf_trace_wsi<true, true>, 128);
    // (sized for the default variants; the others share their launch bounds)
    ctx->gridShade[0] = persistentGrid(ctx, k_wf_shade<false, false, true, true>, KF_SHADE_THREADS);
    ctx->gridShade[1] = persistentGrid(ctx, k_wf_shade<false, true, true, true>, KF_SHADE_THREADS);
    ctx->gridShade[2] = persistentGrid(ctx, k_wf_shade<true, false, true, true>, KF_SHADE_THREADS);
    ctx->gridShade[3] = persistentGrid(ctx, k_wf_shade<true, true, true, true>, KF_SHADE_THREADS);
    ctx->gridShadow[0] = persistentGrid(ctx, k_wf_shadow_resolve<false, false, true>, 128);
    ctx->gridShadow[1] = persistentGrid(ctx, k_wf_shadow_resolve<false, true, true>, 128);
    ctx->gridShadow[2] = persistentGrid(ctx, k_wf_shadow_resolve<true, false, true>, 128);
    ctx->gridShadow[3] = persistentGrid(ctx, k_wf_shadow_resolve<true, true, true>, 128);
  }

  WfArgs a;
  a.sc = ra.sc;
  a.b.rayO = ctx->wfRayO.p;
  a.b.rayD = ctx->wfRayD.p;
  a.b.hitA = ctx->wfHitA.p;
  a.b.hitB = ctx->wfHitB.p;
  a.b.stateW = ctx->wfStateW.p;
  a.b.stateC = ctx->wfStateC.p;
  a.b.shadowL = ctx->wfShadowL.p;
  a.b.shadowC = ctx->wfShadowC.p;
  a.b.ctx = ctx->wfCtx.p;
  a.b.queue[0] = ctx->wfQueue0.p;
  a.b.queue[1] = ctx->wfQueue1.p;
  a.b.shadowQueue[0] = ctx->wfShadowQ0.p;
  a.b.shadowQueue[1] = multi ? ctx->wfShadowQ1.p : ctx->wfShadowQ0.p;
  a.b.counts = ctx->wfCounts.p;
  a.cams = ra.cams;
  a.nCams = ra.nCams;
  a.w = ra.w;
  a.h = ra.h;
  a.tilesX = tilesX;
  a.tilesY = tilesY;
  a.slotsPerSample = uint32_t(slotsPerSample);
  a.firstSample = ra.s0;
  a.pc = ra.pc;
  a.clockBase = ra.clockBase;
  a.sum = ra.sum;
  a.albedo = ra.albedo;
  a.normal = ra.normal;
  a.hitIds = ra.hitIds;
  a.hitT = ra.hitT;
  a.depth = ra.depth;
  a.counters = ra.counters;
  const int d = ctx->detail ? 1 : 0;
  const int variant = (multi ? 2 : 0) + d;
  cudaStream_t st = ctx->stream;
  // debugging aid (KFRT_TRACE_LOG): synchronous per-launch timing of the traversal stages
  cudaEvent_t logEv[2] = {nullptr, nullptr};
  if (ctx->traceLog) { cudaEventCreate(&logEv[0]); cudaEventCreate(&logEv[1]); }
  auto logBegin = [&]() { if (ctx->traceLog) cudaEventRecord(logEv[0], st); };
  auto logEnd = [&](const char* what, uint32_t depth, const uint32_t* countDev) {
    if (!ctx->traceLog) return;
    cudaEventRecord(logEv[1], st);
    cudaEventSynchronize(logEv[1]);
    float ms = 0.0f;
    uint32_t n = 0;
    cudaEventElapsedTime(&ms, logEv[0], logEv[1]);
    cudaMemcpy(&n, countDev, sizeof(n), cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[kfrt] %s depth %u: %u rays, %.3f ms, %.1f Mrays/s\n", what, depth, n, ms, n / (ms * 1e3 + 1e-9));
  };
  for (uint32_t b0 = ra.s0; b0 < ra.s1; b0 += batch) {
    a.batchBegin = b0;
    a.batchCount = std::min(batch, ra.s1 - b0);
    KF_CUDA(ctx, cudaMemsetAsync(a.b.counts, 0, 8 * sizeof(uint32_t), st));
    stageMark(ctx, KFRT_STAGE_RAYGEN);
    // one thread per pixel slot (it loops over the batch's samples): the hardware block scheduler balances
    // the pixels better than a persistent grid with ~9 pixels x 32 samples per thread (-4 % on the stage)
    k_wf_raygen<<<gridFor(slotsPerSample, 256), 256, 0, st>>>(a);
    ctx->launches++;
    for (uint32_t depth = 0; depth <= ra.pc.maxPathDepth; depth++) {
      const int q = int(depth & 1u);
      // closest hit of the extension queue; resets the queues the later stages of this bounce fill
      TraceArgs te;
      te.sc = a.sc;
      te.queue = a.b.queue[q];
      te.count = a.b.counts + q;
      te.fetch = a.b.counts + 4;
      te.rayO = a.b.rayO;
      te.rayD = a.b.rayD;
      te.seedSrc = a.b.stateW;
      te.hitA = a.b.hitA;
      te.hitB = a.b.hitB;
      te.clear0 = a.b.counts + (q ^ 1);
      te.clear1 = a.b.counts + 2;
      te.clear2 = a.b.counts + 3;
      te.counters = a.counters;
      te.rayCounter = 1;
      te.detailBase = 4;
      stageMark(ctx, KFRT_STAGE_TRACE_CLOSEST);
      logBegin();
      launchTrace(ctx, te, false, d != 0);
      logEnd("closest", depth, te.count);
      stageMark(ctx, KFRT_STAGE_SHADE);
      launchShade(ctx, variant, a, q, depth);
      ctx->launches += 2;
      const uint32_t rounds = ctx->nLightSlots;
      for (uint32_t l = 0; l < rounds; l++) {
        const int sq = multi ? int(l & 1u) : 0;
        TraceArgs ts = te;
        ts.queue = a.b.shadowQueue[sq];
        ts.count = a.b.counts + 2 + sq;
        ts.fetch = a.b.counts + 5;
        ts.rayD = a.b.shadowL;
        ts.clear0 = ts.clear1 = ts.clear2 = nullptr;
        ts.rayCounter = 2;
        ts.detailBase = 8;
        stageMark(ctx, KFRT_STAGE_TRACE_OCCLUSION);
        logBegin();
        launchTrace(ctx, ts, true, d != 0);
        logEnd("occlusion", depth, ts.count);
        stageMark(ctx, KFRT_STAGE_SHADOW_RESOLVE);
        launchShadowResolve(ctx, variant, a, sq, q ^ 1, depth);
        ctx->launches += 2;
        if (multi && l + 1 < rounds) {
          stageMark(ctx, KFRT_STAGE_OTHER);
          k_wf_clear_count<<<1, 1, 0, st>>>(a.b.counts + 2 + sq);
          ctx->launches++;
        }
      }
    }
    stageMark(ctx, KFRT_STAGE_FINISH);
    k_wf_finish<<<gridFor(slotsPerSample, 256), 256, 0, st>>>(a);
    ctx->launches++;
  }
  stageMark(ctx, KFRT_STAGE_END);
  if (ctx->traceLog) { cudaEventDestroy(logEv[0]); cudaEventDestroy(logEv[1]); }
  KF_CUDA(ctx, cudaGetLastError());
  return KFRT_OK;
}

extern "C" {

// Mode 1 of kfrtSetInstanceSubtrees: which structure is faster depends on the scene AND on how many rays
// a launch carries (instance subtrees save the instance phase and a few dependent fetches per ray; the
// two-level walk has the exact per-instance culls, fewer node steps and the smaller working set --
// measured on the BASELINE configs at full size: config 2 -15 % and config 1 -5 % of the closest-hit stage
// with subtrees, config 3 +4 %, config 4 +10 %; yet a ONE-sample frame of config 4 is 6 % faster with
// subtrees, because small launches are bound by latency, not by throughput).  So the first kfrtRender after
// a build renders a batch of the frame it was asked for -- up to 32 Mi path slots of it, the size at
// which the ratio has settled -- once with each structure, and keeps the faster; the subtrees must win by
// 2 %.  Both give the same buffers bit for bit, and the frame proper overwrites what the probes wrote.
#define KF_PROBE_SLOTS (size_t(32) << 20)
static int probeStructures(KfrtContext* ctx, const RenderArgs& full) {
  RenderArgs p = full;
  const size_t slotsPerSample = size_t(full.nCams) * ((full.w + 7) / 8) * ((full.h + 3) / 4) * 32;
  const uint32_t samples = uint32_t(std::max<size_t>(1, KF_PROBE_SLOTS / slotsPerSample));
  p.s1 = std::min(full.s1, p.s0 + samples);
  const bool timers = ctx->stageTimers, log = ctx->traceLog;
  ctx->stageTimers = false;
  ctx->traceLog = false;  // (its per-launch synchronisation would be what the probe measures)
  cudaEvent_t ev[2] = {nullptr, nullptr};
  KF_CUDA(ctx, cudaEventCreate(&ev[0]));
  KF_CUDA(ctx, cudaEventCreate(&ev[1]));
  float ms[2] = {3.0e38f, 3.0e38f};
  // the path state of the frame proper is sized first (so that no allocation falls into a timed pass), and a
  // one-sample pass with each structure has loaded its kernels
  int rc = KFRT_OK;
  {
    const uint32_t nSamples = full.s1 - full.s0;
    const uint32_t batch = std::min(uint32_t(std::max<size_t>(1, ctx->batchSlotTarget / slotsPerSample)), nSamples);
    rc = ensurePathState(ctx, slotsPerSample * batch, ctx->nLightSlots > 1);
  }
  for (int pass = 0; pass < 2 && rc == KFRT_OK; pass++) {
    RenderArgs q = p;
    if (pass == 0) q.s1 = q.s0 + 1;
    for (int which = 0; which < 2 && rc == KFRT_OK; which++) {
      q.sc.wsiNodes = which ? ctx->wsiNodes.p : nullptr;
      q.sc.wsiTris = which ? ctx->wsiTris.p : nullptr;
      cudaEventRecord(ev[0], ctx->stream);
      rc = renderWavefront(ctx, q);
      cudaEventRecord(ev[1], ctx->stream);
      if (rc == KFRT_OK && pass == 1 && cudaEventSynchronize(ev[1]) == cudaSuccess) cudaEventElapsedTime(&ms[which], ev[0], ev[1]);
    }
  }
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  ctx->stageTimers = timers;
  ctx->traceLog = log;
  ctx->stageUsed = 0;
  if (rc) return rc;
  ctx->wsiProbeMs[0] = ms[0];
  ctx->wsiProbeMs[1] = ms[1];
  ctx->wsiChoice = ms[1] < 0.98f * ms[0] ? 2 : 1;
  if (ctx->traceLog)
    std::fprintf(stderr, "[kfrt] structure probe (%u samples): two-level %.3f ms, instance subtrees %.3f ms -> %s\n",
                 p.s1 - p.s0, ms[0], ms[1], ctx->wsiChoice == 2 ? "instance subtrees" : "two-level");
  if (ctx->wsiChoice == 1) ctx->wsiValid = false;  // (its storage stays for the next build)
  KF_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * 16, ctx->stream));
  return KFRT_OK;
}

static int ensureOutputs(KfrtContext* ctx, uint32_t nCams, uint32_t w, uint32_t h) {
  const size_t np = size_t(nCams) * w * h;
  const bool changed = nCams != ctx->nCams || w != ctx->width || h != ctx->height;
  KF_CUDA(ctx, ctx->cams.ensure(nCams));
  KF_CUDA(ctx, ctx->sum.ensure(np));
  KF_CUDA(ctx, ctx->rgba.ensure(np));
  KF_CUDA(ctx, ctx->albedo.ensure(np));
  KF_CUDA(ctx, ctx->normal.ensure(np));
  KF_CUDA(ctx, ctx->hitIds.ensure(np));
  KF_CUDA(ctx, ctx->hitT.ensure(np));
  KF_CUDA(ctx, ctx->depth.ensure(np));
  KF_CUDA(ctx, ctx->bgra.ensure(np));
  if (changed) {
    if (ctx->bgraStaged) KF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->stagedEvent, 0));
    ctx->bgraStaged = false;
    KF_CUDA(ctx, cudaMemsetAsync(ctx->rgba.p, 0, sizeof(float4) * np, ctx->stream));
    KF_CUDA(ctx, cudaMemsetAsync(ctx->bgra.p, 0, sizeof(uchar4) * np, ctx->stream));
  }
  ctx->nCams = nCams;
  ctx->width = w;
  ctx->height = h;
  return KFRT_OK;
}

static SceneDev sceneDev(KfrtContext* ctx) {
  SceneDev sc{};
  sc.tlasNodes = ctx->nTlasNodes ? ctx->tlasNodes.p : nullptr;
  sc.wsiNodes = ctx->wsiValid ? ctx->wsiNodes.p : nullptr;
  sc.wsiTris = ctx->wsiValid ? ctx->wsiTris.p : nullptr;
  sc.inst = ctx->instRec.p;
  sc.instSsbo = ctx->instDev.p;
  sc.geoms = ctx->geomTable.p;
  sc.mats = ctx->mats.p;
  sc.texs = ctx->texTable.p;
  sc.nTex = uint32_t(ctx->texs.size());
  sc.nInst = uint32_t(ctx->instHost.size());
  sc.envFaces = ctx->envSize ? ctx->env.p : nullptr;
  sc.envSize = ctx->envSize;
  sc.srgbToLinear = ctx->srgbToLinear.p;
  sc.dl = ctx->dl.p;
  sc.pl = ctx->pl.p;
  sc.al = ctx->al.p;
  sc.alProjView = ctx->alProjView.p;
  sc.lightMask = ctx->lightMask;
  return sc;
}

int kfrtRender(KfrtContext* ctx, const KfrtCamera* cameras, uint32_t nCameras, uint32_t width, uint32_t height,
               const KfrtPushConstants* pc, uint32_t sampleBegin, uint32_t sampleEnd, uint32_t clockBase) {
  KF_CHECK_CTX(ctx);
  if (!cameras || !nCameras || !width || !height || !pc) KF_FAIL(ctx, KFRT_ERR_INVALID, "bad render arguments");
  if (sampleEnd < sampleBegin || sampleEnd > pc->sampleRatePerPixel)
    KF_FAIL(ctx, KFRT_ERR_INVALID, "sample range must lie inside [0, sampleRatePerPixel]");
  if (!ctx->tlasBuilt) KF_FAIL(ctx, KFRT_ERR_NOT_BUILT, "kfrtRender before kfrtBuildTlas");
  // material indices are trusted like the reference trusts its SSBO; materials must exist
  if (ctx->nMats == 0 && !ctx->instHost.empty()) KF_FAIL(ctx, KFRT_ERR_INVALID, "no materials uploaded");
  if (ctx->texDirty) {
    int rc = uploadTexTable(ctx);
    if (rc) return rc;
  }
  int rc = ensureOutputs(ctx, nCameras, width, height);
  if (rc) return rc;
  if (ctx->wsiDirty) {
    rc = buildWsi(ctx);
    if (rc) return rc;
  }
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->cams.p, cameras, sizeof(KfrtCamera) * nCameras, cudaMemcpyHostToDevice, ctx->stream));
  KF_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * 16, ctx->stream));
  RenderArgs a;
  a.sc = sceneDev(ctx);
  a.cams = ctx->cams.p;
  a.nCams = nCameras;
  a.w = width;
  a.h = height;
  a.pc = *pc;
  a.s0 = sampleBegin;
  a.s1 = sampleEnd;
  a.clockBase = clockBase;
  a.sum = ctx->sum.p;
  a.albedo = ctx->albedo.p;
  a.normal = ctx->normal.p;
  a.hitIds = ctx->hitIds.p;
  a.hitT = ctx->hitT.p;
  a.depth = ctx->depth.p;
  a.counters = ctx->counters.p;
  ctx->stageUsed = 0;
  if (ctx->wsiMode == 1 && ctx->wsiValid && ctx->wsiChoice == 0 && sampleEnd > sampleBegin) {
    rc = probeStructures(ctx, a);
    if (rc) return rc;
    a.sc = sceneDev(ctx);
  }
  rc = renderWavefront(ctx, a);
  if (rc) return rc;
  ctx->lastPc = *pc;
  ctx->rendered = true;
  ctx->lastCounters = KfrtCounters{};
  ctx->lastCounters.paths = uint64_t(nCameras) * width * height * (sampleEnd - sampleBegin);
  return KFRT_OK;
}

int kfrtResolve(KfrtContext* ctx) {
  KF_CHECK_CTX(ctx);
  if (!ctx->rendered) KF_FAIL(ctx, KFRT_ERR_INVALID, "kfrtResolve before kfrtRender");
  const size_t np = size_t(ctx->nCams) * ctx->width * ctx->height;
  if (!ctx->copyStream) {
    KF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
    KF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->resolvedEvent, cudaEventDisableTiming));
    KF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stagedEvent, cudaEventDisableTiming));
  }
  // the staging copy of the previous frame reads the buffer this launch overwrites
  if (ctx->bgraStaged) KF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->stagedEvent, 0));
  k_resolve<<<gridFor(np, 256), 256, 0, ctx->stream>>>(ctx->sum.p, ctx->rgba.p, ctx->bgra.p, np,
                                                       ctx->lastPc.sampleRatePerPixel, ctx->lastPc.frameCount,
                                                       ctx->srgbThreshold.p);
  KF_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  if (ctx->bgraStageCap < np * 4) {
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->copyStream));
    if (ctx->bgraStage) cudaFreeHost(ctx->bgraStage);
    ctx->bgraStage = nullptr;
    ctx->bgraStageCap = 0;
    KF_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->bgraStage), np * 4));
    ctx->bgraStageCap = np * 4;
  }
  KF_CUDA(ctx, cudaEventRecord(ctx->resolvedEvent, ctx->stream));
  KF_CUDA(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->resolvedEvent, 0));
  KF_CUDA(ctx, cudaMemcpyAsync(ctx->bgraStage, ctx->bgra.p, np * 4, cudaMemcpyDeviceToHost, ctx->copyStream));
  KF_CUDA(ctx, cudaEventRecord(ctx->stagedEvent, ctx->copyStream));
  ctx->bgraStaged = true;
  return KFRT_OK;
}

// NCCL is reached the way the reference reaches libcuda (src/cuda_dl.cpp:10-71): dlopen + dlsym,
// so libkfrt has no link-time dependency on it.
int kfrtReduceNccl(KfrtContext* ctx, void* ncclComm, int root) {
  KF_CHECK_CTX(ctx);
  if (!ctx->rendered || !ncclComm) KF_FAIL(ctx, KFRT_ERR_INVALID, "kfrtReduceNccl needs a rendered frame and a communicator");
  typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  typedef int (*ReduceFn)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);
  // resolved once per process (contexts on several threads may get here together); a failed lookup
  // is remembered with its reason and never leaves a half-filled table behind
  static std::once_flag once;
  static AllReduceFn allReduce = nullptr;
  static ReduceFn reduce = nullptr;
  static std::string loadError;
  std::call_once(once, [] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      const char* why = dlerror();
      loadError = std::string("cannot load libnccl.so.2: ") + (why ? why : "unknown error");
      return;
    }
    AllReduceFn ar = reinterpret_cast<AllReduceFn>(dlsym(lib, "ncclAllReduce"));
    ReduceFn rd = reinterpret_cast<ReduceFn>(dlsym(lib, "ncclReduce"));
    if (!ar || !rd) {
      loadError = "ncclAllReduce/ncclReduce not found in libnccl.so.2";
      return;
    }
    allReduce = ar;
    reduce = rd;
  });
  if (!allReduce || !reduce) KF_FAIL(ctx, KFRT_ERR_NCCL, loadError);
  const size_t count = size_t(ctx->nCams) * ctx->width * ctx->height * 4;
  const int ncclFloat32 = 7, ncclSum = 0;
  int r = root < 0 ? allReduce(ctx->sum.p, ctx->sum.p, count, ncclFloat32, ncclSum, ncclComm, ctx->stream)
                   : reduce(ctx->sum.p, ctx->sum.p, count, ncclFloat32, ncclSum, root, ncclComm, ctx->stream);
  if (r != 0) KF_FAIL(ctx, KFRT_ERR_NCCL, "NCCL reduce failed with code " + std::to_string(r));
  return KFRT_OK;
}

static int bufferOf(KfrtContext* ctx, int kind, void** p, size_t* perPixel) {
  switch (kind) {
    case KFRT_AUX_RGBA32F: *p = ctx->rgba.p; *perPixel = 16; return KFRT_OK;
    case KFRT_AUX_ALBEDO32F: *p = ctx->albedo.p; *perPixel = 16; return KFRT_OK;
    case KFRT_AUX_NORMAL32F: *p = ctx->normal.p; *perPixel = 16; return KFRT_OK;
    case KFRT_AUX_HIT_IDS: *p = ctx->hitIds.p; *perPixel = 8; return KFRT_OK;
    case KFRT_AUX_HIT_T: *p = ctx->hitT.p; *perPixel = 4; return KFRT_OK;
    case KFRT_AUX_DEPTH: *p = ctx->depth.p; *perPixel = 4; return KFRT_OK;
    case KFRT_AUX_SUM32F: *p = ctx->sum.p; *perPixel = 16; return KFRT_OK;
    case KFRT_AUX_BGRA8: *p = ctx->bgra.p; *perPixel = 4; return KFRT_OK;
    default: return KFRT_ERR_INVALID;
  }
}

int kfrtDownloadAux(KfrtContext* ctx, uint32_t camera, int kind, void* dst, size_t nbytes) {
  KF_CHECK_CTX(ctx);
  if (!ctx->rendered) KF_FAIL(ctx, KFRT_ERR_INVALID, "nothing rendered yet");
  if (camera >= ctx->nCams || !dst) KF_FAIL(ctx, KFRT_ERR_INVALID, "bad camera index or destination");
  const size_t np = size_t(ctx->width) * ctx->height;
  if (kind == KFRT_AUX_SEGMENTATION) {
    if (nbytes != np * 4) KF_FAIL(ctx, KFRT_ERR_INVALID, "destination size mismatch");
    std::vector<int2> ids(np);
    KF_CUDA(ctx, cudaMemcpyAsync(ids.data(), ctx->hitIds.p + camera * np, np * 8, cudaMemcpyDeviceToHost, ctx->stream));
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int32_t* o = static_cast<int32_t*>(dst);
    for (size_t i = 0; i < np; i++) o[i] = ids[i].x;
    return KFRT_OK;
  }
  void* p = nullptr;
  size_t pp = 0;
  if (bufferOf(ctx, kind, &p, &pp)) KF_FAIL(ctx, KFRT_ERR_INVALID, "unknown aux kind");
  if (nbytes != np * pp) KF_FAIL(ctx, KFRT_ERR_INVALID, "destination size mismatch");
  if (kind == KFRT_AUX_BGRA8 && ctx->bgraStaged) {  // already on its way to (or in) pinned host memory
    KF_CUDA(ctx, cudaEventSynchronize(ctx->stagedEvent));
    std::memcpy(dst, ctx->bgraStage + camera * np * 4, nbytes);
    return KFRT_OK;
  }
  KF_CUDA(ctx, cudaMemcpyAsync(dst, static_cast<char*>(p) + camera * np * pp, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KFRT_OK;
}

int kfrtDownloadBGRA8(KfrtContext* ctx, uint32_t camera, uint8_t* dst, size_t nbytes) {
  return kfrtDownloadAux(ctx, camera, KFRT_AUX_BGRA8, dst, nbytes);
}

int kfrtMapBGRA8(KfrtContext* ctx, uint32_t camera, const uint8_t** bytes, size_t* nbytes) {
  KF_CHECK_CTX(ctx);
  if (!bytes || !nbytes) KF_FAIL(ctx, KFRT_ERR_INVALID, "null out pointer");
  if (!ctx->rendered || !ctx->bgraStaged) KF_FAIL(ctx, KFRT_ERR_INVALID, "no resolved frame to map (kfrtResolve first)");
  if (camera >= ctx->nCams) KF_FAIL(ctx, KFRT_ERR_INVALID, "bad camera index");
  KF_CUDA(ctx, cudaEventSynchronize(ctx->stagedEvent));
  const size_t np = size_t(ctx->width) * ctx->height;
  *bytes = ctx->bgraStage + camera * np * 4;
  *nbytes = np * 4;
  return KFRT_OK;
}

int kfrtGetDeviceBuffer(KfrtContext* ctx, int kind, void** devicePtr, size_t* nbytes) {
  KF_CHECK_CTX(ctx);
  if (!devicePtr || !nbytes) KF_FAIL(ctx, KFRT_ERR_INVALID, "null out pointer");
  void* p = nullptr;
  size_t pp = 0;
  if (bufferOf(ctx, kind, &p, &pp)) KF_FAIL(ctx, KFRT_ERR_INVALID, "unknown aux kind");
  *devicePtr = p;
  *nbytes = size_t(ctx->nCams) * ctx->width * ctx->height * pp;
  return KFRT_OK;
}

int kfrtSetDetailCounters(KfrtContext* ctx, int detail) {
  KF_CHECK_CTX(ctx);
  ctx->detail = detail ? 1 : 0;
  return KFRT_OK;
}

int kfrtSetStageTimers(KfrtContext* ctx, int on) {
  KF_CHECK_CTX(ctx);
  ctx->stageTimers = on != 0;
  ctx->stageUsed = 0;
  return KFRT_OK;
}

int kfrtGetStageTimes(KfrtContext* ctx, KfrtStageTimes* out) {
  KF_CHECK_CTX(ctx);
  if (!out) KF_FAIL(ctx, KFRT_ERR_INVALID, "null stage times");
  std::memset(out, 0, sizeof(*out));
  if (ctx->stageUsed < 2) return KFRT_OK;
  KF_CUDA(ctx, cudaEventSynchronize(ctx->stageEvents[ctx->stageUsed - 1]));
  for (size_t i = 0; i + 1 < ctx->stageUsed; i++) {
    const int stage = ctx->stageOf[i];
    if (stage < 0 || stage >= KFRT_STAGE_END) continue;
    float ms = 0.0f;
    KF_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->stageEvents[i], ctx->stageEvents[i + 1]));
    out->ms[stage] += ms;
    out->launches[stage]++;
  }
  return KFRT_OK;
}

int kfrtGetCounters(KfrtContext* ctx, KfrtCounters* out) {
  KF_CHECK_CTX(ctx);
  if (!out) KF_FAIL(ctx, KFRT_ERR_INVALID, "null counters");
  unsigned long long h[16] = {0};
  if (ctx->rendered) {
    KF_CUDA(ctx, cudaMemcpyAsync(h, ctx->counters.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  *out = ctx->lastCounters;
  out->extensionRays = h[1];
  out->shadowRays = h[2];
  out->extensionHits = h[3];
  out->nodeVisits = h[4] + h[8];
  out->triangleTests = h[5] + h[9];
  out->instanceVisits = h[6] + h[10];
  out->shadowNodeVisits = h[8];
  out->shadowTriangleTests = h[9];
  out->shadowInstanceVisits = h[10];
  out->textureFetches = h[7];
  out->tlasNodeVisits = h[12];
  out->instanceEntries = h[13];
  out->shadowRaysSkipped = h[11];
  out->kernelLaunches = ctx->launches;
  return KFRT_OK;
}

}  // extern "C"
