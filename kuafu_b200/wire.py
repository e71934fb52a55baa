"""numpy dtypes of the wire structs in include/kf_rt.h (bit-for-bit the reference's UBO/SSBO
layouts: vertex.hpp:28-33, geometry.hpp:70-99, camera.hpp:181-192, light.hpp:41-58, rt.hpp:31-42)."""
import numpy as np

VERTEX = np.dtype([("pos", "<f4", 3), ("normal", "<f4", 3), ("color", "<f4", 3),
                   ("texCoord", "<f4", 2), ("padding0", "<f4")])
MATERIAL = np.dtype([("diffuse", "<f4", 4), ("emission", "<f4", 4), ("alpha", "<f4"),
                     ("metallic", "<f4"), ("specular", "<f4"), ("roughness", "<f4"), ("ior", "<f4"),
                     ("transmission", "<f4"), ("diffuseTexIdx", "<i4"), ("metallicTexIdx", "<i4"),
                     ("roughnessTexIdx", "<i4"), ("transmissionTexIdx", "<i4"), ("padding0", "<i4"),
                     ("padding1", "<i4")])
INSTANCE = np.dtype([("transform", "<f4", 16), ("geometryIndex", "<u4"), ("padding", "<u4", 3)])
CAMERA = np.dtype([("view", "<f4", 16), ("projection", "<f4", 16), ("viewInverse", "<f4", 16),
                   ("projectionInverse", "<f4", 16), ("position", "<f4", 4), ("front", "<f4", 4),
                   ("padding", "<f4", 8)])
DIRECTIONAL_LIGHT = np.dtype([("direction", "<f4", 4), ("rgbs", "<f4", 4)])
POINT_LIGHTS = np.dtype([("posr", "<f4", (32, 4)), ("rgbs", "<f4", (32, 4))])
ACTIVE_LIGHTS = np.dtype([("viewMat", "<f4", (8, 16)), ("projMat", "<f4", (8, 16)),
                          ("front", "<f4", (8, 4)), ("rgbs", "<f4", (8, 4)),
                          ("position", "<f4", (8, 4)), ("sftp", "<f4", (8, 4))])
PUSH_CONSTANTS = np.dtype([("clearColor", "<f4", 4), ("frameCount", "<i4"),
                           ("sampleRatePerPixel", "<u4"), ("maxPathDepth", "<u4"),
                           ("useEnvironmentMap", "<u4"), ("russianRoulette", "<u4"),
                           ("russianRouletteMinBounces", "<u4"), ("nextEventEstimation", "<u4"),
                           ("nextEventEstimationMinBounces", "<u4")])
COUNTERS = np.dtype([("paths", "<u8"), ("extensionRays", "<u8"), ("shadowRays", "<u8"),
                     ("extensionHits", "<u8"), ("nodeVisits", "<u8"), ("triangleTests", "<u8"),
                     ("instanceVisits", "<u8"), ("textureFetches", "<u8"), ("kernelLaunches", "<u8"),
                     ("shadowNodeVisits", "<u8"), ("shadowTriangleTests", "<u8"), ("shadowInstanceVisits", "<u8"),
                     ("tlasNodeVisits", "<u8"), ("instanceEntries", "<u8"), ("shadowRaysSkipped", "<u8"), ("reserved", "<u8", 1)])
BVH_STATS = np.dtype([("blasCount", "<u4"), ("instanceCount", "<u4"), ("triangleCount", "<u8"),
                      ("instancedTriangles", "<u8"), ("blasNodeCount", "<u8"),
                      ("tlasNodeCount", "<u8"), ("nodeBytes", "<u4"), ("triangleBytes", "<u4"),
                      ("instanceBytes", "<u4"), ("tlasRebuilds", "<u4"),
                      ("subtreeNodeCount", "<u8"), ("subtreeTriangles", "<u8"), ("subtreeDepth", "<u4"),
                      ("subtreeBuilds", "<u4")])

assert VERTEX.itemsize == 48 and MATERIAL.itemsize == 80 and INSTANCE.itemsize == 80
assert CAMERA.itemsize == 320 and DIRECTIONAL_LIGHT.itemsize == 32
assert POINT_LIGHTS.itemsize == 1024 and ACTIVE_LIGHTS.itemsize == 1536
assert PUSH_CONSTANTS.itemsize == 48 and COUNTERS.itemsize == 128 and BVH_STATS.itemsize == 80

AUX_RGBA32F, AUX_ALBEDO32F, AUX_NORMAL32F, AUX_HIT_IDS, AUX_HIT_T = 0, 1, 2, 3, 4
AUX_DEPTH, AUX_SEGMENTATION, AUX_SUM32F, AUX_BGRA8 = 5, 6, 7, 8


def push_constants(clear_color=(0, 0, 0, 1), frame_count=0, spp=1, max_depth=8, use_env=False,
                   rr=False, rr_min=4):
    pc = np.zeros((), PUSH_CONSTANTS)
    pc["clearColor"] = clear_color
    pc["frameCount"] = frame_count
    pc["sampleRatePerPixel"] = spp
    pc["maxPathDepth"] = max_depth
    pc["useEnvironmentMap"] = int(use_env)
    pc["russianRoulette"] = int(rr)
    pc["russianRouletteMinBounces"] = rr_min
    pc["nextEventEstimation"] = 1
    return pc

STAGE_TIMES = np.dtype([("ms", "<f8", 8), ("launches", "<u8", 8)])
STAGE_NAMES = ("raygen", "trace_closest", "shade", "trace_occlusion", "shadow_resolve", "finish", "other")
