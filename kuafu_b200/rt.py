"""ctypes binding of the C ABI in include/kf_rt.h (libkfrt.so).

This is what a maintainer's FFI stub would look like; the tests and bench.py
drive the CUDA core through it with host buffers, exactly like the C++ facade does.  There is no
fallback: if the library or a B200 is missing every call raises.
"""
import ctypes as C
import os

import numpy as np

from . import wire
from .build import lib_path

_lib = None


class KfrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"kfrt error {code}: {msg}")
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path("libkfrt.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m kuafu_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t
    sig = {
        "kfrtCreate": [i32, C.POINTER(vp)], "kfrtDestroy": [vp], "kfrtSetStream": [vp, vp],
        "kfrtSynchronize": [vp], "kfrtSetLimits": [vp, u32, u32, u32, u32],
        "kfrtUploadGeometry": [vp, u32, vp, u32, vp, u32, vp, u32, i32, i32],
        "kfrtClearGeometries": [vp], "kfrtUploadMaterials": [vp, vp, u32],
        "kfrtUploadTexture": [vp, u32, vp, u32, u32],
        "kfrtSetEnvironmentCube": [vp, C.POINTER(vp), u32], "kfrtClearEnvironment": [vp],
        "kfrtSetLights": [vp, vp, vp, vp], "kfrtBuildBlas": [vp], "kfrtSetInstances": [vp, vp, u32],
        "kfrtBuildTlas": [vp], "kfrtRefitTlas": [vp, vp, u32], "kfrtGetBvhStats": [vp, vp],
        "kfrtSetInstanceSubtrees": [vp, i32, C.c_uint64], "kfrtSetLightSampleCulling": [vp, i32], "kfrtSetOwnInstanceSkip": [vp, i32],
        "kfrtRender": [vp, vp, u32, u32, u32, vp, u32, u32, u32], "kfrtResolve": [vp],
        "kfrtReduceNccl": [vp, vp, i32], "kfrtDownloadBGRA8": [vp, u32, vp, sz],
        "kfrtMapBGRA8": [vp, u32, C.POINTER(vp), C.POINTER(sz)],
        "kfrtDownloadAux": [vp, u32, i32, vp, sz],
        "kfrtGetDeviceBuffer": [vp, i32, C.POINTER(vp), C.POINTER(sz)],
        "kfrtSetDetailCounters": [vp, i32], "kfrtGetCounters": [vp, vp],
        "kfrtSetStageTimers": [vp, i32], "kfrtGetStageTimes": [vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.kfrtLastError.argtypes = [vp]
    lib.kfrtLastError.restype = C.c_char_p
    lib.kfrtVersion.argtypes = []
    lib.kfrtVersion.restype = C.c_char_p
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _arr(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


_AUX = {
    wire.AUX_RGBA32F: ("<f4", 4), wire.AUX_ALBEDO32F: ("<f4", 4), wire.AUX_NORMAL32F: ("<f4", 4),
    wire.AUX_HIT_IDS: ("<i4", 2), wire.AUX_HIT_T: ("<f4", 1), wire.AUX_DEPTH: ("<f4", 1),
    wire.AUX_SEGMENTATION: ("<i4", 1), wire.AUX_SUM32F: ("<f4", 4), wire.AUX_BGRA8: ("u1", 4),
}


class Context:
    """One KfrtContext.  Method names follow the C entry points."""

    def __init__(self, device=0, handle=None):
        self.lib = load()
        self.owned = handle is None
        self.shape = None
        if handle is not None:  # borrow the KfrtContext that lives under a facade renderer
            self.h = C.c_void_p(handle)
            return
        h = C.c_void_p()
        rc = self.lib.kfrtCreate(int(device), C.byref(h))
        if rc:
            raise KfrtError(rc, self.lib.kfrtLastError(None).decode())
        self.h = h
        self.shape = None

    def _ck(self, rc):
        if rc:
            raise KfrtError(rc, self.lib.kfrtLastError(self.h).decode())

    def close(self):
        if self.h and self.owned:
            self.lib.kfrtDestroy(self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_handle):
        self._ck(self.lib.kfrtSetStream(self.h, C.c_void_p(stream_handle or 0)))

    def synchronize(self):
        self._ck(self.lib.kfrtSynchronize(self.h))

    def set_limits(self, geometry, instances, textures, materials):
        self._ck(self.lib.kfrtSetLimits(self.h, geometry, instances, textures, materials))

    def upload_geometry(self, index, vertices, indices, mat_index, opaque=True, hide=False):
        v = _arr(vertices, wire.VERTEX)
        i = _arr(indices, "<u4").reshape(-1)
        m = _arr(mat_index, "<u4").reshape(-1)
        self._ck(self.lib.kfrtUploadGeometry(self.h, index, _ptr(v), v.size, _ptr(i), i.size, _ptr(m),
                                             m.size, int(opaque), int(hide)))

    def clear_geometries(self):
        self._ck(self.lib.kfrtClearGeometries(self.h))

    def upload_materials(self, mats):
        m = _arr(mats, wire.MATERIAL).reshape(-1)
        self._ck(self.lib.kfrtUploadMaterials(self.h, _ptr(m), m.size))

    def upload_texture(self, index, rgba8):
        t = _arr(rgba8, "u1")
        assert t.ndim == 3 and t.shape[2] == 4
        self._ck(self.lib.kfrtUploadTexture(self.h, index, _ptr(t), t.shape[1], t.shape[0]))

    def set_environment_cube(self, faces):
        fs = [_arr(f, "u1") for f in faces]
        assert len(fs) == 6
        size = fs[0].shape[0]
        arr = (C.c_void_p * 6)(*[f.ctypes.data for f in fs])
        self._ck(self.lib.kfrtSetEnvironmentCube(self.h, arr, size))

    def clear_environment(self):
        self._ck(self.lib.kfrtClearEnvironment(self.h))

    def set_lights(self, directional=None, points=None, actives=None):
        d = _arr(directional, wire.DIRECTIONAL_LIGHT) if directional is not None else None
        p = _arr(points, wire.POINT_LIGHTS) if points is not None else None
        a = _arr(actives, wire.ACTIVE_LIGHTS) if actives is not None else None
        self._ck(self.lib.kfrtSetLights(self.h, _ptr(d), _ptr(p), _ptr(a)))

    def build_blas(self):
        self._ck(self.lib.kfrtBuildBlas(self.h))

    def set_instances(self, instances):
        i = _arr(instances, wire.INSTANCE).reshape(-1)
        self._ck(self.lib.kfrtSetInstances(self.h, _ptr(i), i.size))

    def build_tlas(self):
        self._ck(self.lib.kfrtBuildTlas(self.h))

    def refit_tlas(self, transforms):
        t = _arr(transforms, "<f4").reshape(-1, 16)
        self._ck(self.lib.kfrtRefitTlas(self.h, _ptr(t), t.shape[0]))

    def set_instance_subtrees(self, mode, max_triangles=0):
        """0: two-level structure only; 1: world-space instance subtrees when the scene fits the budget (default)."""
        self._ck(self.lib.kfrtSetInstanceSubtrees(self.h, int(mode), int(max_triangles)))

    def set_light_sample_culling(self, on):
        self._ck(self.lib.kfrtSetLightSampleCulling(self.h, int(on)))

    def set_own_instance_skip(self, on):
        self._ck(self.lib.kfrtSetOwnInstanceSkip(self.h, int(on)))

    def bvh_stats(self):
        s = np.zeros((), wire.BVH_STATS)
        self._ck(self.lib.kfrtGetBvhStats(self.h, _ptr(s)))
        return s

    def render(self, cameras, width, height, pc, sample_begin=0, sample_end=None, clock_base=0):
        cams = _arr(cameras, wire.CAMERA).reshape(-1)
        pc = _arr(pc, wire.PUSH_CONSTANTS)
        if sample_end is None:
            sample_end = int(pc.reshape(-1)[0]["sampleRatePerPixel"])
        self._ck(self.lib.kfrtRender(self.h, _ptr(cams), cams.size, width, height, _ptr(pc), sample_begin,
                                     sample_end, clock_base & 0xFFFFFFFF))
        self.shape = (cams.size, height, width)

    def resolve(self):
        self._ck(self.lib.kfrtResolve(self.h))

    def reduce_nccl(self, comm, root=-1):
        self._ck(self.lib.kfrtReduceNccl(self.h, C.c_void_p(comm), root))

    def download_bgra8(self, camera=0):
        _, h, w = self.shape
        out = np.empty((h, w, 4), "u1")
        self._ck(self.lib.kfrtDownloadBGRA8(self.h, camera, _ptr(out), out.nbytes))
        return out

    def map_bgra8(self, camera=0):
        """Zero-copy view of the encoded frame in the context's pinned staging buffer (valid until the next
        resolve)."""
        _, h, w = self.shape
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.kfrtMapBGRA8(self.h, camera, C.byref(p), C.byref(n)))
        assert n.value == h * w * 4
        return np.frombuffer((C.c_ubyte * n.value).from_address(p.value), "u1").reshape(h, w, 4)

    def download_aux(self, kind, camera=0, out=None):
        _, h, w = self.shape
        dt, k = _AUX[kind]
        if out is None:
            out = np.empty((h, w, k) if k > 1 else (h, w), dt)
        self._ck(self.lib.kfrtDownloadAux(self.h, camera, kind, _ptr(out), out.nbytes))
        return out

    def device_buffer(self, kind):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.kfrtGetDeviceBuffer(self.h, kind, C.byref(p), C.byref(n)))
        return p.value, n.value

    def set_detail_counters(self, on):
        self._ck(self.lib.kfrtSetDetailCounters(self.h, int(on)))

    def set_stage_timers(self, on):
        self._ck(self.lib.kfrtSetStageTimers(self.h, int(on)))

    def stage_times(self):
        """{stage name: (total device ms, launches)} of the last render (needs set_stage_timers(True))."""
        t = np.zeros((), wire.STAGE_TIMES)
        self._ck(self.lib.kfrtGetStageTimes(self.h, _ptr(t)))
        return {n: (float(t["ms"][i]), int(t["launches"][i])) for i, n in enumerate(wire.STAGE_NAMES)}

    def counters(self):
        c = np.zeros((), wire.COUNTERS)
        self._ck(self.lib.kfrtGetCounters(self.h, _ptr(c)))
        return c
