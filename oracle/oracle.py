"""ctypes binding of the CPU oracle (oracle/libkf_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under kuafu_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "libkf_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", _HERE])
    lib = C.CDLL(path)
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    lib.kfo_create.restype = vp
    lib.kfo_create.argtypes = []
    lib.kfo_destroy.argtypes = [vp]
    lib.kfo_destroy.restype = None
    lib.kfo_tea.argtypes = [u32, u32]
    lib.kfo_tea.restype = u32
    lib.kfo_lcg.argtypes = [C.POINTER(u32)]
    lib.kfo_lcg.restype = u32
    lib.kfo_rnd.argtypes = [C.POINTER(u32)]
    lib.kfo_rnd.restype = C.c_float
    lib.kfo_set_geometry.argtypes = [vp, u32, vp, u32, vp, u32, vp, u32, i32, i32]
    lib.kfo_set_materials.argtypes = [vp, vp, u32]
    lib.kfo_set_texture.argtypes = [vp, u32, vp, u32, u32]
    lib.kfo_set_env_cube.argtypes = [vp, C.POINTER(vp), u32]
    lib.kfo_set_instances.argtypes = [vp, vp, u32]
    lib.kfo_set_transforms.argtypes = [vp, vp, u32]
    lib.kfo_set_lights.argtypes = [vp, vp, vp, vp]
    lib.kfo_set_skip_own_instance.argtypes = [vp, i32]
    lib.kfo_set_cull_light_samples.argtypes = [vp, i32]
    lib.kfo_last_shadow_skipped.argtypes = [vp]
    lib.kfo_last_shadow_skipped.restype = C.c_uint64
    lib.kfo_render.argtypes = [vp, vp, u32, u32, u32, vp, u32, u32, u32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.kfo_resolve.argtypes = [vp, vp, vp, vp, C.c_uint64, u32, C.c_int32]
    lib.kfo_hardware_threads.restype = i32
    for n in ("kfo_set_geometry", "kfo_set_materials", "kfo_set_texture", "kfo_set_env_cube",
              "kfo_set_instances", "kfo_set_transforms", "kfo_set_lights", "kfo_render", "kfo_resolve"):
        getattr(lib, n).restype = i32
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def tea(a, b):
    return load().kfo_tea(a & 0xFFFFFFFF, b & 0xFFFFFFFF)


def lcg_stream(seed, n):
    """(n lcg outputs, n rnd outputs) from two copies of the same seed."""
    lib = load()
    s1, s2 = C.c_uint32(seed), C.c_uint32(seed)
    ints = [lib.kfo_lcg(C.byref(s1)) for _ in range(n)]
    flts = [lib.kfo_rnd(C.byref(s2)) for _ in range(n)]
    return ints, flts


def hardware_threads():
    return load().kfo_hardware_threads()


class Oracle:
    """Scene holder mirroring the order of calls a kfrt context receives (raw wire buffers)."""

    def __init__(self):
        self.lib = load()
        self.h = C.c_void_p(self.lib.kfo_create())

    def close(self):
        if self.h:
            self.lib.kfo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError(f"oracle {what} failed with code {rc}")

    def load(self, view):
        """Takes the packed wire buffers of a scene (kuafu_b200.host.WireSceneView or anything with the
        same attributes): the oracle receives bit for bit what the C ABI receives."""
        for gi, (v, idx, mi, op, hide) in enumerate(view.geoms):
            self.set_geometry(gi, v, idx, mi, op, hide)
        self.set_materials(view.mats)
        for ti, t in enumerate(view.textures):
            self.set_texture(ti, t)
        if view.env is not None:
            self.set_env_cube(view.env)
        self.set_lights(view.dl, view.pl, view.al)
        self.set_instances(view.insts)

    def set_geometry(self, index, vertices, indices, mat_index, opaque=True, hide=False):
        v = np.ascontiguousarray(vertices)
        assert v.dtype.itemsize == 48
        i = np.ascontiguousarray(indices, "<u4").reshape(-1)
        m = np.ascontiguousarray(mat_index, "<u4").reshape(-1)
        self._ck(self.lib.kfo_set_geometry(self.h, index, _p(v), v.size, _p(i), i.size, _p(m), m.size,
                                           int(opaque), int(hide)), "set_geometry")

    def set_materials(self, mats):
        m = np.ascontiguousarray(mats).reshape(-1)
        assert m.dtype.itemsize == 80
        self._ck(self.lib.kfo_set_materials(self.h, _p(m), m.size), "set_materials")

    def set_texture(self, index, rgba8):
        t = np.ascontiguousarray(rgba8, "u1")
        self._ck(self.lib.kfo_set_texture(self.h, index, _p(t), t.shape[1], t.shape[0]), "set_texture")

    def set_env_cube(self, faces):
        fs = [np.ascontiguousarray(f, "u1") for f in faces]
        arr = (C.c_void_p * 6)(*[f.ctypes.data for f in fs])
        self._ck(self.lib.kfo_set_env_cube(self.h, arr, fs[0].shape[0]), "set_env_cube")

    def set_instances(self, instances):
        i = np.ascontiguousarray(instances).reshape(-1)
        assert i.dtype.itemsize == 80
        self._ck(self.lib.kfo_set_instances(self.h, _p(i), i.size), "set_instances")

    def set_transforms(self, transforms):
        t = np.ascontiguousarray(transforms, "<f4").reshape(-1, 16)
        self._ck(self.lib.kfo_set_transforms(self.h, _p(t), t.shape[0]), "set_transforms")

    def set_lights(self, directional=None, points=None, actives=None):
        d = np.ascontiguousarray(directional) if directional is not None else None
        p = np.ascontiguousarray(points) if points is not None else None
        a = np.ascontiguousarray(actives) if actives is not None else None
        self._ck(self.lib.kfo_set_lights(self.h, _p(d), _p(p), _p(a)), "set_lights")

    def set_skip_own_instance(self, on):
        """Mirror of the CUDA path's declared deviation D6 (kf_oracle.cpp header); off by default."""
        self._ck(self.lib.kfo_set_skip_own_instance(self.h, int(on)), "set_skip_own_instance")

    def set_cull_light_samples(self, on):
        """Mirror of the CUDA path's light-sample culling (kf_oracle.cpp, Scene::cullLightSamples); off by default."""
        self._ck(self.lib.kfo_set_cull_light_samples(self.h, int(on)), "set_cull_light_samples")

    def last_shadow_skipped(self):
        return int(self.lib.kfo_last_shadow_skipped(self.h))

    def render(self, cameras, width, height, pc, sample_begin=0, sample_end=None, clock_base=0,
               brute=False, threads=0):
        cams = np.ascontiguousarray(cameras).reshape(-1)
        assert cams.dtype.itemsize == 320
        pc = np.ascontiguousarray(pc)
        assert pc.dtype.itemsize == 48
        if sample_end is None:
            sample_end = int(pc.reshape(-1)[0]["sampleRatePerPixel"])
        n = cams.size
        out = {
            "sum": np.zeros((n, height, width, 4), "<f4"),
            "albedo": np.zeros((n, height, width, 4), "<f4"),
            "normal": np.zeros((n, height, width, 4), "<f4"),
            "hit_ids": np.full((n, height, width, 2), -1, "<i4"),
            "hit_t": np.zeros((n, height, width), "<f4"),
            "depth": np.zeros((n, height, width), "<f4"),
        }
        counters = np.zeros(4, "<u8")
        self._ck(self.lib.kfo_render(self.h, _p(cams), n, width, height, _p(pc), sample_begin, sample_end,
                                     clock_base & 0xFFFFFFFF, int(brute), int(threads), _p(out["sum"]),
                                     _p(out["albedo"]), _p(out["normal"]), _p(out["hit_ids"]),
                                     _p(out["hit_t"]), _p(out["depth"]), _p(counters)), "render")
        out["counters"] = {"paths": int(counters[0]), "extensionRays": int(counters[1]),
                           "shadowRays": int(counters[2]), "extensionHits": int(counters[3])}
        return out

    def resolve(self, sum_, rgba, spp, frame_count):
        """In-place running mean on `rgba`; returns the BGRA8 frame(s)."""
        s = np.ascontiguousarray(sum_, "<f4")
        assert rgba.flags["C_CONTIGUOUS"] and rgba.dtype == np.float32 and rgba.shape == s.shape
        bgra = np.zeros(s.shape[:-1] + (4,), "u1")
        self._ck(self.lib.kfo_resolve(self.h, _p(s), _p(rgba), _p(bgra), s.size // 4, spp, frame_count), "resolve")
        return bgra
