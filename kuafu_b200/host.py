"""ctypes binding of the C++ host facade through its flat C view (include/kuafu_c.h, libkuafu.so).

The facade is the drop-in `kuafu.hpp` surface; Python only uses it to build the BASELINE scenes with
the real facade code, to drive `Kuafu::run()` / `downloadLatestFrame()`, and to read back the packed
wire buffers (so the CPU oracle in tests/ receives the same bits the device does).
"""
import ctypes as C
import os

import numpy as np

from . import wire
from .build import lib_path

_lib = None
(WIRE_MATERIALS, WIRE_INSTANCES, WIRE_DIRECTIONAL, WIRE_POINTS, WIRE_ACTIVES, WIRE_CAMERA, WIRE_PUSH,
 WIRE_TEXTURE, WIRE_ENV_FACE) = range(9)


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path("libkuafu.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m kuafu_b200.build`")
    lib = C.CDLL(path)
    vp, u32, i32, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t
    lib.kfcCreate.argtypes = [i32, i32]
    lib.kfcCreate.restype = vp
    lib.kfcDestroy.argtypes = [vp]
    lib.kfcDestroy.restype = None
    lib.kfcLastError.restype = C.c_char_p
    lib.kfcLoadScene.argtypes = [vp, C.c_char_p, i32, i32, i32, i32, i32]
    lib.kfcAnimate.argtypes = [vp, i32]
    lib.kfcNumCameras.argtypes = [vp]
    lib.kfcSetCamera.argtypes = [vp, i32]
    lib.kfcRun.argtypes = [vp]
    lib.kfcRunAll.argtypes = [vp]
    lib.kfcCameraShard.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.kfcRunRange.argtypes = [vp, i32, i32]
    lib.kfcRunShard.argtypes = [vp, i32, i32, i32, C.POINTER(i32), i32, C.POINTER(i32)]
    lib.kfcSetEnvironmentMap.argtypes = [vp, C.c_char_p]
    lib.kfcReadTexture.argtypes = [C.c_char_p, C.POINTER(u32), C.POINTER(u32), vp, sz]
    lib.kfcReadKtxCube.argtypes = [C.c_char_p, C.POINTER(u32), vp, sz]
    lib.kfcSetSampleShard.argtypes = [vp, u32, u32, i32]
    lib.kfcResolve.argtypes = [vp]
    lib.kfcDownloadFrame.argtypes = [vp, i32, vp, sz]
    lib.kfcDownloadFrameByValue.argtypes = [vp, i32, vp, sz]
    lib.kfcDownloadAux.argtypes = [vp, i32, i32, vp, sz]
    lib.kfcClockBase.argtypes = [vp]
    lib.kfcClockBase.restype = u32
    lib.kfcSetClockBase.argtypes = [vp, u32]
    lib.kfcFrameCount.argtypes = []
    lib.kfcDeviceContext.argtypes = [vp]
    lib.kfcDeviceContext.restype = vp
    lib.kfcPack.argtypes = [vp]
    lib.kfcWireCounts.argtypes = [vp, C.POINTER(u32)]
    lib.kfcWireGeometry.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u32), C.POINTER(vp), C.POINTER(u32),
                                    C.POINTER(vp), C.POINTER(u32), C.POINTER(i32), C.POINTER(i32)]
    lib.kfcWireBuffer.argtypes = [vp, i32, u32, C.POINTER(vp), C.POINTER(sz)]
    lib.kfcTextureDims.argtypes = [vp, u32, C.POINTER(u32), C.POINTER(u32)]
    _lib = lib
    return lib


def _copy(ptr, nbytes, dtype):
    if not nbytes:
        return np.zeros(0, dtype)
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def camera_shard(n_cameras, rank, world):
    """Kuafu::cameraShard: the contiguous camera range [begin, end) of `rank` among `world` processes."""
    lib = load()
    b, e = C.c_int(), C.c_int()
    if lib.kfcCameraShard(n_cameras, rank, world, C.byref(b), C.byref(e)):
        raise RuntimeError("kfcCameraShard: " + lib.kfcLastError().decode())
    return b.value, e.value


def read_texture(path):
    """The facade's texture-file reader (PNG / PNM): (h, w, 4) uint8 RGBA."""
    lib = load()
    w, h = C.c_uint32(), C.c_uint32()
    if lib.kfcReadTexture(str(path).encode(), C.byref(w), C.byref(h), None, 0):
        raise RuntimeError("kfcReadTexture: " + lib.kfcLastError().decode())
    out = np.empty((h.value, w.value, 4), "u1")
    if lib.kfcReadTexture(str(path).encode(), C.byref(w), C.byref(h), out.ctypes.data_as(C.c_void_p), out.nbytes):
        raise RuntimeError("kfcReadTexture: " + lib.kfcLastError().decode())
    return out


def read_ktx_cube(path):
    """The facade's KTX1 cube-map reader: (6, size, size, 4) uint8 RGBA."""
    lib = load()
    s = C.c_uint32()
    if lib.kfcReadKtxCube(str(path).encode(), C.byref(s), None, 0):
        raise RuntimeError("kfcReadKtxCube: " + lib.kfcLastError().decode())
    out = np.empty((6, s.value, s.value, 4), "u1")
    if lib.kfcReadKtxCube(str(path).encode(), C.byref(s), out.ctypes.data_as(C.c_void_p), out.nbytes):
        raise RuntimeError("kfcReadKtxCube: " + lib.kfcLastError().decode())
    return out


class WireSceneView:
    """Packed wire buffers of a facade scene; `.upload(ctx)` feeds a kfrt Context (the test-side oracle
    reads the same view through `oracle.Oracle.load`)."""

    def __init__(self):
        self.geoms, self.textures, self.env = [], [], None
        self.mats = self.insts = self.dl = self.pl = self.al = self.pc = None
        self.cams = []
        self.w = self.h = 0

    def upload(self, ctx):
        """Feeds a kuafu_b200.rt.Context (the C ABI) in the facade's own call order."""
        ctx.set_limits(1025, 8192, 257, 4097)  # Config limits of the facade renderer that packed this scene
        for gi, (v, idx, mi, op, hide) in enumerate(self.geoms):
            ctx.upload_geometry(gi, v, idx, mi, op, hide)
        ctx.upload_materials(self.mats)
        for ti, t in enumerate(self.textures):
            ctx.upload_texture(ti, t)
        if self.env is not None:
            ctx.set_environment_cube(self.env)
        ctx.set_lights(self.dl, self.pl, self.al)
        ctx.build_blas()
        ctx.set_instances(self.insts)
        ctx.build_tlas()

    def n_tris(self):
        return int(sum(self.geoms[int(i["geometryIndex"])][1].size // 3 for i in self.insts))


class Renderer:
    """kuafu::Kuafu behind the C view.  device=None builds a host-only renderer (no GPU needed)."""

    def __init__(self, device=0, accumulate=False):
        self.lib = load()
        self.h = self.lib.kfcCreate(-1 if device is None else int(device), int(accumulate))
        if not self.h:
            raise RuntimeError("kfcCreate: " + self.lib.kfcLastError().decode())

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError(f"{what}: {self.lib.kfcLastError().decode()}")

    def close(self):
        if self.h:
            self.lib.kfcDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_scene(self, name, width=0, height=0, spp=0, depth=0, scale=0):
        self._ck(self.lib.kfcLoadScene(self.h, name.encode(), width, height, spp, depth, scale), "kfcLoadScene")
        return self.lib.kfcNumCameras(self.h)

    def animate(self, frame):
        self._ck(self.lib.kfcAnimate(self.h, frame), "kfcAnimate")

    def set_camera(self, cam):
        self._ck(self.lib.kfcSetCamera(self.h, cam), "kfcSetCamera")

    def run(self):
        self._ck(self.lib.kfcRun(self.h), "kfcRun")

    def run_all(self):
        self._ck(self.lib.kfcRunAll(self.h), "kfcRunAll")

    def run_range(self, begin, end):
        """Kuafu::run() on the recipe cameras [begin, end) in one launch (camera-batch shard)."""
        self._ck(self.lib.kfcRunRange(self.h, begin, end), "kfcRunRange")

    def run_shard(self, rank, world, interleaved=True):
        """Kuafu::run() on this rank's share of the recipe cameras (Kuafu::cameraShardIndices); returns the
        camera indices rendered, in device slot order."""
        cap = self.lib.kfcNumCameras(self.h)
        idx = (C.c_int * max(cap, 1))()
        n = C.c_int()
        self._ck(self.lib.kfcRunShard(self.h, rank, world, int(interleaved), idx, cap, C.byref(n)), "kfcRunShard")
        return [int(idx[k]) for k in range(n.value)]

    def set_environment_map(self, path):
        self._ck(self.lib.kfcSetEnvironmentMap(self.h, str(path).encode()), "kfcSetEnvironmentMap")

    def set_sample_shard(self, begin, end, defer_resolve=True):
        self._ck(self.lib.kfcSetSampleShard(self.h, begin, end, int(defer_resolve)), "kfcSetSampleShard")

    def resolve(self):
        self._ck(self.lib.kfcResolve(self.h), "kfcResolve")

    @property
    def clock_base(self):
        return self.lib.kfcClockBase(self.h)

    @clock_base.setter
    def clock_base(self, v):
        self.lib.kfcSetClockBase(self.h, v & 0xFFFFFFFF)

    def frame_count(self):
        return self.lib.kfcFrameCount()

    def device_context(self):
        return self.lib.kfcDeviceContext(self.h)

    def counts(self):
        out = (C.c_uint32 * 8)()
        self._ck(self.lib.kfcWireCounts(self.h, out), "kfcWireCounts")
        return dict(zip(("geometries", "materials", "textures", "instances", "cameras", "envSize", "width",
                         "height"), [int(x) for x in out]))

    def download_frame(self, cam=0, out=None, by_value=False):
        """Camera::downloadLatestFrameInto a numpy array (`out` to reuse one); by_value: through
        Kuafu::downloadLatestFrame, the reference's signature (a std::vector, then copied)."""
        c = self.counts()
        if out is None:
            out = np.empty((c["height"], c["width"], 4), "u1")
        fn = self.lib.kfcDownloadFrameByValue if by_value else self.lib.kfcDownloadFrame
        self._ck(fn(self.h, cam, out.ctypes.data_as(C.c_void_p), out.nbytes), "kfcDownloadFrame")
        return out

    def download_aux(self, kind, cam=0):
        from .rt import _AUX
        c = self.counts()
        dt, k = _AUX[kind]
        out = np.empty((c["height"], c["width"], k) if k > 1 else (c["height"], c["width"]), dt)
        self._ck(self.lib.kfcDownloadAux(self.h, cam, kind, out.ctypes.data_as(C.c_void_p), out.nbytes), "kfcDownloadAux")
        return out

    def _buffer(self, kind, index, dtype):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.lib.kfcWireBuffer(self.h, kind, index, C.byref(p), C.byref(n)), "kfcWireBuffer")
        return _copy(p.value, n.value, dtype)

    def wire_scene(self):
        """Pack on the host and copy every wire buffer out."""
        self._ck(self.lib.kfcPack(self.h), "kfcPack")
        c = self.counts()
        ws = WireSceneView()
        ws.w, ws.h = c["width"], c["height"]
        for gi in range(c["geometries"]):
            v, i, m = C.c_void_p(), C.c_void_p(), C.c_void_p()
            nv, ni, nm = C.c_uint32(), C.c_uint32(), C.c_uint32()
            op, hide = C.c_int(), C.c_int()
            self._ck(self.lib.kfcWireGeometry(self.h, gi, C.byref(v), C.byref(nv), C.byref(i), C.byref(ni),
                                              C.byref(m), C.byref(nm), C.byref(op), C.byref(hide)), "kfcWireGeometry")
            ws.geoms.append((_copy(v.value, nv.value * 48, wire.VERTEX), _copy(i.value, ni.value * 4, "<u4"),
                             _copy(m.value, nm.value * 4, "<u4"), bool(op.value), bool(hide.value)))
        ws.mats = self._buffer(WIRE_MATERIALS, 0, wire.MATERIAL)
        ws.insts = self._buffer(WIRE_INSTANCES, 0, wire.INSTANCE)
        ws.dl = self._buffer(WIRE_DIRECTIONAL, 0, wire.DIRECTIONAL_LIGHT)[0]
        ws.pl = self._buffer(WIRE_POINTS, 0, wire.POINT_LIGHTS)[0]
        ws.al = self._buffer(WIRE_ACTIVES, 0, wire.ACTIVE_LIGHTS)[0]
        ws.pc = self._buffer(WIRE_PUSH, 0, wire.PUSH_CONSTANTS)[0]
        ws.cams = [self._buffer(WIRE_CAMERA, k, wire.CAMERA)[0] for k in range(c["cameras"])]
        for ti in range(c["textures"]):
            w, h = C.c_uint32(), C.c_uint32()
            self._ck(self.lib.kfcTextureDims(self.h, ti, C.byref(w), C.byref(h)), "kfcTextureDims")
            ws.textures.append(self._buffer(WIRE_TEXTURE, ti, "u1").reshape(h.value, w.value, 4))
        if c["envSize"]:
            s = c["envSize"]
            ws.env = [self._buffer(WIRE_ENV_FACE, f, "u1").reshape(s, s, 4) for f in range(6)]
        return ws
