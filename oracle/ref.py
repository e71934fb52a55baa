"""ctypes binding of oracle/_ref/libkf_ref.so: the reference's own ray-tracing shaders
(/root/reference/resources/shaders), compiled for the CPU by `make -C oracle ref`.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py): tests/, __graft_entry__ and bench.py's
cpu_baseline / --impl reference legs.  The library takes an oracle scene handle: scenes are uploaded
once through oracle.Oracle and rendered either by the hand-written restatement (Oracle.render) or by
the reference's shaders (render below); that pair is what pins the restatement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as _oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libkf_ref.so")
REFERENCE = os.environ.get("KUAFU_REFERENCE", "/root/reference")
_lib = None


def available():
    """True when the library exists or can be built here (the reference tree is present)."""
    return os.path.exists(_PATH) or os.path.isdir(os.path.join(REFERENCE, "resources", "shaders"))


def load():
    global _lib
    if _lib is not None:
        return _lib
    _oracle.load()  # libkf_ref.so resolves kfo_* from libkf_oracle.so
    if not os.path.exists(_PATH):
        subprocess.check_call(["make", "-C", _HERE, "ref", f"REFERENCE={REFERENCE}"])
    lib = C.CDLL(_PATH)
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    lib.kfref_render.argtypes = [vp, vp, u32, u32, u32, vp, u32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.kfref_render.restype = i32
    _lib = lib
    return lib


def render(orc, cameras, width, height, pc, clock_base=0, brute=False, threads=0, image=None):
    """One vkCmdTraceRaysKHR(width, height, 1) per camera over the scene held by `orc` (oracle.Oracle).
    Returns the three storage images of PathTrace.rgen plus the primary-hit buffers and ray counters.
    `image` (optional, camera-major rgba32f) is the accumulation image of the previous frames, read
    when pc.frameCount > 0."""
    lib = load()
    cams = np.ascontiguousarray(cameras).reshape(-1)
    assert cams.dtype.itemsize == 320
    pc = np.ascontiguousarray(pc)
    assert pc.dtype.itemsize == 48
    n = cams.size
    out = {
        "image": np.zeros((n, height, width, 4), "<f4") if image is None else np.ascontiguousarray(image, "<f4").copy(),
        "albedo": np.zeros((n, height, width, 4), "<f4"),
        "normal": np.zeros((n, height, width, 4), "<f4"),
        "hit_ids": np.full((n, height, width, 2), -1, "<i4"),
        "hit_t": np.zeros((n, height, width), "<f4"),
    }
    counters = np.zeros(4, "<u8")
    p = _oracle._p
    rc = lib.kfref_render(orc.h, p(cams), n, width, height, p(pc), clock_base & 0xFFFFFFFF, int(brute),
                          int(threads), p(out["image"]), p(out["albedo"]), p(out["normal"]), p(out["hit_ids"]),
                          p(out["hit_t"]), p(counters))
    if rc:
        raise RuntimeError(f"kfref_render failed with code {rc}")
    out["counters"] = {"paths": int(counters[0]), "extensionRays": int(counters[1]),
                       "shadowRays": int(counters[2]), "extensionHits": int(counters[3])}
    return out
