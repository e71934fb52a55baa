#!/usr/bin/env python
"""Acceleration-structure build / refit times of the scene recipes (rows a13 / a14).
usage: python tools/build_times.py [names...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kuafu_b200 import host, rt, wire

names = sys.argv[1:] or ["spheres", "cornell", "million", "active", "articulated"]
# one untimed build first: the first launch of every build kernel pays CUDA's lazy module loading
# (tens of ms once per process), which is not build time
names = ["spheres"] + names
for k, name in enumerate(names):
    r = host.Renderer(device=None)
    r.load_scene(name)
    ws = r.wire_scene()
    ctx = rt.Context(0)
    ctx.set_limits(1025, 8192, 257, 4097)
    for gi, (v, idx, mi, op, hide) in enumerate(ws.geoms):
        ctx.upload_geometry(gi, v, idx, mi, op, hide)
    ctx.upload_materials(ws.mats)
    ctx.synchronize()
    def timed(fn, reps=1):
        ts = []
        for _ in range(reps):
            ctx.synchronize(); t0 = time.perf_counter(); fn(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
        return min(ts) * 1e3
    t_blas_first = timed(ctx.build_blas)   # fresh context: includes the allocation of the build scratch
    for gi, (v, idx, mi, op, hide) in enumerate(ws.geoms):
        ctx.upload_geometry(gi, v, idx, mi, op, hide)
    ctx.synchronize()
    t_blas = timed(ctx.build_blas)         # rebuild of every geometry with the scratch in place
    ctx.set_instances(ws.insts)
    t_tlas = timed(ctx.build_tlas, 3)
    tr = np.ascontiguousarray(np.array(ws.insts)["transform"], np.float32)
    t_refit = timed(lambda: ctx.refit_tlas(tr), 5)
    st = ctx.bvh_stats()
    if k == 0:
        ctx.close(); r.close()
        continue
    print(f"{name:12s} tris {int(st['triangleCount']):9d} (instanced {int(st['instancedTriangles']):9d}) "
          f"blas {int(st['blasCount']):3d} nodes {int(st['blasNodeCount']):7d} | build BLAS {t_blas:8.2f} ms "
          f"({int(st['triangleCount'])/t_blas/1e3:7.2f} Mtris/s; first in context {t_blas_first:7.2f} ms) | TLAS {int(st['instanceCount']):5d} inst build {t_tlas:6.2f} ms refit {t_refit:6.3f} ms")
    ctx.close(); r.close()
