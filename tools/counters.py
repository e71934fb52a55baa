#!/usr/bin/env python
"""Per-ray traversal statistics of a scene recipe (detail counters), split closest-hit / occlusion.
usage: python tools/counters.py [scene] [width height spp]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kuafu_b200 import host, rt, wire

scene = sys.argv[1] if len(sys.argv) > 1 else "million"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = int(sys.argv[3]) if len(sys.argv) > 3 else 0
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 4
r = host.Renderer(device=0, accumulate=False)
r.load_scene(scene, w, h, spp)
ctx = rt.Context(handle=r.device_context())
r.run()
ctx.set_detail_counters(True)
r.run()
c = ctx.counters()
ctx.set_detail_counters(False)
ext, sh = int(c["extensionRays"]), int(c["shadowRays"])
sn, st, si = int(c["shadowNodeVisits"]), int(c["shadowTriangleTests"]), int(c["shadowInstanceVisits"])
n, t, i = int(c["nodeVisits"]) - sn, int(c["triangleTests"]) - st, int(c["instanceVisits"]) - si
print(f"{scene}: paths {int(c['paths'])}, extension rays {ext} ({int(c['extensionHits'])/max(ext,1):.2f} hit), shadow rays {sh}")
print(f"  closest-hit per ray: nodes {n/max(ext,1):.2f} tris {t/max(ext,1):.2f} instances {i/max(ext,1):.2f}")
print(f"  both kinds  per ray: top-level nodes {int(c['tlasNodeVisits'])/max(ext+sh,1):.2f}, instance entries {int(c['instanceEntries'])/max(ext+sh,1):.2f} of {(i+si)/max(ext+sh,1):.2f} visits")
print(f"  occlusion   per ray: nodes {sn/max(sh,1):.2f} tris {st/max(sh,1):.2f} instances {si/max(sh,1):.2f}")
ts = []
ctx.set_stage_timers(True)
acc = {}
for k in range(5):
    t0 = time.perf_counter(); r.run(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    for n, (ms, _) in ctx.stage_times().items():
        acc.setdefault(n, []).append(ms)
print(f"  frame {min(ts)*1e3:.2f} ms -> {(ext+sh)/min(ts)/1e6:.1f} Mrays/s (wall, incl. resolve)")
print("  stages (min of 5, ms): " + ", ".join(f"{n} {min(v):.2f}" for n, v in acc.items() if min(v) > 0))
