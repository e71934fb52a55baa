import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from kuafu_b200 import host, rt, wire
def stats(r, ctx):
    ctx.set_detail_counters(True); r.run(); c = ctx.counters(); ctx.set_detail_counters(False)
    rays = int(c["extensionRays"]) + int(c["shadowRays"])
    return int(c["tlasNodeVisits"]) / rays, int(c["nodeVisits"]) / rays
r = host.Renderer(device=0, accumulate=False)
r.load_scene("articulated", 0, 0, 4)
ctx = rt.Context(handle=r.device_context())
r.run()
print("frame 0 (fresh build): tlas/ray %.2f nodes/ray %.2f" % stats(r, ctx))
for f in range(1, 161):
    r.animate(f); r.run()
    if f not in (10, 40, 160): continue
    print("frame %d (after refits): tlas/ray %.2f nodes/ray %.2f, rebuilds so far %d" % ((f,) + stats(r, ctx) + (int(ctx.bvh_stats()["tlasRebuilds"]),)))
r2 = host.Renderer(device=0, accumulate=False)
r2.load_scene("articulated", 0, 0, 4); r2.animate(160)
ctx2 = rt.Context(handle=r2.device_context()); r2.run()
print("frame 160 (fresh build): tlas/ray %.2f nodes/ray %.2f" % stats(r2, ctx2))
