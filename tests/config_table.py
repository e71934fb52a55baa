#!/usr/bin/env python
"""Measures the five BASELINE.json configs at full size through the facade (Kuafu::run on all recipe
cameras, stage timers on) and prints one JSON line per config plus a markdown table.
usage: python tests/config_table.py [--cpu] [names...]   (--cpu also times the oracle on a 1-spp sample)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kuafu_b200 import host, rt, wire

CONFIGS = [("spheres", "1 eSpheres 800x600 4 spp d8"), ("cornell", "2 Cornell 1024^2 64 spp d12"),
           ("million", "3 1M tris 1080p 64 spp d8"), ("active", "4 eActive stereo 720p 32 spp"),
           ("articulated", "5 64 cams x 512^2, 10M tris, 32 spp, refit")]
args = [a for a in sys.argv[1:] if not a.startswith("--")]
with_cpu = "--cpu" in sys.argv
rows = []
for name, label in CONFIGS:
    if args and name not in args:
        continue
    r = host.Renderer(device=0, accumulate=False)
    ncam = r.load_scene(name)
    ctx = rt.Context(handle=r.device_context())
    t0 = time.perf_counter(); r.run_all(); ctx.synchronize(); t_first = time.perf_counter() - t0  # upload + build + frame
    stats = ctx.bvh_stats()
    ctx.set_stage_timers(True)
    ts, rays = [], 0
    for k in range(3):
        if name == "articulated":
            r.animate(k + 1)  # per-frame TLAS refit inside the timed region
        t0 = time.perf_counter(); r.run_all(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
        c = ctx.counters(); rays = int(c["extensionRays"]) + int(c["shadowRays"])
    st = ctx.stage_times()
    ctx.set_stage_timers(False)
    ms = float(np.median(ts)) * 1e3
    row = {"config": label, "name": name, "cameras": ncam, "triangles_instanced": int(stats["instancedTriangles"]),
           "instances": int(stats["instanceCount"]), "rays_per_frame": rays, "ms_per_frame": ms,
           "mrays_s": rays / ms / 1e3, "first_frame_ms_incl_build": t_first * 1e3,
           "stage_ms": {k: v[0] for k, v in st.items() if v[0] > 0}}
    if with_cpu:
        from oracle import oracle
        ws = r.wire_scene()
        orc = oracle.Oracle(); orc.load(ws)
        cores = oracle.hardware_threads()
        t0 = time.perf_counter()
        out = orc.render(np.array(ws.cams[:1], wire.CAMERA), ws.w, ws.h, ws.pc, 0, 1, clock_base=0, threads=cores)
        dt = time.perf_counter() - t0
        cr = out["counters"]["extensionRays"] + out["counters"]["shadowRays"]
        row["cpu_mrays_s"] = cr / dt / 1e6
        row["cpu_cores"] = cores
        row["cpu_sample"] = "1 spp of camera 0"
    rows.append(row)
    print(json.dumps(row), flush=True)
    r.close()
print("\n| Config | tris (instanced) | rays/frame | B200 x1 ms/frame | B200 x1 Mrays/s | CPU oracle Mrays/s |")
print("|---|---|---|---|---|---|")
for w in rows:
    cpu = f"{w['cpu_mrays_s']:.2f} ({w['cpu_cores']} thr)" if "cpu_mrays_s" in w else "-"
    print(f"| {w['config']} | {w['triangles_instanced']:,} | {w['rays_per_frame']/1e6:.1f} M | {w['ms_per_frame']:.1f} | {w['mrays_s']:.0f} | {cpu} |")
