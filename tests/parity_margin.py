#!/usr/bin/env python
"""Radiance parity margins of the five configs at the test sizes: how far the CUDA path is from the
tolerance written in tests/test_gpu_configs.py (frac_gt_1e-3 < 0.02, mean_rel_diff < 2e-3)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
from kuafu_b200 import host, rt
from oracle import oracle
CASES = [("spheres", 200, 150, 4, 0), ("cornell", 128, 128, 8, 0), ("million", 240, 135, 2, 0),
         ("active", 160, 90, 4, 0), ("articulated", 96, 96, 2, 8)]
for name, w, h, spp, scale in CASES:
    r = host.Renderer(device=0, accumulate=False)
    r.load_scene(name, w, h, spp, 0, scale)
    ws = r.wire_scene(); ws.cams = ws.cams[:1]
    ctx, orc = rt.Context(0), oracle.Oracle()
    ws.upload(ctx); orc.load(ws)
    got, ref = parity.render_both(ws, ctx, orc, clock_base=0)
    parity.assert_hits_bit_exact(got, ref)
    st = parity.radiance_stats(got["sum"], ref["sum"], spp)
    print(f"{name:12s} " + " ".join(f"{k}={v:.3g}" for k, v in st.items()), flush=True)
    ctx.close(); r.close()
