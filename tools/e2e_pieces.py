#!/usr/bin/env python
"""Where the end-to-end frame time goes on one GPU: when Kuafu::run() returns (launches are asynchronous),
when downloadLatestFrame() returns, and the device time of the frame itself (config 3, 1080p).
usage: python tools/e2e_pieces.py [spp]"""
import os, sys, time, ctypes
sys.path.insert(0, os.getcwd())
libc = ctypes.CDLL("libc.so.6"); libc.mallopt(-3, 1 << 30); libc.mallopt(-1, 1 << 30)
import numpy as np, torch
from kuafu_b200 import host, rt
r = host.Renderer(device=0)
r.load_scene("million", 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 64)
ctx = rt.Context(handle=r.device_context())
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
for k in range(3):
    r.run(); r.download_frame(0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for k in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ev[0].record(stream)
    r.run()
    t1 = time.perf_counter(); ev[1].record(stream)
    f = r.download_frame(0)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"run() returns after {1e3*(t1-t0):7.2f} ms | download_frame returns after {1e3*(t2-t0):7.2f} ms (took {1e3*(t2-t1):6.2f}) | sync {1e3*(t3-t2):5.2f} | device run {ev[0].elapsed_time(ev[1]):7.2f} ms | total {1e3*(t3-t0):7.2f}")
