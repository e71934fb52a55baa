#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Kuafu path-tracing core.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode spp|cameras]

Metric (BASELINE.json): Mrays/s and ms/frame at 1080p, 64 spp on config 3 (the ~1 M-triangle
instanced PrincipledBSDF scene with textures and an environment cube, path depth 8, Russian roulette
on).  A "step" is one full frame.  A ray is one traceRayEXT equivalent: every extension ray plus every
shadow ray actually traced, counted on the device.

N > 1 (one process per GPU under torchrun): the 64 samples of every pixel are split across ranks
(replicated scene + BVH), the float4 sample sums are reduced onto rank 0 over NCCL and rank 0 runs the
accumulate + encode epilogue -- strong scaling of one frame.  The reduce is the C ABI's
own collective (kfrtReduceNccl on a communicator from ncclCommInitRank), and after the timed steps rank 0
renders the same frame alone and checks the sharded one against it (`frame_check`, at most 1 BGRA8 LSB).

`--mode cameras` is the second natural shard (BASELINE config 5): the 64 cameras of the
articulated scene are split across ranks (Kuafu::cameraShard), every rank refits its own top level once per
frame, no collective on the data path; the per-rank BGRA8 frames are gathered onto rank 0 inside the e2e
region.  Its line goes to profiles/, the driver's bench is the default mode.

`value` is timed on the device with everything resident in HBM; `e2e` is the same frame through the
public facade call (Kuafu::run + downloadLatestFrame) with host buffers, H2D/D2H inside the timed
region.  `--impl reference` times the reference's own shaders compiled for the CPU (oracle/_ref, built from
/root/reference by `make -C oracle ref`; the hand-written oracle port when that library is absent) on the
box's host cores -- the reference's Vulkan-RT pipeline itself cannot run on a B200.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = {"name": "million", "width": 1920, "height": 1080, "spp": 64, "depth": 8}
WORKLOAD = ("config 3: 204 x createSphere(4900 tris) + floor = 999602 instanced triangles, 17 textured "
            "PrincipledBSDF materials (37 512^2 textures), 6x256^2 env cube, directional light, "
            "1920x1080, 64 spp, path depth 8, Russian roulette on (min 4 bounces)")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="spp", choices=["spp", "cameras"])
    ap.add_argument("--camera-split", default="interleaved", choices=["interleaved", "contiguous"])
    # debugging overrides (a run that uses them is not a bench value; they are echoed in `config`)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--scene", default="")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-spp", type=int, default=8)
    return ap.parse_args()


def effective_config(args):
    cfg = dict(CONFIG)
    if args.scene:
        cfg["name"] = args.scene
    if args.width:
        cfg["width"] = args.width
    if args.height:
        cfg["height"] = args.height
    if args.spp:
        cfg["spp"] = args.spp
    return cfg


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks and throttle reasons while the timed region runs (nvidia-smi, 200 ms)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def trace_closest_bytes(cnt, stats):
    """Algorithmic bytes of the closest-hit traversal stage (the dominant kernel, k_wf_trace<closest>):
    its share of SURVEY.md §8.5 -- nodes, triangles and instance records it fetched, plus 32 B ray read
    and 16 B hit record write per extension ray."""
    ext = int(cnt["extensionRays"])
    n = int(cnt["nodeVisits"]) - int(cnt["shadowNodeVisits"])
    t = int(cnt["triangleTests"]) - int(cnt["shadowTriangleTests"])
    i = int(cnt["instanceVisits"]) - int(cnt["shadowInstanceVisits"])
    b = int(stats["nodeBytes"]) * n + int(stats["triangleBytes"]) * t + int(stats["instanceBytes"]) * i
    return b + ext * (32 + 16), {"nodes": n / max(ext, 1), "triangles": t / max(ext, 1), "instances": i / max(ext, 1)}


def alu_line(per_ray, bytes_per_launch, launch_ms, stats):
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mhz = float(json.load(f).get("sm_max_mhz", 1965.0))
    except (OSError, ValueError):
        mhz = 1965.0
    peak = sms * 128 * 2 * mhz * 1e6 / 1e12
    bytes_per_ray = (int(stats["nodeBytes"]) * per_ray["nodes"] + int(stats["triangleBytes"]) * per_ray["triangles"] +
                     int(stats["instanceBytes"]) * per_ray["instances"] + 48)
    rays_per_launch = bytes_per_launch / max(bytes_per_ray, 1.0)
    flops = rays_per_launch * (136.0 * per_ray["nodes"] + 45.0 * per_ray["triangles"])
    achieved = flops / (launch_ms * 1e-3) / 1e12 if launch_ms > 0 else 0.0
    return {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": f"{sms} SMs x 128 FP32 lanes x 2 x {mhz:.0f} MHz"}


def algorithmic_bytes(cnt, stats, n_pixels):
    """SURVEY.md §8.5 accounting: bytes a frame must move given what it visited (layout constants from
    kfrtGetBvhStats: 80 B wide node, 48 B triangle, 80 B instance record)."""
    ext, sh, hits = int(cnt["extensionRays"]), int(cnt["shadowRays"]), int(cnt["extensionHits"])
    rays = ext + sh
    b = (int(stats["nodeBytes"]) * int(cnt["nodeVisits"]) + int(stats["triangleBytes"]) * int(cnt["triangleTests"]) +
         int(stats["instanceBytes"]) * int(cnt["instanceVisits"]))
    b += rays * (32 + 16)
    b += hits * 240 + 16 * int(cnt["textureFetches"])
    b += ext * 96
    b += n_pixels * 36
    return b, rays


def cpu_arm():
    """The CPU implementation that is timed beside the kernels: the reference's own shaders compiled for
    the CPU (oracle/_ref/libkf_ref.so, built from /root/reference by `make -C oracle ref`; kind
    "reference") when that library travelled with the snapshot, else the hand-written oracle port (kind
    "port").  Both take traversal, triangle test and texture sampling from libkf_oracle.so."""
    from oracle import oracle, ref
    kind = "reference" if os.path.exists(ref._PATH) else "port"

    def render(orc, cams, w, h, pc, sample_spp, clock, threads):
        pcs = np.array(pc).copy()
        pcs["sampleRatePerPixel"] = sample_spp
        if kind == "reference":
            out = ref.render(orc, cams, w, h, pcs, clock_base=clock, threads=threads)
        else:
            out = orc.render(cams, w, h, pcs, 0, sample_spp, clock_base=clock, threads=threads)
        return out["counters"]["extensionRays"] + out["counters"]["shadowRays"]

    return kind, render


def run_reference(args, cfg, rank):
    """CPU arm: the reference's shaders on all host cores (see cpu_arm), bounded sample."""
    if rank != 0:
        return
    from kuafu_b200 import host
    from oracle import oracle
    r = host.Renderer(device=None)
    r.load_scene(cfg["name"], cfg["width"], cfg["height"], cfg["spp"], cfg["depth"])
    ws = r.wire_scene()
    orc = oracle.Oracle()
    orc.load(ws)
    cores = oracle.hardware_threads()
    sample_spp = max(1, min(args.cpu_sample_spp, cfg["spp"]))
    cams = np.array(ws.cams[:1])

    kind, cpu_render = cpu_arm()

    def step(i):
        t = time.perf_counter()
        rays = cpu_render(orc, cams, ws.w, ws.h, ws.pc, sample_spp, i * (cfg["spp"] + 1), cores)
        return time.perf_counter() - t, rays

    for i in range(args.warmup):
        step(i)
    tot_t, tot_r = 0.0, 0
    for i in range(args.steps):
        dt, rays = step(args.warmup + i)
        tot_t += dt
        tot_r += rays
    mrays = tot_r / tot_t / 1e6
    scale = cfg["spp"] / sample_spp
    sample = (f"{sample_spp} of {cfg['spp']} spp of every pixel per step ({cfg['width']}x{cfg['height']}); "
              f"ms_per_step is the sample time x {scale:g}")
    line = {
        "impl": "reference", "metric": "Mrays/s (path tracing, 1080p 64 spp, config 3)", "value": mrays,
        "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_t / args.steps * 1e3 * scale, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, **cfg},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class DevBytes:
    """__cuda_array_interface__ view of a kfrt device buffer as bytes."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


CAMERA_CONFIG = {"name": "articulated", "width": 512, "height": 512, "spp": 32, "depth": 8}
CAMERA_WORKLOAD = ("config 5: 64 chains x 32 links + floor = 2049 instances (10035202 instanced triangles), "
                   "64 cameras x 512x512, 32 spp, path depth 8, every transform rewritten per frame -> one "
                   "top-level refit per frame per GPU")


def run_cameras(args, rank, world, local_rank):
    """Camera-batch shard (SURVEY.md §8.6 way 2): cameras [b, e) of 64 per rank, replicated scene, one
    refit per frame per rank (reference context.cpp:337-341: one update per run()), no collective on the
    data path.  value: device time of Kuafu::run(range) (refit + trace + resolve), max over ranks; e2e: wall
    clock of animate + run + gather of the BGRA8 frames onto rank 0 + copy to pinned host memory."""
    import torch
    import torch.distributed as dist
    from kuafu_b200 import host, rt, wire

    cfg = dict(CAMERA_CONFIG)
    if args.width:
        cfg["width"] = args.width
    if args.height:
        cfg["height"] = args.height
    if args.spp:
        cfg["spp"] = args.spp
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    renderer = host.Renderer(device=local_rank, accumulate=False)
    ncam = renderer.load_scene(cfg["name"], cfg["width"], cfg["height"], cfg["spp"], cfg["depth"])
    interleaved = args.camera_split == "interleaved"
    b, e = host.camera_shard(ncam, rank, world)  # (the size of this rank's share is the same either way)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(handle=renderer.device_context())
    ctx.set_stream(stream.cuda_stream)
    spp = cfg["spp"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, ev=None):
        renderer.animate(i)
        renderer.clock_base = i * (spp + 1)
        if ev:
            ev[0].record(stream)
        renderer.run_shard(rank, world, interleaved)
        if ev:
            ev[1].record(stream)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for i in range(args.warmup):
        step(i)
        flush.zero_()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    my_rays, launches = 0, 0
    for k in range(args.steps):
        step(args.warmup + k, events[k])
        c = ctx.counters()
        my_rays += int(c["extensionRays"]) + int(c["shadowRays"])
        launches += int(c["kernelLaunches"])
        flush.zero_()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(events[k][0].elapsed_time(events[k][1]) for k in range(args.steps))
    tot = torch.tensor([ms, float(my_rays)], dtype=torch.float64, device="cuda")
    mx, sm = tot.clone(), tot.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    total_ms, rays_total = float(mx[0]), float(sm[1])
    value = rays_total / (total_ms * 1e-3) / 1e6
    stats = ctx.bvh_stats()
    # per-rank device time and rays: the cameras of a contiguous range do not see equal amounts of scene
    per_rank = [tot.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, tot)
    rank_ms = [float(t[0]) / args.steps for t in per_rank]
    rank_rays = [float(t[1]) / args.steps for t in per_rank]

    # e2e: frames of all ranks end up in pinned host memory on rank 0
    frame_bytes = cfg["width"] * cfg["height"] * 4
    per_rank = (e - b) * frame_bytes
    assert world == 1 or all(host.camera_shard(ncam, k, world)[1] - host.camera_shard(ncam, k, world)[0] == e - b
                             for k in range(world)), "camera mode wants equal shards"
    pinned = torch.empty(ncam * frame_bytes, dtype=torch.uint8, pin_memory=True) if rank == 0 else None
    gathered = [torch.empty(per_rank, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 and world > 1 else None
    e2e_t, e2e_rays = 0.0, 0.0
    for k in range(args.steps + 1):
        i = 1000 + k
        barrier()
        t = time.perf_counter()
        renderer.animate(i)
        renderer.clock_base = i * (spp + 1)
        renderer.run_shard(rank, world, interleaved)
        ptr, nbytes = ctx.device_buffer(wire.AUX_BGRA8)
        mine = torch.as_tensor(DevBytes(ptr, nbytes), device="cuda")
        if world > 1:
            dist.gather(mine, gathered, dst=0)
            if rank == 0:
                for g, buf in enumerate(gathered):
                    pinned[g * per_rank:(g + 1) * per_rank].copy_(buf, non_blocking=True)
        else:
            pinned.copy_(mine, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        c = ctx.counters()
        r2 = torch.tensor([dt, float(int(c["extensionRays"]) + int(c["shadowRays"]))], dtype=torch.float64, device="cuda")
        m2, s2 = r2.clone(), r2.clone()
        if world > 1:
            dist.all_reduce(m2, op=dist.ReduceOp.MAX)
            dist.all_reduce(s2, op=dist.ReduceOp.SUM)
        if k > 0:  # the first pass sizes the staging buffers
            e2e_t += float(m2[0])
            e2e_rays += float(s2[1])
    if rank == 0:
        assert int(pinned.view(-1, 4)[:, 3].min()) == 255  # every camera's frame arrived (alpha is 255 everywhere)
        line = {
            "metric": "Mrays/s (path tracing, config 5: 64 cameras x 512x512, 32 spp, per-frame TLAS refit)",
            "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": CAMERA_WORKLOAD, **cfg, "cameras": ncam, "parallelism": f"camera-shard x{world} ({args.camera_split})",
                       "cameras_per_gpu": e - b, "l2": "flushed between timed steps (512 MiB memset)",
                       "triangles_instanced": int(stats["instancedTriangles"]), "instances": int(stats["instanceCount"])},
            "per_gpu_mrays": value / world, "rays_per_step": rays_total / args.steps,
            "tlas_rebuilds_rank0": int(stats["tlasRebuilds"]),
            "rank_ms_per_step": rank_ms, "rank_rays_per_step": rank_rays,
            "e2e": {"value": e2e_rays / e2e_t / 1e6, "unit": "Mrays/s",
                    "h2d_bytes_per_step": 64 * int(stats["instanceCount"]) + (e - b) * 320 + 48 + 2592,
                    "d2h_bytes_per_step": ncam * frame_bytes, "ms_per_step": e2e_t / args.steps * 1e3},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def keep_freed_memory():
    """Host-process allocator policy for the e2e legs: Camera::downloadLatestFrame returns the frame BY VALUE
    (the reference's signature), i.e. a fresh 8 MB std::vector per frame, and with glibc's defaults every other
    such allocation is mmap'ed and page-faulted in again (measured on the box: 5.2 ms vs 1.8 ms per download,
    alternating).  A host that downloads frames in a loop keeps freed memory instead; SAPIEN-side this is one
    mallopt call at start-up."""
    import ctypes
    if os.environ.get("KFB_DEFAULT_MALLOC"):  # A/B switch for the measurement in DESIGN.md §6
        return False
    try:
        libc = ctypes.CDLL("libc.so.6")
        M_TRIM_THRESHOLD, M_MMAP_THRESHOLD = -1, -3
        libc.mallopt(M_MMAP_THRESHOLD, 1 << 30)
        libc.mallopt(M_TRIM_THRESHOLD, 1 << 30)
        return True
    except OSError:
        return False


def main():
    args = parse_args()
    keep_freed_memory()
    cfg = effective_config(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if args.mode == "cameras":
        run_cameras(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from kuafu_b200 import host, nccl, rt, wire

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = nccl.Communicator()  # raw ncclComm_t for kfrtReduceNccl
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    spp = cfg["spp"]
    s0, s1 = rank * spp // world, (rank + 1) * spp // world

    # ---- scene through the facade (public API), device context borrowed for the resident-timed arm
    renderer = host.Renderer(device=local_rank, accumulate=False)
    renderer.load_scene(cfg["name"], cfg["width"], cfg["height"], spp, cfg["depth"])
    ws = renderer.wire_scene()
    if world > 1:
        renderer.set_sample_shard(s0, s1, defer_resolve=True)
    # a dedicated (non-default) stream: kfrt launches on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = rt.Context(handle=renderer.device_context())
    ctx.set_stream(stream.cuda_stream)
    t_first = time.perf_counter()
    renderer.run()  # uploads, builds BLAS/TLAS, renders once
    torch.cuda.synchronize()
    first_frame_ms = (time.perf_counter() - t_first) * 1e3  # geometry + texture upload, BLAS + TLAS build, first frame
    stats = ctx.bvh_stats()
    cams = np.array(ws.cams[:1], wire.CAMERA)
    pc = ws.pc
    n_pixels = cfg["width"] * cfg["height"]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def frame(i, ev=None):
        """One resident step: trace this rank's samples, reduce, accumulate + encode."""
        if ev:
            ev[0].record(stream)
        ctx.render(cams, cfg["width"], cfg["height"], pc, s0, s1, clock_base=i * (spp + 1))
        if ev:
            ev[1].record(stream)
        if world > 1:
            ctx.reduce_nccl(comm.handle, root=0)  # ncclReduce of the float4 sample sums, issued by the C ABI
        if rank == 0:
            ctx.resolve()
        if ev:
            ev[2].record(stream)

    for i in range(args.warmup):
        frame(i)
        flush.zero_()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    rays_total, launches = 0, 0
    counters = []
    stage_ms, stage_launches = {}, {}
    ctx.set_stage_timers(True)  # one event record per stage launch, on the same stream
    for k in range(args.steps):
        frame(args.warmup + k, events[k])
        c = ctx.counters()  # synchronises; outside the event brackets
        counters.append(c)
        for name, (ms, n) in ctx.stage_times().items():
            stage_ms[name] = stage_ms.get(name, 0.0) + ms
            stage_launches[name] = stage_launches.get(name, 0) + n
        launches += int(c["kernelLaunches"]) if rank == 0 else 1
        flush.zero_()  # L2 flush between timed steps
    barrier()
    ctx.set_stage_timers(False)
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [events[k][0].elapsed_time(events[k][2]) for k in range(args.steps)]
    render_ms = [events[k][0].elapsed_time(events[k][1]) for k in range(args.steps)]
    my_rays = sum(int(c["extensionRays"]) + int(c["shadowRays"]) for c in counters)
    tot = torch.tensor([sum(step_ms), float(my_rays), sum(render_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, rays_total, render_total_ms = float(mx[0]), float(sm[1]), float(mx[2])
    else:
        total_ms, rays_total, render_total_ms = float(tot[0]), float(tot[1]), float(tot[2])
    value = rays_total / (total_ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the trace kernel), rank 0: counts from one untimed
    # detail-counter pass with the seeds of the last timed step, duration from the timed steps
    roof = None
    if rank == 0:
        ctx.set_detail_counters(True)
        ctx.render(cams, cfg["width"], cfg["height"], pc, s0, s1, clock_base=(args.warmup + args.steps - 1) * (spp + 1))
        detail = ctx.counters()
        ctx.set_detail_counters(False)
        abytes, rays = algorithmic_bytes(detail, stats, n_pixels)
        peak, peak_src = measured_peaks()
        # dominant kernel: the closest-hit traversal stage.  Bytes from the detail pass (same seeds as
        # the last timed step), launch count and duration from the stage events of the timed steps.
        kname = "trace_closest"
        k_launches = max(stage_launches.get(kname, 0) // args.steps, 1)
        k_ms = stage_ms.get(kname, 0.0) / max(stage_launches.get(kname, 0), 1)
        kbytes, kper = trace_closest_bytes(detail, stats)
        kbytes_per_launch = kbytes / k_launches
        achieved = kbytes_per_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        # second line: what ncu measures as the binding limit of this kernel.  The BVH of the BASELINE
        # scenes lives in L1/L2, so the byte line above is an accounting yardstick; the kernel is bound by
        # instruction issue, and the useful fraction of the SM's SIMT issue capacity is
        # issue-slot utilisation x active lanes per instruction / 32 (profiles/issue.json, from the ncu
        # --set full capture of the same command; see profiles/README.md).
        issue = None
        ipath = os.path.join(ROOT, "profiles", "issue.json")
        if os.path.exists(ipath):
            with open(ipath) as f:
                ij = json.load(f)
            issue = {"bound": "issue", "achieved": ij["issue_active_pct"] / 100.0 * ij["lanes_per_inst"] / 32.0,
                     "peak": 1.0, "unit": "fraction of SIMT lane-issue slots", "issue_active_pct": ij["issue_active_pct"],
                     "lanes_per_inst": ij["lanes_per_inst"], "source": ij.get("source")}
            issue["frac"] = issue["achieved"]
        dur = statistics.mean(render_ms) * 1e-3
        frame_achieved = abytes / dur / 1e9
        total_stage = sum(stage_ms.values()) or 1.0
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_wf_trace<closest hit> (extension-ray traversal stage)",
                "peak_source": peak_src + ", burst copy figure",
                "algorithmic_bytes_per_launch": kbytes_per_launch, "launches_per_step": k_launches,
                "launch_ms": k_ms, "per_ray": kper,
                "share_of_step": stage_ms.get(kname, 0.0) / total_stage,
                # SURVEY.md §8.5 secondary line: algorithmic FP32 work of the traversal (136 flops per
                # node step: 8 boxes x 6 slabs x 2 + min/max/compare; 45 per triangle test) against
                # SMs x 128 lanes x 2 x clock.  Most of a node step is byte unpacking and mask logic,
                # not flops, so this line sits far below the issue-slot utilisation ncu reports.
                "alu": alu_line(kper, kbytes_per_launch, k_ms, stats), "issue": issue,
                "stages_ms_per_step": {n: v / args.steps for n, v in stage_ms.items() if v > 0},
                "frame": {"achieved": frame_achieved, "frac": frame_achieved / peak,
                          "algorithmic_bytes_per_step": abytes, "bytes_per_ray": abytes / max(rays, 1),
                          "per_ray": {"nodes": int(detail["nodeVisits"]) / max(rays, 1),
                                      "triangles": int(detail["triangleTests"]) / max(rays, 1),
                                      "instances": int(detail["instanceVisits"]) / max(rays, 1)},
                          "render_ms": statistics.mean(render_ms)}}

    # ---- e2e: the user-facing calls with host buffers.  Kuafu::run(), then the frame into a host buffer the
    # caller owns (Camera::downloadLatestFrameInto == the C ABI's kfrtDownloadBGRA8); a second loop takes the
    # frame through the reference's by-value signature (Kuafu::downloadLatestFrame: a fresh std::vector per
    # frame, then copied out) and is reported beside it.
    h2d = 320 + 48 + 2592 + 64 * int(stats["instanceCount"])  # camera, push constants, lights, transforms
    d2h = n_pixels * 4
    host_frame = np.empty((cfg["height"], cfg["width"], 4), "u1")
    for i in range(2):
        renderer.clock_base = (1000 + i) * (spp + 1)
        renderer.run()
    barrier()

    def e2e_loop(by_value, clock0):
        tot_t, tot_rays = 0.0, 0.0
        for k in range(args.steps):
            renderer.clock_base = (clock0 + k) * (spp + 1)
            barrier()
            t = time.perf_counter()
            renderer.run()
            if world > 1:
                ctx.reduce_nccl(comm.handle, root=0)
            if rank == 0:
                if world > 1:
                    renderer.resolve()
                frame_bytes = renderer.download_frame(0, out=host_frame, by_value=by_value)
                assert frame_bytes.nbytes == d2h
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            c = ctx.counters()
            r = torch.tensor([dt, float(int(c["extensionRays"]) + int(c["shadowRays"]))], dtype=torch.float64, device="cuda")
            if world > 1:
                m = r.clone()
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
                s = r.clone()
                dist.all_reduce(s, op=dist.ReduceOp.SUM)
                tot_t += float(m[0])
                tot_rays += float(s[1])
            else:
                tot_t += dt
                tot_rays += float(r[1])
        return tot_t, tot_rays

    e2e_t, e2e_rays = e2e_loop(False, 2000)
    e2e_value = e2e_rays / e2e_t / 1e6
    bv_t, bv_rays = e2e_loop(True, 2000)
    assert int(host_frame[..., 3].min()) == 255 or rank != 0

    # ---- N > 1: the sharded, reduced and resolved frame against the same frame rendered by rank 0 alone
    frame_check = None
    if world > 1:
        renderer.clock_base = 3000 * (spp + 1)
        renderer.run()
        ctx.reduce_nccl(comm.handle, root=0)
        if rank == 0:
            renderer.resolve()
            sharded = renderer.download_frame(0).copy()
            renderer.set_sample_shard(0, 0)
            renderer.clock_base = 3000 * (spp + 1)
            renderer.run()
            alone = renderer.download_frame(0)
            diff = np.abs(alone.astype(np.int16) - sharded.astype(np.int16))
            frame_check = {"max_lsb": int(diff.max()), "bytes_differing": float((diff > 0).mean()),
                           "against": "the same 64-spp frame rendered by rank 0 alone"}
            assert frame_check["max_lsb"] <= 1, frame_check
        barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        orc = oracle.Oracle()
        orc.load(ws)
        cores = oracle.hardware_threads()
        sample_spp = max(1, min(args.cpu_sample_spp, spp))
        kind, cpu_render = cpu_arm()
        t = time.perf_counter()
        cr = cpu_render(orc, cams, ws.w, ws.h, pc, sample_spp, 0, cores)
        dt = time.perf_counter() - t
        cpu = {"value": cr / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
               "sample": f"{sample_spp} of {spp} spp of every pixel ({ws.w}x{ws.h}), one pass, {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": "Mrays/s (path tracing, 1080p 64 spp, config 3)", "value": value, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, **cfg, "parallelism": f"spp-shard x{world}",
                       "l2": "flushed between timed steps (512 MiB memset)",
                       "triangles_instanced": int(stats["instancedTriangles"]),
                       "instances": int(stats["instanceCount"])},
            "per_gpu_mrays": value / world,
            "rays_per_step": rays_total / args.steps,
            # light samples whose occlusion ray cannot change colour or RNG state (a transmissive surface seen from
            # inside) are answered in the shade stage; the reference traces them.  They are NOT in rays_per_step or
            # in `value`: only rays that were traversed count (rank 0's share shown)
            "light_samples_not_traced_per_step_rank0": sum(int(c["shadowRaysSkipped"]) for c in counters) / args.steps,
            "first_frame_ms": first_frame_ms, "frame_check": frame_check,
            "traversal_structure": "instance subtrees" if int(stats["subtreeNodeCount"]) else "two-level",
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_t / args.steps * 1e3,
                    "call": "Kuafu::run() + Camera::downloadLatestFrameInto(host buffer) [== kfrtDownloadBGRA8]",
                    "by_value": {"value": bv_rays / bv_t / 1e6, "ms_per_step": bv_t / args.steps * 1e3,
                                 "call": "Kuafu::run() + Kuafu::downloadLatestFrame() (std::vector by value, the reference's signature)"}},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
