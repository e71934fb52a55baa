"""world_size-2 test of the N > 1 path on CPU (gloo): spp sharding + sum-reduce + resolve on the root
reproduces the single-process frame.  The per-rank renderer is the CPU oracle here (test
infrastructure); on GPUs bench.py runs the same sharding arithmetic with kfrtRender + NCCL."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(spp, rank, world):
    """The sample range rank `rank` of `world` traces (bench.py uses the same expression)."""
    return rank * spp // world, (rank + 1) * spp // world


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pyscene
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = pyscene.small_scene(seed=4, w=32, h=24, spp=6, depth=4)
    orc = oracle.Oracle()
    sc.upload(orc)
    s0, s1 = shard(6, rank, world)
    part = orc.render(np.array(sc.cams), sc.w, sc.h, sc.pc, s0, s1, clock_base=9, threads=1)
    t = torch.from_numpy(part["sum"].copy())
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    cnt = torch.tensor([part["counters"]["extensionRays"] + part["counters"]["shadowRays"]], dtype=torch.int64)
    dist.all_reduce(cnt)
    if rank == 0:
        full = orc.render(np.array(sc.cams), sc.w, sc.h, sc.pc, 0, 6, clock_base=9, threads=1)
        ok_sum = np.allclose(t.numpy()[..., :3], full["sum"][..., :3], rtol=1e-5, atol=1e-6)
        ok_cnt = int(cnt[0]) == full["counters"]["extensionRays"] + full["counters"]["shadowRays"]
        rgba_a, rgba_b = np.zeros_like(full["sum"]), np.zeros_like(full["sum"])
        a = orc.resolve(t.numpy(), rgba_a, 6, 0)
        b = orc.resolve(full["sum"], rgba_b, 6, 0)
        ok_img = (np.abs(a.astype(int) - b.astype(int)) > 1).sum() == 0
        ok_hits = np.array_equal(part["hit_ids"], full["hit_ids"])  # sample 0 lives on rank 0
        q.put((ok_sum, ok_cnt, ok_img, ok_hits))
    dist.destroy_process_group()


def test_shard_ranges_partition():
    for spp in (1, 4, 6, 64):
        for world in (1, 2, 3, 4, 8):
            r = [shard(spp, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))


def _camera_worker(rank, world, port, q):
    """Camera-batch shard on CPU: each rank asks the facade (Kuafu::cameraShard through the C view) for
    its range of the config 5 camera batch and packs exactly those cameras; rank 0 checks that the
    ranges tile the batch and that the packed cameras are the batch's own, in order."""
    sys.path.insert(0, ROOT)
    from kuafu_b200 import host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = host.Renderer(device=None)
    ncam = r.load_scene("articulated", 32, 32, 1, 0, 7)  # 7 cameras: an uneven split
    r.animate(2)
    ws = r.wire_scene()
    b, e = host.camera_shard(ncam, rank, world)
    mine = np.stack([np.frombuffer(np.array(c).tobytes(), "u1") for c in ws.cams[b:e]]) if e > b else np.zeros((0, 320), "u1")
    ranges = [None] * world
    dist.all_gather_object(ranges, (b, e))
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    tr = torch.from_numpy(np.ascontiguousarray(np.array(ws.insts)["transform"], np.float32).copy())
    tr0 = tr.clone()
    dist.broadcast(tr0, 0)
    same_scene = bool(torch.equal(tr, tr0))  # every rank refits to the same transforms
    if rank == 0:
        tiled = ranges[0][0] == 0 and ranges[-1][1] == ncam and all(ranges[k][1] == ranges[k + 1][0] for k in range(world - 1))
        sizes = [hi - lo for lo, hi in ranges]
        even = max(sizes) - min(sizes) <= 1
        whole = np.concatenate(parts)
        allc = np.stack([np.frombuffer(np.array(c).tobytes(), "u1") for c in ws.cams])
        q.put((tiled, even, bool(np.array_equal(whole, allc)), same_scene, ncam))
    dist.barrier()
    r.close()
    dist.destroy_process_group()


def test_gloo_world2_camera_shard():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_camera_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True, True, True, 7), res


def test_camera_shard_ranges_partition():
    from kuafu_b200 import host
    for n in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            r = [host.camera_shard(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    with pytest.raises(RuntimeError):
        host.camera_shard(8, 2, 2)


def test_interleaved_camera_shard_covers_the_batch():
    """Kuafu::cameraShardIndices through a host-only renderer: rank, rank + world, ... ; every camera once."""
    from kuafu_b200 import host
    import ctypes as C
    r = host.Renderer(device=None)
    ncam = r.load_scene("articulated", 32, 32, 1, 0, 7)
    lib = host.load()
    for world in (1, 2, 3):
        seen = []
        for rank in range(world):
            for inter in (0, 1):
                idx = (C.c_int * ncam)()
                n = C.c_int()
                # host-only: the indices come back, then run() refuses to render (no device, no CPU fallback)
                rc = lib.kfcRunShard(r.h, rank, world, inter, idx, ncam, C.byref(n))
                assert rc != 0 and b"no CUDA device" in lib.kfcLastError() or n.value == 0
                got = [idx[k] for k in range(n.value)]
                if inter:
                    assert got == list(range(rank, ncam, world))
                    seen += got
                else:
                    b, e = host.camera_shard(ncam, rank, world)
                    assert got == list(range(b, e))
        assert sorted(seen) == list(range(ncam))
    r.close()


def test_gloo_world2_spp_shard_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True, True, True), res
