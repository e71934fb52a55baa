"""One rank of the N > 1 product path on real GPUs (launched by tests/test_gpu_multi.py under
torch.distributed.run, one process per GPU):

  A. spp shard: this rank traces its samples through the facade, the float4 sample sums are reduced by the
     C ABI's own collective (kfrtReduceNccl on a communicator from ncclCommInitRank), rank 0 resolves and
     downloads -- the frame must equal the frame rank 0 renders alone within 1 BGRA8 LSB (the only
     difference is the float association of the per-rank partial sums);
  B. all-reduce variant (root < 0): every rank ends with the same sums, bit for bit;
  C. camera-batch shard (config 5 shape): Kuafu::cameraShard + Kuafu::run(range) after actor motion, no
     collective on the data path; every rank's frames equal the same cameras of an all-camera launch.
Prints one line `MULTI_GPU_OK ...` on rank 0; any failed assertion makes the launcher exit non-zero."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kuafu_b200 import host, nccl, rt, wire  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = nccl.Communicator()

    # ---- A / B: spp shard + kfrtReduceNccl ---------------------------------------------------------
    w, h, spp = 480, 270, 8
    r = host.Renderer(device=local, accumulate=False)
    r.load_scene("million", w, h, spp, 8)
    ctx = rt.Context(handle=r.device_context())
    s0, s1 = rank * spp // world, (rank + 1) * spp // world
    r.set_sample_shard(s0, s1, defer_resolve=True)
    r.clock_base = 77
    r.run()
    own = r.download_aux(wire.AUX_SUM32F, 0).copy()
    assert (own[..., 3] == s1 - s0).all()
    ctx.reduce_nccl(comm.handle, root=0)
    ctx.synchronize()
    if rank == 0:
        r.resolve()
        sharded = r.download_frame(0).copy()
        summed = r.download_aux(wire.AUX_SUM32F, 0).copy()
        assert (summed[..., 3] == spp).all()
    # all-reduce: same shard again, every rank gets the sums
    r.clock_base = 77
    r.run()
    ctx.reduce_nccl(comm.handle, root=-1)
    ctx.synchronize()
    mine = torch.from_numpy(r.download_aux(wire.AUX_SUM32F, 0).copy()).cuda()
    ref0 = mine.clone()
    dist.broadcast(ref0, 0)
    assert torch.equal(mine.view(torch.int32), ref0.view(torch.int32)), "all-reduced sums differ between ranks"
    worst = 0
    if rank == 0:
        # ncclReduce and ncclAllReduce may add the per-rank partial sums in different orders (they do from
        # three ranks on): equal up to float association, not bit for bit
        assert np.allclose(mine.cpu().numpy()[..., :3], summed[..., :3], rtol=2e-5, atol=1e-6), "reduce and all-reduce disagree on the root"
        r.set_sample_shard(0, 0)
        r.clock_base = 77
        r.run()
        alone = r.download_frame(0)
        worst = int(np.abs(alone.astype(int) - sharded.astype(int)).max())
        assert worst <= 1, f"sharded frame differs from the 1-GPU frame by {worst} LSB"
        assert (alone != sharded).mean() < 0.02
        one = r.download_aux(wire.AUX_SUM32F, 0)
        assert np.allclose(one[..., :3], summed[..., :3], rtol=2e-5, atol=1e-6)
    r.close()

    # ---- C: camera-batch shard ---------------------------------------------------------------------
    r = host.Renderer(device=local, accumulate=False)
    ncam = r.load_scene("articulated", 96, 96, 2, 0, 8)
    b, e = host.camera_shard(ncam, rank, world)
    assert 0 <= b <= e <= ncam
    for frame in range(3):  # one refit per run() on every rank, same transforms everywhere
        r.animate(frame)
        r.clock_base = 200 + frame
        r.run_range(b, e)
    mine = np.stack([r.download_frame(c) for c in range(b, e)]) if e > b else np.zeros((0, 96, 96, 4), "u1")
    r.clock_base = 202
    r.run_all()
    everything = np.stack([r.download_frame(c) for c in range(ncam)])
    assert np.array_equal(mine, everything[b:e]), "a camera rendered in a shard differs from the all-camera launch"
    # the batch as rank 0 would hand it on: gathered per-rank frames == the all-camera launch
    counts = [host.camera_shard(ncam, k, world) for k in range(world)]
    # the interleaved split (rank, rank + world, ...) renders the same frames
    r.clock_base = 202
    idx = r.run_shard(rank, world, interleaved=True)
    assert idx == list(range(rank, ncam, world))
    for c in idx:
        assert np.array_equal(r.download_frame(c), everything[c]), f"camera {c} of the interleaved shard differs"
    if len({ce - cb for cb, ce in counts}) == 1:  # equal shards: one all_gather of the BGRA8 frames
        bufs = [torch.empty((ce - cb, 96, 96, 4), dtype=torch.uint8, device="cuda") for cb, ce in counts]
        dist.all_gather(bufs, torch.from_numpy(mine).cuda())
        if rank == 0:
            assert np.array_equal(torch.cat(bufs).cpu().numpy(), everything)
    r.close()

    ok = torch.ones(1, device="cuda")
    dist.all_reduce(ok)
    comm.close()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} sharded_vs_alone_max_lsb={worst} cameras={ncam}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
