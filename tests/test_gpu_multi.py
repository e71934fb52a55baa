"""N > 1 on real GPUs: the spp-sharded frame through kfrtReduceNccl and the camera-batch shard, each
against the 1-GPU result (tests/multi_gpu_worker.py, one process per GPU).  Skipped on a box with one
GPU; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it.  Also here: a context
that lives on GPU 1 while the calling thread's current device is GPU 0."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_frames_equal_the_single_gpu_frames(built, world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MULTI_GPU_OK" in p.stdout, p.stdout[-2000:]


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_context_on_another_device_and_thread(built):
    """A context created on GPU 1, driven from a thread whose current device is GPU 0 (every entry point
    makes the context's device current), gives the frame the same scene gives on GPU 0."""
    import pyscene
    from kuafu_b200 import rt, wire
    sc = pyscene.small_scene(seed=8, w=64, h=48, spp=2, depth=4, lights="dir point", textures=True)
    frames = {}

    def work(dev, key):
        import torch
        torch.cuda.set_device(0)  # the thread's current device is never the context's own for dev 1
        ctx = rt.Context(dev)
        torch.cuda.set_device(0)
        sc.upload(ctx)
        ctx.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, 0, 2, 5)
        ctx.resolve()
        frames[key] = (ctx.download_bgra8(0).copy(), ctx.download_aux(wire.AUX_HIT_IDS).copy())
        ctx.close()

    work(0, "a")
    t = threading.Thread(target=work, args=(1, "b"))
    t.start()
    t.join()
    assert "b" in frames, "the thread died"
    assert np.array_equal(frames["a"][0], frames["b"][0]) and np.array_equal(frames["a"][1], frames["b"][1])
