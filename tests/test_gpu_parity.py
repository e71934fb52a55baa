"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded wire buffers.  Bit-exact for hit indices / primitive ids / t / depth at sample 0;
toleranced for radiance (libm vs CUDA transcendentals; tolerance stated in each test)."""
import numpy as np
import pytest

import parity
import pyscene
from kuafu_b200 import wire

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt(built):
    from kuafu_b200 import rt as _rt
    return _rt


@pytest.fixture(scope="module")
def orc_mod():
    from oracle import oracle
    return oracle


def _pair(sc, rt, orc_mod):
    ctx = rt.Context(0)
    orc = orc_mod.Oracle()
    sc.upload(ctx)
    sc.upload(orc)
    return ctx, orc


# Radiance tolerance: per-pixel relative error |gpu-cpu|/(|cpu|+1e-3) may exceed 1e-3 on at most 2 %
# of pixels (paths whose discrete decisions flip on a 1-ulp sin/cos/pow difference) and the image
# means must agree to 1e-3 relative.
def _check_radiance(got, ref, spp):
    st = parity.radiance_stats(got["sum"], ref["sum"], spp)
    assert st["frac_gt_1e-3"] < 0.02, st
    assert st["mean_rel_diff"] < 1e-3, st
    return st


@pytest.mark.parametrize("lights", ["dir", "point", "active", "dir point active"])
def test_small_scene_hits_and_radiance(rt, orc_mod, lights):
    sc = pyscene.small_scene(seed=2, w=128, h=96, spp=2, depth=6, lights=lights, textures=True)
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    _check_radiance(got, ref, 2)
    # ray counts follow the same stochastic decisions; allow the same 2 % slack
    for k in ("extensionRays", "shadowRays", "extensionHits"):
        assert abs(got["counters"][k] - ref["counters"][k]) <= 0.02 * max(ref["counters"][k], 1), (k, got["counters"], ref["counters"])
    assert got["counters"]["paths"] == ref["counters"]["paths"]
    ctx.close()


@pytest.mark.parametrize("fov_deg,softness", [(150.0, 0.0), (40.0, 0.0), (40.0, 0.3)])
def test_spot_light_without_texture(rt, orc_mod, fov_deg, softness):
    """The reference's "spot" light is an ActiveLight without a projector texture (scene.cpp:407,
    PathTrace.rchit:262-316 with texID < 0): cone test only, no projection, colour = rgb.  The narrow
    cone leaves most of the scene outside it (lights that fail the cone test draw no further random
    numbers), the soft one perturbs the light position first."""
    sc = pyscene.small_scene(seed=2, w=128, h=96, spp=2, depth=6, lights="active", textures=False)
    assert int(sc.al["sftp"][0][2]) == -1
    sc.al["sftp"][0][0] = softness
    sc.al["sftp"][0][1] = np.radians(fov_deg)
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    _check_radiance(got, ref, 2)
    assert ref["counters"]["shadowRays"] > 0
    for k in ("extensionRays", "shadowRays", "extensionHits"):
        assert abs(got["counters"][k] - ref["counters"][k]) <= 0.02 * max(ref["counters"][k], 1), (k, got["counters"], ref["counters"])
    ctx.close()


def test_mapped_frame_equals_the_downloaded_frame(rt, orc_mod):
    """kfrtMapBGRA8 (the pinned staging copy kfrtResolve starts) hands out the bytes kfrtDownloadBGRA8
    copies, for every camera of a batch, and the oracle's encode of the same sums."""
    sc = pyscene.small_scene(seed=12, w=64, h=40, spp=2, depth=3, lights="dir")
    sc.cams = [sc.cams[0], pyscene.camera([-10.0, 3.0, 6.0], [0.7, -0.2, -0.4], [0, 0, 1], sc.w, sc.h)]
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    ctx.resolve()
    rgba = np.zeros_like(got["sum"])
    want = orc.resolve(got["sum"], rgba, 2, -1)
    for cam in range(2):
        mapped = ctx.map_bgra8(cam).copy()
        assert np.array_equal(mapped, ctx.download_bgra8(cam))
        assert np.array_equal(mapped, want[cam])
    ctx.close()


def test_brute_force_oracle_agrees(rt, orc_mod):
    """BVH layout must not change the answer: GPU (8-wide LBVH) == oracle brute force over all triangles."""
    sc = pyscene.small_scene(seed=5, w=64, h=48, spp=1, depth=3, lights="dir", stacks=6, slices=8)
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc, brute=True)
    parity.assert_hits_bit_exact(got, ref)
    ctx.close()


def test_env_emissive_rr(rt, orc_mod):
    sc = pyscene.small_scene(seed=3, w=96, h=64, spp=4, depth=8, lights="dir", env=True, emissive=True, rr=True)
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    _check_radiance(got, ref, 4)
    ctx.close()


def test_resolve_and_bgra_bit_exact(rt, orc_mod):
    """Accumulate + sRGB/BGRA encode of the SAME float sums must be byte-identical."""
    sc = pyscene.small_scene(seed=4, w=64, h=48, spp=2, depth=4, lights="dir")
    ctx, orc = _pair(sc, rt, orc_mod)
    rgba_ref = np.zeros((1, sc.h, sc.w, 4), np.float32)
    for frame in range(3):
        sc.pc["frameCount"] = frame
        ctx.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, clock_base=10 + frame * 3)
        ctx.resolve()
        gsum = ctx.download_aux(wire.AUX_SUM32F)[None]
        bgra_ref = orc.resolve(gsum, rgba_ref, 2, frame)
        assert np.array_equal(ctx.download_bgra8(), bgra_ref[0])
        assert np.array_equal(ctx.download_aux(wire.AUX_RGBA32F).view(np.uint32), rgba_ref[0].view(np.uint32))
    ctx.close()


def test_spp_shards_union_equals_full(rt, orc_mod):
    """spp sharding: samples [0,2) + [2,4) traced separately sum to the 4-spp frame."""
    sc = pyscene.small_scene(seed=6, w=64, h=48, spp=4, depth=4, lights="dir")
    ctx, orc = _pair(sc, rt, orc_mod)
    cams = np.array(sc.cams, wire.CAMERA)
    ctx.render(cams, sc.w, sc.h, sc.pc, 0, 4, 21)
    full = ctx.download_aux(wire.AUX_SUM32F).astype(np.float64)
    ctx.render(cams, sc.w, sc.h, sc.pc, 0, 2, 21)
    a = ctx.download_aux(wire.AUX_SUM32F).astype(np.float64)
    ctx.render(cams, sc.w, sc.h, sc.pc, 2, 4, 21)
    b = ctx.download_aux(wire.AUX_SUM32F).astype(np.float64)
    assert np.allclose(a[..., :3] + b[..., :3], full[..., :3], rtol=1e-5, atol=1e-6)
    ctx.close()


def test_refit_equals_rebuild(rt, orc_mod):
    """Moving actors: refit of the top level must give the same hits as a fresh build (and the oracle)."""
    sc = pyscene.small_scene(seed=7, w=96, h=64, spp=1, depth=2, lights="dir")
    ctx, orc = _pair(sc, rt, orc_mod)
    insts = np.array(sc.insts, wire.INSTANCE)
    rng = np.random.default_rng(0)
    for step in range(3):
        tr = insts["transform"].copy().reshape(-1, 16)
        tr[1:, 12:15] += rng.normal(scale=0.7, size=(tr.shape[0] - 1, 3)).astype(np.float32)
        ctx.refit_tlas(tr)
        orc.set_transforms(tr)
        got, ref = parity.render_both(sc, ctx, orc, clock_base=step)
        parity.assert_hits_bit_exact(got, ref)
        insts["transform"] = tr
        ctx2 = rt.Context(0)
        sc2 = pyscene.small_scene(seed=7, w=96, h=64, spp=1, depth=2, lights="dir")
        sc2.insts = list(insts)
        sc2.upload(ctx2)
        ctx2.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, clock_base=step)
        assert np.array_equal(ctx2.download_aux(wire.AUX_HIT_IDS), got["hit_ids"][0])
        ctx2.close()
    ctx.close()


def test_error_behaviour(rt):
    ctx = rt.Context(0)
    with pytest.raises(rt.KfrtError) as e:
        ctx.render(np.zeros(1, wire.CAMERA), 8, 8, wire.push_constants())
    assert e.value.code == 4  # KFRT_ERR_NOT_BUILT
    ctx.set_limits(16, 2, 4, 4)
    v, i = pyscene.quad()
    ctx.upload_geometry(0, v, i, np.zeros(i.size, np.uint32))
    with pytest.raises(rt.KfrtError) as e:
        ctx.upload_geometry(16, v, i, np.zeros(i.size, np.uint32))
    assert e.value.code == 3  # KFRT_ERR_LIMIT
    with pytest.raises(rt.KfrtError) as e:
        ctx.set_instances(np.zeros(3, wire.INSTANCE))
    assert e.value.code == 3
    ctx.close()


@pytest.mark.parametrize("keep", [1, 2, 3])
def test_tiny_top_level(rt, orc_mod, keep):
    """1-3 instances: a root with a single InstNode child (k_single_instance_root) / the smallest LBVHs."""
    sc = pyscene.small_scene(seed=8, w=96, h=64, spp=2, depth=5, lights="dir point", textures=True)
    sc.insts = sc.insts[1:1 + keep]
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    _check_radiance(got, ref, 2)
    ctx.close()


def test_small_batches_equal_one_batch(rt, orc_mod, monkeypatch):
    """Wavefront batching must not change the result: sample order of the per-pixel sum is preserved."""
    sc = pyscene.small_scene(seed=9, w=64, h=40, spp=5, depth=4, lights="dir")
    ctx, orc = _pair(sc, rt, orc_mod)
    cams = np.array(sc.cams, wire.CAMERA)
    ctx.render(cams, sc.w, sc.h, sc.pc, clock_base=2)
    one = ctx.download_aux(wire.AUX_SUM32F).copy()
    ctx.close()
    monkeypatch.setenv("KFRT_BATCH_SLOTS", str(64 * 40 * 2))  # two samples per batch -> 3 batches
    ctx2 = rt.Context(0)
    sc.upload(ctx2)
    ctx2.render(cams, sc.w, sc.h, sc.pc, clock_base=2)
    assert np.array_equal(ctx2.download_aux(wire.AUX_SUM32F).view(np.uint32), one.view(np.uint32))
    ctx2.close()


def test_frame_into_a_caller_buffer_equals_the_by_value_frame(built):
    """Camera::downloadLatestFrameInto (additive) and Kuafu::downloadLatestFrame (the reference's by-value
    signature) hand out the same bytes, also for a frame that was displaced by a later render (stash)."""
    from kuafu_b200 import host
    r = host.Renderer(device=0, accumulate=False)
    r.load_scene("active", 96, 54, 1)
    r.set_camera(0)
    r.run()
    a_into = r.download_frame(0).copy()
    a_val = r.download_frame(0, by_value=True).copy()
    assert np.array_equal(a_into, a_val) and int(a_into[..., 3].min()) == 255
    r.set_camera(1)
    r.run()  # camera 0's frame now lives in its stash
    assert np.array_equal(r.download_frame(0), a_into)
    assert np.array_equal(r.download_frame(0, by_value=True), a_into)
    reuse = np.zeros_like(a_into)
    assert r.download_frame(1, out=reuse) is reuse and not np.array_equal(reuse, a_into)
    r.close()
