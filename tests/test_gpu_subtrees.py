"""World-space instance subtrees (kf_wsi.cuh, kfrtSetInstanceSubtrees) against the two-level structure and the
oracle: the two structures must give the SAME buffers bit for bit -- hit ids, t, depth, and the radiance sums
too, because every path takes the same decisions on the same hits."""
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


def _buffers(ctx, sc, clock_base=3):
    ctx.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, 0, None, clock_base)
    n = len(sc.cams)
    out = {k: np.stack([ctx.download_aux(a, c) for c in range(n)])
           for k, a in (("sum", wire.AUX_SUM32F), ("hit_ids", wire.AUX_HIT_IDS), ("hit_t", wire.AUX_HIT_T),
                        ("depth", wire.AUX_DEPTH), ("albedo", wire.AUX_ALBEDO32F), ("normal", wire.AUX_NORMAL32F))}
    cnt = ctx.counters()
    out["counters"] = {k: int(cnt[k]) for k in ("paths", "extensionRays", "shadowRays", "extensionHits")}
    return out


def _same(a, b):
    for k in ("hit_ids", "hit_t", "depth", "sum", "albedo", "normal"):
        assert np.array_equal(a[k].view(np.uint32) if a[k].dtype == np.float32 else a[k],
                              b[k].view(np.uint32) if b[k].dtype == np.float32 else b[k]), k
    assert a["counters"] == b["counters"]


@pytest.mark.parametrize("lights,textures", [("dir", True), ("dir point active", True), ("point", False)])
def test_subtrees_equal_two_level_and_oracle(rt, orc_mod, lights, textures):
    sc = pyscene.small_scene(seed=5, w=160, h=120, spp=3, depth=6, lights=lights, textures=textures)
    ctx = rt.Context(0)
    sc.upload(ctx)
    ctx.set_instance_subtrees(0)
    two = _buffers(ctx, sc)
    assert int(ctx.bvh_stats()["subtreeNodeCount"]) == 0
    ctx.set_instance_subtrees(2)
    sub = _buffers(ctx, sc)
    st = ctx.bvh_stats()
    assert int(st["subtreeNodeCount"]) > 0 and int(st["subtreeTriangles"]) == int(st["instancedTriangles"])
    _same(two, sub)
    orc = orc_mod.Oracle()
    sc.upload(orc)
    ref = orc.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, 0, None, 3)
    parity.assert_hits_bit_exact(sub, ref)
    ctx.close()


def test_subtrees_follow_refit_and_probe_picks_one(rt):
    """Mode 2 rebuilds the subtrees after kfrtRefitTlas; mode 1 (the default) probes a freshly built scene
    once, keeps one structure, and leaves a scene that is being refitted on the two-level walk."""
    sc = pyscene.small_scene(seed=7, w=128, h=96, spp=2, depth=5, lights="dir", textures=True)
    ctx = rt.Context(0)
    sc.upload(ctx)
    insts = np.array(sc.insts, wire.INSTANCE)
    moved = np.array(insts["transform"], "<f4").reshape(-1, 16).copy()
    moved[1:, 12:15] += np.float32(0.125)
    ctx.set_instance_subtrees(2)
    ctx.refit_tlas(moved)
    sub = _buffers(ctx, sc)
    builds = int(ctx.bvh_stats()["subtreeBuilds"])
    assert builds >= 1 and int(ctx.bvh_stats()["subtreeNodeCount"]) > 0
    ctx.set_instance_subtrees(0)
    two = _buffers(ctx, sc)
    _same(two, sub)
    # default mode: a moving scene is not probed
    ctx.set_instance_subtrees(1)
    ctx.refit_tlas(moved)
    auto = _buffers(ctx, sc)
    assert int(ctx.bvh_stats()["subtreeNodeCount"]) == 0
    _same(two, auto)
    # after a full build the probe runs once and the frame it precedes is unchanged
    ctx.set_instances(insts)
    ctx.build_tlas()
    first = _buffers(ctx, sc)
    again = _buffers(ctx, sc)
    _same(first, again)
    ctx.close()


def test_budget_keeps_large_scenes_two_level(rt):
    sc = pyscene.small_scene(seed=3, w=64, h=48, spp=1, depth=3, lights="dir", textures=False)
    ctx = rt.Context(0)
    sc.upload(ctx)
    ctx.set_instance_subtrees(2, max_triangles=8)  # far below the scene's triangle count
    _buffers(ctx, sc)
    assert int(ctx.bvh_stats()["subtreeNodeCount"]) == 0
    ctx.close()


@pytest.mark.parametrize("lights", ["dir", "dir point active"])
def test_culled_light_samples_are_the_ones_the_oracle_would_cull(rt, orc_mod, lights):
    """The kernel's light-sample culling against the oracle's mirror of the rule (kfo_set_cull_light_samples, which
    tests/test_cpu_oracle.py shows to change no bit of the frame): the same light samples are traced and the
    same are answered without a ray (measured: 46 777 / 1 156 and 192 454 / 5 496 on both sides; the bound
    leaves room for a path whose length flips on an ulp)."""
    sc = pyscene.small_scene(seed=11, w=160, h=120, spp=4, depth=8, lights=lights, textures=True, glass=True)
    ctx = rt.Context(0)
    sc.upload(ctx)
    ctx.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, 0, None, 3)
    c = ctx.counters()
    orc = orc_mod.Oracle()
    sc.upload(orc)
    orc.set_cull_light_samples(True)
    ref = orc.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, 0, None, 3)
    traced, skipped = ref["counters"]["shadowRays"], orc.last_shadow_skipped()
    assert skipped > 0
    assert abs(int(c["shadowRays"]) - traced) <= 0.005 * traced, (int(c["shadowRays"]), traced)
    assert abs(int(c["shadowRaysSkipped"]) - skipped) <= max(8, 0.02 * skipped), (int(c["shadowRaysSkipped"]), skipped)
    ctx.close()


@pytest.mark.parametrize("lights", ["dir", "dir point active"])
def test_culled_light_samples_change_nothing(rt, lights):
    """kfrtSetLightSampleCulling / kfrtSetOwnInstanceSkip: with both off every light sample the reference traces
    is traced and every ray walks every instance on its way; with them on (the default) the first-hit buffers
    come out bit for bit and the radiance within the rounding of the shading arithmetic (the switches select
    other instantiations of the shade kernel, whose float contraction differs: same tolerance as against the
    oracle), and the traced + skipped light samples add up to the number traced without the culling.  The scene
    has a glass cube and glass spheres, whose inside surfaces are what the culling rule is about, and convex
    meshes throughout, which is what the own-instance skip is about."""
    sc = pyscene.small_scene(seed=11, w=160, h=120, spp=4, depth=8, lights=lights, textures=True, glass=True)
    ctx = rt.Context(0)
    sc.upload(ctx)
    ctx.set_light_sample_culling(0)
    ctx.set_own_instance_skip(0)
    full = _buffers(ctx, sc)
    c_full = ctx.counters()
    assert int(c_full["shadowRaysSkipped"]) == 0
    for cull, own in ((1, 0), (0, 1), (1, 1)):
        ctx.set_light_sample_culling(cull)
        ctx.set_own_instance_skip(own)
        got = _buffers(ctx, sc)
        c = ctx.counters()
        for k in ("hit_ids", "hit_t", "depth", "albedo", "normal"):
            a, b = full[k], got[k]
            assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                                  b.view(np.uint32) if b.dtype == np.float32 else b), (k, cull, own)
        st = parity.radiance_stats(got["sum"], full["sum"], 4)
        assert st["frac_gt_1e-3"] < 0.02 and st["mean_rel_diff"] < 1e-3, (st, cull, own)
        assert (int(c["shadowRaysSkipped"]) > 0) == bool(cull)
        total = int(c["shadowRays"]) + int(c["shadowRaysSkipped"])
        assert abs(total - int(c_full["shadowRays"])) <= 0.002 * int(c_full["shadowRays"]), (total, int(c_full["shadowRays"]))
        assert abs(int(c["extensionRays"]) - int(c_full["extensionRays"])) <= 0.002 * int(c_full["extensionRays"])
    ctx.close()
