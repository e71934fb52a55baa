"""BASELINE configs through the C++ facade (kuafu.hpp API via the flat C view), at reduced resolution so
the CPU oracle finishes in seconds.  Two checks per config:
  * the wire buffers the facade packs, fed to the raw C ABI and to the oracle, give bit-exact hit
    buffers and toleranced radiance;
  * Kuafu::run() + downloadLatestFrame() produce exactly what the raw C-ABI calls produce.
The configurations at their own size: test_gpu_fullsize.py (against the oracle) and the size-independent
properties in test_gpu_edge.py."""
import numpy as np
import pytest

import parity
from kuafu_b200 import wire

pytestmark = pytest.mark.gpu

CASES = [
    # name, w, h, spp, scale
    ("spheres", 200, 150, 4, 0),
    ("cornell", 128, 128, 8, 0),
    ("million", 240, 135, 2, 0),
    ("million_obj", 240, 135, 2, 0),   # config 3 "loaded via the Assimp path": meshes through loadScene()
    ("active", 160, 90, 4, 0),
    ("articulated", 96, 96, 2, 8),
]


@pytest.fixture(scope="module")
def mods(built):
    from kuafu_b200 import host, rt
    from oracle import oracle
    return host, rt, oracle


@pytest.mark.parametrize("name,w,h,spp,scale", CASES)
def test_config_parity_and_facade(mods, name, w, h, spp, scale):
    host, rt, oracle = mods
    r = host.Renderer(device=0, accumulate=False)
    ncam = r.load_scene(name, w, h, spp, 0, scale)
    ws = r.wire_scene()
    assert ws.pc["sampleRatePerPixel"] == spp
    ncheck = min(ncam, 2)
    ws.cams = ws.cams[:ncheck]

    ctx = rt.Context(0)
    orc = oracle.Oracle()
    ws.upload(ctx)
    orc.load(ws)
    got, ref = parity.render_both(ws, ctx, orc, clock_base=0)
    parity.assert_hits_bit_exact(got, ref)
    st = parity.radiance_stats(got["sum"], ref["sum"], spp)
    # tolerance: <= 2 % of pixels off by more than 1e-3 relative, image means within 2e-3
    assert st["frac_gt_1e-3"] < 0.02 and st["mean_rel_diff"] < 2e-3, st
    assert (got["hit_ids"][..., 0] >= 0).mean() > 0.1  # the camera actually sees the scene

    # the facade path: same clock base (0 on a fresh renderer), current camera = recipe camera 0
    ctx.resolve()
    r.run()
    assert np.array_equal(r.download_frame(0), ctx.download_bgra8(0))
    assert np.array_equal(r.download_aux(wire.AUX_HIT_IDS, 0), got["hit_ids"][0])
    assert np.array_equal(r.download_aux(wire.AUX_DEPTH, 0).view(np.uint32), got["depth"][0].view(np.uint32))
    seg = r.download_aux(wire.AUX_SEGMENTATION, 0)
    assert np.array_equal(seg, got["hit_ids"][0][..., 0])
    ctx.close()
    r.close()


def test_articulated_refit_through_facade(mods):
    """Config 5 shape: per-frame actor motion -> setTransform -> refit inside Kuafu::run()."""
    host, rt, oracle = mods
    r = host.Renderer(device=0, accumulate=False)
    r.load_scene("articulated", 96, 96, 1, 0, 6)
    orc = oracle.Oracle()
    orc.load(r.wire_scene())
    for frame in range(3):
        r.animate(frame)
        ws = r.wire_scene()
        orc.set_transforms(ws.insts["transform"])
        r.clock_base = 100 + frame
        r.run_all()
        for cam in (0, 3):
            ref = orc.render(np.array(ws.cams[cam:cam + 1]), ws.w, ws.h, ws.pc, clock_base=100 + frame)
            assert np.array_equal(r.download_aux(wire.AUX_HIT_IDS, cam), ref["hit_ids"][0])
            assert np.array_equal(r.download_aux(wire.AUX_HIT_T, cam).view(np.uint32), ref["hit_t"][0].view(np.uint32))
    r.close()


def test_refit_quality_watch_rebuilds_without_changing_hits(mods):
    """Forty frames of actor motion: kfrtRefitTlas rebuilds the top level when the refitted hierarchy has
    stretched (area sum past 1.1x its value at build time); whichever of the two it did, the frame still
    hits what the oracle hits for the same transforms."""
    host, rt, oracle = mods
    r = host.Renderer(device=0, accumulate=False)
    r.load_scene("articulated", 64, 64, 1, 0, 6)
    ctx = rt.Context(handle=r.device_context())
    for frame in range(1, 41):
        r.animate(frame)
        r.clock_base = 7
        r.run()
        r.download_frame(0)  # a renderer loop looks at its frames; the watch reads its area sum back asynchronously
    assert int(ctx.bvh_stats()["tlasRebuilds"]) >= 1
    ws = r.wire_scene()
    orc = oracle.Oracle()
    orc.load(ws)
    ref = orc.render(np.array(ws.cams[:1]), ws.w, ws.h, ws.pc, clock_base=7)
    assert np.array_equal(r.download_aux(wire.AUX_HIT_IDS, 0), ref["hit_ids"][0])
    assert np.array_equal(r.download_aux(wire.AUX_HIT_T, 0).view(np.uint32), ref["hit_t"][0].view(np.uint32))
    r.close()


def test_accumulation_frames(mods):
    """frameCount semantics: accumulate on -> 0,1,2..; running mean equals the oracle's resolve."""
    host, rt, oracle = mods
    r = host.Renderer(device=0, accumulate=True)
    r.load_scene("cornell", 64, 64, 2, 0, 0)
    orc = oracle.Oracle()
    rgba = np.zeros((1, 64, 64, 4), np.float32)
    for f in range(3):
        r.run()
        assert r.frame_count() == f
        s = r.download_aux(wire.AUX_SUM32F, 0)[None]
        bgra = orc.resolve(s, rgba, 2, f)
        assert np.array_equal(r.download_frame(0), bgra[0])
    r.close()


def test_displaced_frame_is_stashed(mods):
    """render(A); render(B); A.downloadLatestFrame() still returns A's frame (per-camera images)."""
    host, rt, oracle = mods
    r = host.Renderer(device=0, accumulate=False)
    r.load_scene("active", 96, 54, 1, 0, 0)
    r.set_camera(0)
    r.run()
    a = r.download_frame(0)
    r.set_camera(1)
    r.run()
    b = r.download_frame(1)
    assert np.array_equal(r.download_frame(0), a)
    assert not np.array_equal(a, b)
    r.close()


def test_config3_import_path_hits_the_same_triangles(mods):
    """million_obj (every mesh written to a Wavefront file and read back through kuafu::loadScene)
    against million (procedural): positions and face order survive the round trip bit for bit, so the
    primary-hit buffers are identical; v moved by at most one ulp (two FlipUVs), so radiance agrees to
    the texture-lookup tolerance."""
    host, rt, oracle = mods
    out = {}
    for name in ("million", "million_obj"):
        r = host.Renderer(device=0, accumulate=False)
        r.load_scene(name, 240, 135, 2)
        r.run()
        out[name] = {k: r.download_aux(kind, 0) for k, kind in
                     (("ids", wire.AUX_HIT_IDS), ("t", wire.AUX_HIT_T), ("depth", wire.AUX_DEPTH), ("sum", wire.AUX_SUM32F))}
        r.close()
    a, b = out["million"], out["million_obj"]
    assert np.array_equal(a["ids"], b["ids"])
    assert np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    st = parity.radiance_stats(a["sum"][None], b["sum"][None], 2)
    assert st["frac_gt_1e-3"] < 0.02 and st["mean_rel_diff"] < 2e-3, st
