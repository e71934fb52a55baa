"""Edge cases of the hot path, CUDA (through the C ABI) vs the CPU oracle: empty and degenerate inputs,
hidden and alpha-tested geometry (reference PathTrace.rahit:30-48, rt.cpp:101-102,153-157), depth of
field (PathTrace.rgen:40-56), deep instance stacks, and size-independent properties at the BASELINE
frame size (determinism, sample-shard linearity, hit-buffer sanity)."""
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


def _check_radiance(got, ref, spp, frac=0.02):
    st = parity.radiance_stats(got["sum"], ref["sum"], spp)
    assert st["frac_gt_1e-3"] < frac, st
    assert st["mean_rel_diff"] < 2e-3, st


def test_empty_scene_is_all_miss(rt, orc_mod):
    """No instances at all: every path ends in the miss shader, hit buffers are all -1 / 0."""
    sc = pyscene.small_scene(seed=1, w=64, h=48, spp=2, depth=4, lights="dir")
    sc.insts = []
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    assert (got["hit_ids"] == -1).all() and (got["hit_t"] == 0).all()
    assert np.array_equal(got["sum"].view(np.uint32), ref["sum"].view(np.uint32))  # clear colour only: exact
    assert got["counters"]["extensionRays"] == 64 * 48 * 2 and got["counters"]["shadowRays"] == 0
    ctx.close()


def test_hidden_geometry_is_untouchable(rt, orc_mod):
    """hideRender geometries get no bottom level (reference rt.cpp:153-157): rays pass through them."""
    sc = pyscene.small_scene(seed=3, w=96, h=64, spp=2, depth=5, lights="dir point")
    v, i, mi, op, _ = sc.geoms[1]
    sc.geoms[1] = (v, i, mi, op, True)
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    hidden = [k for k, inst in enumerate(sc.insts) if int(inst["geometryIndex"]) == 1]
    assert hidden and not np.isin(got["hit_ids"][..., 0], hidden).any()
    _check_radiance(got, ref, 2)
    ctx.close()


@pytest.mark.parametrize("alpha", [0.0, 0.5])
def test_alpha_tested_geometry(rt, orc_mod, alpha):
    """Non-opaque geometry runs the stochastic any-hit test on extension rays and fully blocks
    occlusion rays (reference PathTrace.rahit:30-48, rchit:186-197)."""
    sc = pyscene.small_scene(seed=4, w=96, h=64, spp=3, depth=4, lights="dir", glass=False)
    for g in (1, 2):
        v, i, mi, _, hide = sc.geoms[g]
        sc.geoms[g] = (v, i, mi, False, hide)
        sc.mats[int(mi[0])]["alpha"] = alpha
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    if alpha == 0.0:
        gone = [k for k, inst in enumerate(sc.insts) if int(inst["geometryIndex"]) in (1, 2)]
        assert not np.isin(got["hit_ids"][..., 0], gone).any()
    _check_radiance(got, ref, 3)
    ctx.close()


def test_depth_of_field_camera(rt, orc_mod):
    """aperture > 0: origin and direction come from the disk sample (reference PathTrace.rgen:40-56)."""
    sc = pyscene.small_scene(seed=6, w=96, h=64, spp=2, depth=3, lights="dir")
    sc.cams = [pyscene.camera([-12.6, 0.0, 8.4], [0.67, 0.0, -0.5], [0, 0, 1], sc.w, sc.h, aperture=0.3, focus=14.0)]
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    _check_radiance(got, ref, 2)
    ctx.close()


def test_degenerate_triangles(rt, orc_mod):
    """Zero-area and repeated-vertex triangles never hit and never poison the boxes with NaNs."""
    sc = pyscene.small_scene(seed=7, w=80, h=60, spp=1, depth=3, lights="dir", n_spheres=3)
    v, i, mi, op, hide = sc.geoms[1]
    extra = np.array([0, 0, 0, 5, 5, 9, 3, 7, 3], np.uint32)  # point, segment, folded
    i2 = np.concatenate([extra, i, extra])
    sc.geoms[1] = (v, i2, np.full(i2.size, mi[0], np.uint32), op, hide)
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    assert np.isfinite(got["sum"]).all()
    ctx.close()


def test_many_overlapping_instances(rt, orc_mod):
    """A pile of interpenetrating instances of one mesh: deep top-level groups, many box re-tests,
    equal-t ties between coincident instances resolve to the lowest instance index."""
    sc = pyscene.small_scene(seed=9, w=96, h=64, spp=1, depth=2, lights="dir", n_spheres=2, glass=False)
    rng = np.random.default_rng(5)
    base = sc.insts[1]
    for k in range(60):
        m = pyscene.translate(rng.uniform(-1.5, 1.5, 3) + [0, 0, 1.0]) @ pyscene.scale(rng.uniform(0.6, 1.4))
        sc.insts.append(pyscene.instance(m, int(base["geometryIndex"])))
    sc.insts.append(sc.insts[-1].copy())  # an exact duplicate: every hit on it is a tie
    ctx, orc = _pair(sc, rt, orc_mod)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    assert not (got["hit_ids"][..., 0] == len(sc.insts) - 1).any()  # the duplicate never wins a tie
    ctx.close()


def test_batched_blas_build_edge_cases(rt, orc_mod):
    """kfrtBuildBlas builds all uploaded geometries in one pass (kf_blas_batch.cuh): geometries of one, two
    and three triangles next to tessellated spheres, a hidden one in the middle of the batch, more than
    1024 geometries (two batches: the geometry index takes 10 key bits), and a rebuild of a single
    re-uploaded geometry beside structures that stay -- all against the oracle, hit for hit."""
    sc = pyscene.small_scene(seed=11, w=96, h=72, spp=1, depth=3, lights="dir", n_spheres=4, glass=False)
    qv, qi = pyscene.quad()
    mat = sc.mats[1]
    one = sc.add_geometry(qv, qi[:3].copy(), mat)
    two = sc.add_geometry(qv, qi.copy(), mat)
    three = sc.add_geometry(qv, np.concatenate([qi, qi[:3][::-1]]).astype(np.uint32), mat)
    hidden = sc.add_geometry(qv, qi.copy(), mat, hide=True)
    for k, g in enumerate((one, two, three, hidden)):
        sc.insts.append(pyscene.instance(pyscene.translate([-2.0 + 1.5 * k, -3.0, 0.5]) @ pyscene.rotate(0.6, [0, 1, 0]), g))
    rng = np.random.default_rng(3)
    first_tile = len(sc.geoms)
    for k in range(1100):  # tiles of two triangles, each its own geometry
        g = sc.add_geometry(qv, qi.copy(), mat)
        pos = [rng.uniform(-6, 6), rng.uniform(-5, 5), rng.uniform(-0.5, 3.0)]
        sc.insts.append(pyscene.instance(pyscene.translate(pos) @ pyscene.rotate(rng.uniform(0, 3), rng.normal(size=3)) @ pyscene.scale(0.25), g))
    ctx, orc = rt.Context(0), orc_mod.Oracle()
    ctx.set_limits(2048, 8192, 16, 4096)
    sc.upload(ctx)
    sc.upload(orc)
    st = ctx.bvh_stats()
    assert int(st["blasCount"]) == len(sc.geoms) and int(st["instanceCount"]) == len(sc.insts)
    got, ref = parity.render_both(sc, ctx, orc)
    parity.assert_hits_bit_exact(got, ref)
    seen = set(np.unique(got["hit_ids"][..., 0]).tolist())
    hidden_inst = [k for k, inst in enumerate(sc.insts) if int(inst["geometryIndex"]) == hidden]
    assert not seen.intersection(hidden_inst)
    assert len([k for k in seen if k >= 0 and int(sc.insts[k]["geometryIndex"]) >= first_tile]) > 50
    # one geometry uploaded again with other triangles: only it is rebuilt, the rest keep their storage
    sv, si = pyscene.uv_sphere(7, 9)
    v, i, mi, op, hide = sc.geoms[two]
    sc.geoms[two] = (sv, si, np.full(si.size, mi[0], np.uint32), op, hide)
    ctx.upload_geometry(two, sv, si, sc.geoms[two][2], op, hide)
    ctx.build_blas()
    ctx.set_instances(np.array(sc.insts, wire.INSTANCE))
    ctx.build_tlas()
    orc2 = orc_mod.Oracle()
    sc.upload(orc2)
    got, ref = parity.render_both(sc, ctx, orc2)
    parity.assert_hits_bit_exact(got, ref)
    ctx.close()


# ---- properties at the BASELINE frame size (config 3 through the facade; no oracle pass needed) ----
@pytest.fixture(scope="module")
def million(built):
    from kuafu_b200 import host, rt as _rt
    r = host.Renderer(device=0, accumulate=False)
    r.load_scene("million", 1920, 1080, 4, 8)
    ws = r.wire_scene()
    r.run()
    yield r, _rt.Context(handle=r.device_context()), ws
    r.close()


def test_full_size_determinism_and_hit_sanity(million):
    r, ctx, ws = million
    cams = np.array(ws.cams[:1], wire.CAMERA)
    ctx.render(cams, 1920, 1080, ws.pc, 0, 4, clock_base=11)
    a = ctx.download_aux(wire.AUX_SUM32F).copy()
    ids = ctx.download_aux(wire.AUX_HIT_IDS).copy()
    t = ctx.download_aux(wire.AUX_HIT_T).copy()
    depth = ctx.download_aux(wire.AUX_DEPTH).copy()
    ctx.render(cams, 1920, 1080, ws.pc, 0, 4, clock_base=11)
    assert np.array_equal(ctx.download_aux(wire.AUX_SUM32F).view(np.uint32), a.view(np.uint32))  # idempotent
    hit = ids[..., 0] >= 0
    stats = ctx.bvh_stats()
    assert hit.mean() > 0.5
    assert ids[..., 0].max() < int(stats["instanceCount"]) and (ids[..., 1][hit] >= 0).all()
    assert (ids[~hit] == -1).all() and (t[~hit] == 0).all() and (depth[~hit] == 0).all()
    assert (t[hit] > 0.001).all() and (t[hit] < 10000).all() and (depth[hit] > 0).all()
    assert (depth[hit] <= t[hit] * (1 + 1e-5)).all()  # view-space depth never exceeds the ray length
    assert np.isfinite(a).all() and (a[..., :3] >= 0).all() and (a[..., 3] == 4).all()
    c = ctx.counters()
    assert int(c["paths"]) == 1920 * 1080 * 4 and int(c["extensionHits"]) <= int(c["extensionRays"])


def test_full_size_sample_shards_add_up(million):
    """Linearity over the sample range: SUM[0,4) == SUM[0,2) + SUM[2,4) up to float association, and
    the hit buffers of a shard that starts at sample 0 equal those of the full frame."""
    r, ctx, ws = million
    cams = np.array(ws.cams[:1], wire.CAMERA)
    ctx.render(cams, 1920, 1080, ws.pc, 0, 4, clock_base=21)
    full = ctx.download_aux(wire.AUX_SUM32F).astype(np.float64)
    ids = ctx.download_aux(wire.AUX_HIT_IDS).copy()
    ctx.render(cams, 1920, 1080, ws.pc, 0, 2, clock_base=21)
    lo = ctx.download_aux(wire.AUX_SUM32F).astype(np.float64)
    assert np.array_equal(ctx.download_aux(wire.AUX_HIT_IDS), ids)
    ctx.render(cams, 1920, 1080, ws.pc, 2, 4, clock_base=21)
    hi = ctx.download_aux(wire.AUX_SUM32F).astype(np.float64)
    err = np.abs(lo + hi - full)[..., :3]
    assert (err <= 1e-5 * (np.abs(full[..., :3]) + 1.0)).all(), err.max()
    assert (lo[..., 3] == 2).all() and (hi[..., 3] == 2).all()


def test_full_size_refit_with_same_transforms_is_identity(million):
    r, ctx, ws = million
    cams = np.array(ws.cams[:1], wire.CAMERA)
    ctx.render(cams, 1920, 1080, ws.pc, 0, 1, clock_base=31)
    ids = ctx.download_aux(wire.AUX_HIT_IDS).copy()
    t = ctx.download_aux(wire.AUX_HIT_T).copy()
    tr = np.ascontiguousarray(np.array(ws.insts)["transform"], np.float32)
    ctx.refit_tlas(tr)
    ctx.render(cams, 1920, 1080, ws.pc, 0, 1, clock_base=31)
    assert np.array_equal(ctx.download_aux(wire.AUX_HIT_IDS), ids)
    assert np.array_equal(ctx.download_aux(wire.AUX_HIT_T).view(np.uint32), t.view(np.uint32))


def test_imported_asset_parity(rt, orc_mod, tmp_path):
    """A mesh file through kuafu::loadScene (binary STL of a displaced icosphere-like blob written here)
    rendered by the CUDA path and by the oracle from the same packed scene: the import path feeds the
    hot path like any procedural geometry."""
    import struct
    from kuafu_b200 import host
    rng = np.random.default_rng(5)
    n_lat, n_lon = 24, 48
    th = np.linspace(0.0, np.pi, n_lat + 1)[:, None]
    ph = np.linspace(0.0, 2.0 * np.pi, n_lon + 1)[None, :]
    rad = 1.0 + 0.15 * np.sin(3 * th) * np.cos(4 * ph)
    pts = np.stack([rad * np.sin(th) * np.cos(ph), rad * np.sin(th) * np.sin(ph), rad * np.cos(th) * np.ones_like(ph)], -1).astype(np.float32)
    tris = []
    for i in range(n_lat):
        for j in range(n_lon):
            a, b, c, d = pts[i, j], pts[i + 1, j], pts[i + 1, j + 1], pts[i, j + 1]
            tris += [(a, b, c), (a, c, d)]
    path = str(tmp_path / "blob.stl")
    with open(path, "wb") as f:
        f.write(b"\0" * 80 + struct.pack("<I", len(tris)))
        for a, b, c in tris:
            nrm = np.cross(b - a, c - a)
            ln = np.linalg.norm(nrm)
            nrm = nrm / ln if ln > 0 else np.array([0, 0, 1], np.float32)
            f.write(struct.pack("<12fH", *nrm, *a, *b, *c, 0))
    r = host.Renderer(device=None)
    r.load_scene("file:" + path, 96, 72, 2)
    ws = r.wire_scene()
    assert ws.n_tris() == len(tris)
    ctx, orc = rt.Context(0), orc_mod.Oracle()
    ws.upload(ctx)
    orc.load(ws)
    got, ref = parity.render_both(ws, ctx, orc, clock_base=2)
    parity.assert_hits_bit_exact(got, ref)
    assert (got["hit_ids"][..., 0] >= 0).mean() > 0.08     # the framing camera sees the asset
    st = parity.radiance_stats(got["sum"], ref["sum"], 2)
    assert st["frac_gt_1e-3"] < 0.02 and st["mean_rel_diff"] < 2e-3, st
    ctx.close()
    r.close()
