"""Shared comparison helpers: CUDA path (through the C ABI) vs the CPU oracle."""
import numpy as np

from kuafu_b200 import wire


def render_both(sc, ctx, orc, clock_base=3, sample_begin=0, sample_end=None, brute=False):
    ctx.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, sample_begin, sample_end, clock_base)
    ref = orc.render(np.array(sc.cams, wire.CAMERA), sc.w, sc.h, sc.pc, sample_begin, sample_end, clock_base,
                     brute=brute)
    n = len(sc.cams)
    got = {
        "sum": np.stack([ctx.download_aux(wire.AUX_SUM32F, c) for c in range(n)]),
        "albedo": np.stack([ctx.download_aux(wire.AUX_ALBEDO32F, c) for c in range(n)]),
        "normal": np.stack([ctx.download_aux(wire.AUX_NORMAL32F, c) for c in range(n)]),
        "hit_ids": np.stack([ctx.download_aux(wire.AUX_HIT_IDS, c) for c in range(n)]),
        "hit_t": np.stack([ctx.download_aux(wire.AUX_HIT_T, c) for c in range(n)]),
        "depth": np.stack([ctx.download_aux(wire.AUX_DEPTH, c) for c in range(n)]),
    }
    cnt = ctx.counters()
    got["counters"] = {k: int(cnt[k]) for k in ("paths", "extensionRays", "shadowRays", "extensionHits")}
    # light samples whose occlusion ray cannot change the image are not traced by the CUDA path (kf_wavefront.cuh,
    # nextRelevantLight); the oracle, like the reference, traces them: they count as shadow rays in comparisons
    got["counters"]["shadowRaysTraced"] = got["counters"]["shadowRays"]
    got["counters"]["shadowRays"] += int(cnt["shadowRaysSkipped"])
    return got, ref


def assert_hits_bit_exact(got, ref):
    """Hit indices, primitive ids, t and depth of the primary hit: bit-exact (integer compare of the bits)."""
    assert np.array_equal(got["hit_ids"], ref["hit_ids"]), (
        f"{int((got['hit_ids'] != ref['hit_ids']).any(-1).sum())} pixels differ in (instance, primitive)")
    assert np.array_equal(got["hit_t"].view(np.uint32), ref["hit_t"].view(np.uint32)), "hit t bits differ"
    assert np.array_equal(got["depth"].view(np.uint32), ref["depth"].view(np.uint32)), "depth bits differ"
    # (the oracle writes the first-hit buffers only for a sample range that starts at sample 0)
    if "albedo" in got and "albedo" in ref and (ref["albedo"][..., 3] == 1.0).all():
        assert_aux_close(got, ref)


def assert_aux_close(got, ref, atol=2e-5):
    """First-hit albedo and shading normal, the denoiser hand-off buffers (reference PathTrace.rgen:114-117,
    165-166): same hit and same barycentrics on both sides (asserted bit for bit above), so what is left is
    the rounding of the texture interpolation, the division by pi and the normalisation -- a few ulps of
    values in [0, 1].  Alpha is exactly 1; pixels whose first hit is a miss or an emitter are exactly 0."""
    for k in ("albedo", "normal"):
        g, r = got[k], ref[k]
        assert g.shape == r.shape, k
        assert np.array_equal(g[..., 3], r[..., 3]) and (g[..., 3] == 1.0).all(), f"{k} alpha"
        d = np.abs(g[..., :3].astype(np.float64) - r[..., :3].astype(np.float64))
        assert d.max() <= atol, f"{k}: max abs difference {d.max():.3g} at {np.unravel_index(d.argmax(), d.shape)}"
    miss = ref["hit_ids"][..., 0] < 0
    assert not got["albedo"][..., :3][miss].any() and not got["normal"][..., :3][miss].any()
    hit = ~miss
    n = got["normal"][..., :3][hit].astype(np.float64)
    lit = np.abs(n).sum(-1) > 0  # an emissive first hit leaves the normal at zero
    if lit.any():
        assert np.abs(np.linalg.norm(n[lit], axis=-1) - 1.0).max() < 1e-4, "shading normals are not unit length"


def radiance_stats(got_sum, ref_sum, spp):
    g = got_sum[..., :3].astype(np.float64) / spp
    r = ref_sum[..., :3].astype(np.float64) / spp
    assert np.isfinite(g).all() == np.isfinite(r).all()
    ok = np.isfinite(g).all(-1) & np.isfinite(r).all(-1)
    g, r = g[ok], r[ok]
    rel = np.abs(g - r).max(-1) / (np.abs(r).max(-1) + 1e-3)
    return {
        "frac_gt_1e-3": float((rel > 1e-3).mean()),
        "frac_gt_1e-1": float((rel > 1e-1).mean()),
        "median_rel": float(np.median(rel)),
        "mean_rel_diff": float(abs(g.mean() - r.mean()) / (abs(r.mean()) + 1e-12)),
        "rmse": float(np.sqrt(((g - r) ** 2).mean())),
    }
