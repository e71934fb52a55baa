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
    return got, ref


def assert_hits_bit_exact(got, ref):
    """Hit indices, primitive ids, t and depth of the primary hit: bit-exact (integer compare of the bits)."""
    assert np.array_equal(got["hit_ids"], ref["hit_ids"]), (
        f"{int((got['hit_ids'] != ref['hit_ids']).any(-1).sum())} pixels differ in (instance, primitive)")
    assert np.array_equal(got["hit_t"].view(np.uint32), ref["hit_t"].view(np.uint32)), "hit t bits differ"
    assert np.array_equal(got["depth"].view(np.uint32), ref["depth"].view(np.uint32)), "depth bits differ"


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
