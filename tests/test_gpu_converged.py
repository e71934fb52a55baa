"""RMSE of the 4096-spp converged image (north_star's second radiance bar) against the frozen oracle
renders in tests/golden/converged_*.npz (made by tests/golden/make_golden.py --converged).

Each fixture holds the oracle's 4096-spp image at clock surrogate `clock` and `noise_rmse`, the RMSE
between that image and a second oracle render with disjoint RNG streams (`other_clock`) -- the
Monte-Carlo noise floor at 4096 spp.  Two bars, both stated relative to that floor:
  * same streams (same clock): the kernels trace the same paths as the oracle, so only float
    association / transcendental differences remain: RMSE <= 0.25 x noise floor;
  * independent streams (a third clock): the kernels must converge to the same image:
    RMSE <= 1.5 x noise floor and image means within 1 %.
"""
import os

import numpy as np
import pytest

from kuafu_b200 import wire

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("converged_cornell", "cornell", 32, 32), ("converged_spheres", "spheres", 40, 30)]


def _render(rt, ws, clock):
    ctx = rt.Context(0)
    ws.upload(ctx)
    spp = int(ws.pc["sampleRatePerPixel"])
    ctx.render(np.array(ws.cams[:1], wire.CAMERA), ws.w, ws.h, ws.pc, 0, spp, clock)
    img = ctx.download_aux(wire.AUX_SUM32F, 0)[..., :3].astype(np.float64) / spp
    ctx.close()
    return img


@pytest.mark.parametrize("fixture,recipe,w,h", CASES)
def test_converged_image_rmse(built, fixture, recipe, w, h):
    from kuafu_b200 import host, rt
    g = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    spp, noise = int(g["spp"]), float(g["noise_rmse"])
    ref = g["sum"][..., :3].astype(np.float64) / spp
    r = host.Renderer(device=None)
    r.load_scene(recipe, w, h, spp, 0, 0)
    ws = r.wire_scene()

    same = _render(rt, ws, int(g["clock"]))
    rmse_same = float(np.sqrt(((same - ref) ** 2).mean()))
    assert rmse_same <= 0.25 * noise, (rmse_same, noise)

    other = _render(rt, ws, 1000 + int(g["other_clock"]))
    rmse_other = float(np.sqrt(((other - ref) ** 2).mean()))
    assert rmse_other <= 1.5 * noise, (rmse_other, noise)
    assert abs(other.mean() - ref.mean()) <= 0.01 * abs(ref.mean()), (other.mean(), ref.mean())
    print(f"{fixture}: noise floor {noise:.5f}, same streams {rmse_same:.6f}, independent {rmse_other:.5f}")
    r.close()
