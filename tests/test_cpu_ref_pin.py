"""Pins the hand-written oracle (oracle/kf_oracle.cpp) to the reference's own shader code.

oracle/_ref/libkf_ref.so is /root/reference/resources/shaders/PathTrace.{rgen,rchit,rahit,rmiss} +
PathTraceShadow.rmiss + base/*.glsl compiled for the CPU (oracle/Makefile `ref`: glsl2cpp.py +
glsl_shim.hpp + kf_ref_host.cpp).  Both renderers share the black boxes the GLSL delegates to driver
and hardware (traversal, triangle test, texture samplers, clock surrogate), so everything the
reference *wrote* -- camera rays, RNG streams, BSDF sampling, next-event estimation for all light
types, miss / environment lookup, Russian roulette, accumulation -- must agree BIT FOR BIT.

Two layers, so the pin survives on machines without the reference tree:
  * live: shim vs oracle on seeded scenes (needs oracle/_ref, built here from /root/reference);
  * frozen: tests/golden/ref_*.npz, outputs of the shim written by make_golden.py --ref, compared with
    the oracle alone.
"""
import os
import sys

import numpy as np
import pytest

import pyscene
from kuafu_b200 import wire
from oracle import oracle, ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and no reference tree here")


def _feed(orc, sc):
    """A facade scene view goes through Oracle.load, a tests/pyscene.py scene through its own upload."""
    if type(sc).__name__ == "WireSceneView":
        orc.load(sc)
    else:
        sc.upload(orc)


def both(sc, clock, brute=False, frame_count=0, prev=None):
    orc = oracle.Oracle()
    _feed(orc, sc)
    cams = np.array(sc.cams)
    pc = np.array(sc.pc).copy()
    pc["frameCount"] = frame_count
    spp = int(pc.reshape(-1)[0]["sampleRatePerPixel"])
    a = orc.render(cams, sc.w, sc.h, pc, clock_base=clock, brute=brute)
    rgba = np.zeros_like(a["sum"]) if prev is None else prev.copy()
    orc.resolve(a["sum"], rgba, spp, frame_count)
    a["image"] = rgba
    b = ref.render(orc, cams, sc.w, sc.h, pc, clock_base=clock, brute=brute, image=prev)
    return a, b


def assert_identical(a, b):
    for k in ("image", "albedo", "normal", "hit_t"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), f"{k}: bits differ"
    assert np.array_equal(a["hit_ids"], b["hit_ids"])
    assert a["counters"] == b["counters"]


@needs_ref
@pytest.mark.parametrize("lights", ["dir", "point", "active", "dir point active"])
def test_shaders_vs_oracle_small_scene(lights):
    sc = pyscene.small_scene(seed=11, w=48, h=32, spp=2, depth=5, lights=lights, textures=True, env=True,
                             emissive=True, rr=True)
    a, b = both(sc, clock=3)
    assert_identical(a, b)
    assert a["counters"]["extensionHits"] > 0 and (lights == "" or a["counters"]["shadowRays"] > 0)


@needs_ref
def test_shaders_vs_oracle_variants():
    # depth of field (aperture > 0 branch of PathTrace.rgen:46-50), clear colour instead of an
    # environment map, no Russian roulette, brute-force traversal
    sc = pyscene.small_scene(seed=5, w=40, h=30, spp=3, depth=6, lights="dir point", textures=False, env=False, rr=False)
    sc.cams = [pyscene.camera([-12.6, 0.0, 8.4], [0.67, 0.0, -0.5], [0, 0, 1], sc.w, sc.h, aperture=0.4, focus=14.0)]
    a, b = both(sc, clock=9)
    assert_identical(a, b)
    a, b = both(sc, clock=9, brute=True)
    assert_identical(a, b)


@needs_ref
def test_shaders_vs_oracle_accumulation():
    # frameCount > 0: mix(old, new, 1 / (frameCount + 1)) on the storage image (PathTrace.rgen:153-163)
    sc = pyscene.small_scene(seed=7, w=32, h=24, spp=2, depth=4, lights="dir", env=True)
    a0, b0 = both(sc, clock=0)
    assert_identical(a0, b0)
    a1, b1 = both(sc, clock=2, frame_count=1, prev=b0["image"])
    assert_identical(a1, b1)
    assert not np.array_equal(a1["image"], a0["image"])


@needs_ref
def test_alpha_zero_geometry_is_ignored_by_rahit():
    # alpha == 0 -> ignoreIntersectionEXT without a random draw (PathTrace.rahit:38-41): bit-exact.
    sc = pyscene.small_scene(seed=3, w=40, h=30, spp=2, depth=4, lights="dir")
    qv, qi = pyscene.quad()
    g = sc.add_geometry(qv, qi, pyscene.material(diffuse=(1, 0, 0), alpha=0.0), opaque=False)
    sc.insts.append(pyscene.instance(pyscene.translate([0, 0, 3]) @ pyscene.scale(6.0), g))
    a, b = both(sc, clock=4)
    assert_identical(a, b)
    assert not (a["hit_ids"][..., 0] == len(sc.insts) - 1).any()


@needs_ref
def test_stochastic_alpha_is_the_declared_deviation():
    # 0 < alpha < 1: the reference advances ray.seed once per any-hit invocation in hardware
    # traversal order (implementation-defined); the oracle draws from a hash instead (deviation D5).
    # The two must agree statistically, not bit for bit.
    sc = pyscene.small_scene(seed=3, w=48, h=36, spp=16, depth=3, lights="dir")
    qv, qi = pyscene.quad()
    g = sc.add_geometry(qv, qi, pyscene.material(diffuse=(1, 0, 0), alpha=0.5), opaque=False)
    sc.insts.append(pyscene.instance(pyscene.translate([0, 0, 3]) @ pyscene.scale(6.0), g))
    a, b = both(sc, clock=4)
    ma, mb = a["image"][..., :3].mean(), b["image"][..., :3].mean()
    assert abs(ma - mb) < 0.05 * mb, (ma, mb)


CONFIGS = [("spheres", 80, 60, 2, 0, 5), ("cornell", 48, 48, 4, 0, 9), ("million", 64, 36, 2, 40, 2),
           ("active", 64, 36, 2, 0, 4), ("articulated", 48, 48, 2, 4, 1)]


def _config_scene(name, w, h, spp, scale):
    from kuafu_b200 import host
    r = host.Renderer(device=None)
    r.load_scene(name, w, h, spp, 0, scale)
    ws = r.wire_scene()
    ws.cams = ws.cams[:1]
    return ws


@needs_ref
@pytest.mark.parametrize("name,w,h,spp,scale,clock", CONFIGS)
def test_shaders_vs_oracle_baseline_configs(built, name, w, h, spp, scale, clock):
    a, b = both(_config_scene(name, w, h, spp, scale), clock=clock)
    assert_identical(a, b)


REF_ASSET_CONFIGS = [("spheres_ref", 80, 60, 2, 0, 5), ("active_ref", 96, 54, 2, 0, 4)]


@needs_ref
@pytest.mark.parametrize("name,w,h,spp,scale,clock", REF_ASSET_CONFIGS)
def test_shaders_vs_oracle_with_the_reference_assets(built, name, w, h, spp, scale, clock):
    """Configs 1 and 4 with the reference's own assets instead of the procedural stand-ins:
    resources/models/suzanne.dae through the COLLADA importer and, for config 4, the projector pattern
    resources/patterns/fakesense_j415.png (3000 x 3000) through the PNG reader (Example.hpp:292-306, 496).
    The assets cannot ship, so these recipes exist only where /root/reference is mounted."""
    if not os.path.exists("/root/reference/resources/patterns/fakesense_j415.png"):
        pytest.skip("reference assets are not mounted")
    sc = _config_scene(name, w, h, spp, scale)
    assert sc.n_tris() > 270000  # the scanned head is in the scene
    if name == "active_ref":
        assert any(t.shape[:2] == (3000, 3000) for t in sc.textures)
    a, b = both(sc, clock=clock)
    assert_identical(a, b)
    assert a["counters"]["shadowRays"] > 0


@pytest.mark.parametrize("name,w,h,spp,scale,clock", CONFIGS)
def test_oracle_vs_frozen_shader_outputs(built, name, w, h, spp, scale, clock):
    """tests/golden/ref_<name>.npz was written by the shim (make_golden.py --ref); no reference needed."""
    g = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    sc = _config_scene(name, w, h, spp, scale)
    orc = oracle.Oracle()
    _feed(orc, sc)
    a = orc.render(np.array(sc.cams), sc.w, sc.h, sc.pc, clock_base=clock)
    rgba = np.zeros_like(a["sum"])
    orc.resolve(a["sum"], rgba, spp, 0)
    assert np.array_equal(rgba[0].view(np.uint32), g["image"].view(np.uint32))
    assert np.array_equal(a["albedo"][0].view(np.uint32), g["albedo"].view(np.uint32))
    assert np.array_equal(a["normal"][0].view(np.uint32), g["normal"].view(np.uint32))
    assert np.array_equal(a["hit_ids"][0], g["hit_ids"])
    assert [a["counters"][k] for k in ("paths", "extensionRays", "shadowRays", "extensionHits")] == list(g["counters"])
