#!/usr/bin/env python
"""Generates tests/golden/*.npz: frozen outputs of the CPU oracle on small seeded scenes.

The reference ships no golden vectors for this path, and its Vulkan-RT pipeline cannot
run here, so these are OUR oracle's outputs at fixed seeds, frozen so that later changes to the oracle
(or to the facade's scene packing) cannot silently move the goalposts.  Integer outputs (hit ids, t
bits, depth bits, BGRA8) are stored exactly; radiance as float32.

    python tests/golden/make_golden.py        # rewrites the fixtures (run in this container)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (recipe, w, h, spp, scale, clock_base)
    "spheres": ("spheres", 80, 60, 2, 0, 5),
    "cornell": ("cornell", 48, 48, 4, 0, 9),
    "million": ("million", 64, 36, 2, 40, 2),
    "active": ("active", 64, 36, 2, 0, 4),
    "articulated": ("articulated", 48, 48, 2, 4, 1),
}


def render_case(case):
    from kuafu_b200 import host
    from oracle import oracle
    recipe, w, h, spp, scale, clock = CASES[case]
    r = host.Renderer(device=None)
    r.load_scene(recipe, w, h, spp, 0, scale)
    ws = r.wire_scene()
    orc = oracle.Oracle()
    orc.load(ws)
    out = orc.render(np.array(ws.cams[:1]), ws.w, ws.h, ws.pc, clock_base=clock, threads=1)
    rgba = np.zeros_like(out["sum"])
    bgra = orc.resolve(out["sum"], rgba, spp, 0)
    return {
        "hit_ids": out["hit_ids"][0], "hit_t_bits": out["hit_t"][0].view(np.uint32),
        "depth_bits": out["depth"][0].view(np.uint32), "sum": out["sum"][0], "bgra": bgra[0],
        "albedo": out["albedo"][0], "normal": out["normal"][0],
        "counters": np.array([out["counters"][k] for k in ("paths", "extensionRays", "shadowRays", "extensionHits")], np.uint64),
        "n_tris": np.array(ws.n_tris(), np.uint64),
    }


def pyscene_case():
    import pyscene
    from oracle import oracle
    sc = pyscene.small_scene(seed=11, w=48, h=32, spp=2, depth=5, lights="dir point active", textures=True,
                             env=True, emissive=True, rr=True)
    orc = oracle.Oracle()
    sc.upload(orc)
    out = orc.render(np.array(sc.cams), sc.w, sc.h, sc.pc, clock_base=3, threads=1)
    return {"hit_ids": out["hit_ids"][0], "hit_t_bits": out["hit_t"][0].view(np.uint32), "sum": out["sum"][0],
            "counters": np.array([out["counters"][k] for k in ("paths", "extensionRays", "shadowRays", "extensionHits")], np.uint64)}


# 4096-spp converged images (north_star: "RMSE of the 4096-spp converged image").  Two oracle renders
# with different clock surrogates (= disjoint RNG streams) are stored: `sum` is the reference image,
# `noise_rmse` the RMSE between the two oracle renders, i.e. the Monte-Carlo noise floor that any
# correct renderer with independent seeds must reproduce.
CONVERGED = {
    # name: (recipe, w, h, spp, scale, clock_a, clock_b)
    "converged_cornell": ("cornell", 32, 32, 4096, 0, 0, 7),
    "converged_spheres": ("spheres", 40, 30, 4096, 0, 0, 7),
}


def converged_case(case):
    from kuafu_b200 import host
    from oracle import oracle
    recipe, w, h, spp, scale, ca, cb = CONVERGED[case]
    r = host.Renderer(device=None)
    r.load_scene(recipe, w, h, spp, 0, scale)
    ws = r.wire_scene()
    orc = oracle.Oracle()
    orc.load(ws)
    cams = np.array(ws.cams[:1])
    a = orc.render(cams, ws.w, ws.h, ws.pc, clock_base=ca)["sum"][0]
    b = orc.render(cams, ws.w, ws.h, ws.pc, clock_base=cb)["sum"][0]
    ia, ib = a[..., :3].astype(np.float64) / spp, b[..., :3].astype(np.float64) / spp
    return {"sum": a, "spp": np.array(spp), "clock": np.array(ca), "other_clock": np.array(cb),
            "noise_rmse": np.array(np.sqrt(((ia - ib) ** 2).mean())), "mean": np.array(ia.mean())}


def ref_case(case):
    """Outputs of the REFERENCE'S OWN SHADERS (oracle/_ref, built from /root/reference by `make -C oracle
    ref`) for the same small cases: what pins the oracle on machines without the reference tree."""
    from kuafu_b200 import host
    from oracle import oracle, ref
    recipe, w, h, spp, scale, clock = CASES[case]
    r = host.Renderer(device=None)
    r.load_scene(recipe, w, h, spp, 0, scale)
    ws = r.wire_scene()
    orc = oracle.Oracle()
    orc.load(ws)
    out = ref.render(orc, np.array(ws.cams[:1]), ws.w, ws.h, ws.pc, clock_base=clock)
    return {"image": out["image"][0], "albedo": out["albedo"][0], "normal": out["normal"][0],
            "hit_ids": out["hit_ids"][0], "hit_t_bits": out["hit_t"][0].view(np.uint32),
            "counters": np.array([out["counters"][k] for k in ("paths", "extensionRays", "shadowRays", "extensionHits")], np.uint64)}


def main():
    if "--ref" in sys.argv:
        for case in CASES:
            np.savez_compressed(os.path.join(HERE, f"ref_{case}.npz"), **ref_case(case))
            print("wrote ref_" + case)
        return
    for case in CONVERGED:
        if "--converged" in sys.argv or not os.path.exists(os.path.join(HERE, f"{case}.npz")):
            np.savez_compressed(os.path.join(HERE, f"{case}.npz"), **converged_case(case))
            print("wrote", case)
    if "--converged" in sys.argv:
        return
    for case in CASES:
        np.savez_compressed(os.path.join(HERE, f"{case}.npz"), **render_case(case))
        print("wrote", case)
    np.savez_compressed(os.path.join(HERE, "pyscene.npz"), **pyscene_case())
    print("wrote pyscene")


if __name__ == "__main__":
    main()
