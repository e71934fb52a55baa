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
    ws.upload(orc)
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


def main():
    for case in CASES:
        np.savez_compressed(os.path.join(HERE, f"{case}.npz"), **render_case(case))
        print("wrote", case)
    np.savez_compressed(os.path.join(HERE, "pyscene.npz"), **pyscene_case())
    print("wrote pyscene")


if __name__ == "__main__":
    main()
