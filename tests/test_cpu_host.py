"""CPU tests of the host side: C-ABI surface (every symbol declared in include/*.h is exported, no
compute calls), the facade's scene recipes packed on the host, and their golden oracle outputs."""
import ctypes
import os
import re

import numpy as np
import pytest

from kuafu_b200 import build, wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(%s[A-Za-z0-9]+)\s*\(" % prefix, text)))


def test_kfrt_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(build.lib_path("libkfrt.so"))
    names = _declared("kf_rt.h", "kfrt")
    assert len(names) >= 27
    for n in names:
        assert hasattr(lib, n), f"{n} declared in kf_rt.h but not exported"


def test_kuafu_c_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(build.lib_path("libkuafu.so"))
    names = _declared("kuafu_c.h", "kfc")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in kuafu_c.h but not exported"


def test_no_cpu_fallback_without_gpu(built):
    """The product path must fail loudly when no device is usable."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kuafu_b200 import host, rt
    with pytest.raises(rt.KfrtError) as e:
        rt.Context(0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    r = host.Renderer(device=None)
    r.load_scene("cornell", 16, 16, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback|host-only"):
        r.run()


def test_wire_dtypes_match_header_sizes():
    assert wire.VERTEX.itemsize == 48 and wire.MATERIAL.itemsize == 80 and wire.INSTANCE.itemsize == 80
    assert wire.CAMERA.itemsize == 320 and wire.PUSH_CONSTANTS.itemsize == 48
    assert wire.POINT_LIGHTS.itemsize == 1024 and wire.ACTIVE_LIGHTS.itemsize == 1536


EXPECTED = {
    # recipe: (triangles, instances, cameras)  -- SURVEY.md §8.5
    "spheres": (277444, 10, 1),
    "active": (277432, 9, 2),
    "million": (999602, 205, 1),
    "articulated": (10035202, 2049, 64),
}


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_recipe_sizes(built, name):
    from kuafu_b200 import host
    r = host.Renderer(device=None)
    ncam = r.load_scene(name)
    ws = r.wire_scene()
    tris, insts, cams = EXPECTED[name]
    assert (ws.n_tris(), len(ws.insts), ncam) == (tris, insts, cams)
    r.close()


def test_camera_packing_matches_reference_formulas(built):
    """CameraUBO as reference camera.cpp:92-124 / scene.cpp:236-250 fill it (checked against numpy)."""
    import pyscene
    from kuafu_b200 import host
    r = host.Renderer(device=None)
    r.load_scene("spheres", 800, 600, 4)
    ws = r.wire_scene()
    ref = pyscene.camera([-12.6, 0.0, 15.4], [0.67, 0.0, -0.8], [0, 0, 1], 800, 600)
    for k in ("view", "projection", "viewInverse", "projectionInverse"):
        assert np.allclose(ws.cams[0][k], ref[k], rtol=1e-5, atol=1e-5), k
    assert np.allclose(ws.cams[0]["position"], [-12.6, 0.0, 15.4, 0.0])
    assert np.allclose(ws.cams[0]["front"], [0.67, 0.0, -0.8, 5.0])
    pc = ws.pc
    assert (int(pc["sampleRatePerPixel"]), int(pc["maxPathDepth"]), int(pc["russianRoulette"])) == (4, 8, 0)
    assert int(pc["frameCount"]) == -1  # accumulation off -> frameCount stays -1 (context.cpp:346-349)
    d = ws.dl["direction"]
    assert np.allclose(d[:3], np.array([-2, -1, -1]) / np.sqrt(6), atol=1e-6) and d[3] == 0.5


def test_active_light_packing(built):
    import pyscene
    from kuafu_b200 import host
    r = host.Renderer(device=None)
    r.load_scene("active", 160, 90, 1)
    ws = r.wire_scene()
    view = pyscene.look_at([-3.0, -3.0, 8.0], [0, 0, 0], [-1.0, 0.5, 0])
    vinv = np.linalg.inv(view)
    al = ws.al
    assert np.allclose(al["viewMat"][0], pyscene.col_major(view), atol=1e-5)
    assert np.allclose(al["projMat"][0], pyscene.col_major(pyscene.perspective(np.radians(150.0), 1.0, 0.01, 1000.0)), atol=1e-4)
    assert np.allclose(al["front"][0], [-vinv[0, 2], -vinv[1, 2], -vinv[2, 2], 1], atol=1e-5)
    assert np.allclose(al["position"][0][:3], vinv[:3, 3], atol=1e-4)
    assert al["sftp"][0][2] == 0 and len(ws.textures) == 1  # the projector pattern is texture 0
    assert al["front"][1][3] == 0  # unused slots are switched off


@pytest.mark.parametrize("case", ["spheres", "cornell", "million", "active", "articulated"])
def test_config_golden(built, case):
    """Facade packing + oracle reproduce the frozen fixtures (integer outputs exactly)."""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden
    g = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    out = make_golden.render_case(case)
    for k in ("hit_ids", "hit_t_bits", "depth_bits", "counters", "n_tris"):
        assert np.array_equal(out[k], g[k]), k
    assert np.allclose(out["sum"], g["sum"], rtol=1e-5, atol=1e-6)
    assert (np.abs(out["bgra"].astype(int) - g["bgra"].astype(int)) > 1).mean() < 1e-3


def test_header_structs_match_the_python_mirrors(tmp_path):
    """include/kf_rt.h compiled as C: the sizes and field offsets of the structs that cross the ABI equal those of
    the numpy mirrors in kuafu_b200/wire.py (the mirrors are what the tests and bench.py read counters and
    statistics through)."""
    import shutil
    import subprocess
    import numpy as np
    from kuafu_b200 import wire
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"KfrtCounters": wire.COUNTERS, "KfrtBvhStats": wire.BVH_STATS, "KfrtInstance": wire.INSTANCE,
               "KfrtCamera": wire.CAMERA, "KfrtPushConstants": wire.PUSH_CONSTANTS, "KfrtVertex": wire.VERTEX,
               "KfrtMaterial": wire.MATERIAL}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "kf_rt.h"', 'int main(void) {']
    by_field = ("KfrtCounters", "KfrtBvhStats")  # mirrored field by field (the others pad under other names)
    for name, dt in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field in (dt.names if name in by_field else ()):
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([cc, "-std=c99", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, dt in structs.items():
        assert int(got[name]) == dt.itemsize, (name, got[name], dt.itemsize)
        for field in (dt.names if name in by_field else ()):
            assert int(got[f"{name}.{field}"]) == dt.fields[field][1], (name, field)
