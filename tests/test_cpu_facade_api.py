"""The C++ facade's API behaviour (limits, errors, ownership: SURVEY.md §8.3), host-only: compiles
tests/cxx/facade_api.cpp against kuafu_b200/host/include + libkuafu.so and runs it."""
import os
import subprocess

from kuafu_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_facade_api_through_the_cxx_headers(built, tmp_path):
    exe = str(tmp_path / "facade_api")
    lib = os.path.dirname(build.lib_path("libkuafu.so"))
    cxx = os.environ.get("CXX", "g++")
    subprocess.check_call([cxx, "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cxx", "facade_api.cpp"),
                           "-I", os.path.join(ROOT, "kuafu_b200", "host", "include"), "-I", os.path.join(ROOT, "include"),
                           "-L", lib, "-lkuafu", "-lkfrt", f"-Wl,-rpath,{lib}", "-o", exe] + build.CXX_LIBS)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "facade api ok" in out.stdout, out.stdout + out.stderr


REF_GLM = "/root/reference/3rd_party/KTX-Software/other_include"


def test_facade_builds_and_runs_against_real_glm(built, tmp_path):
    """Drop-in proof for a SAPIEN build, which has real glm: `glm::vec3` of glm 0.9.9 mangles differently
    from the in-tree stand-in (kfglm.hpp), so the library a SAPIEN tree links is one compiled with
    -DKUAFU_USE_SYSTEM_GLM.  Where /root/reference is mounted, every facade source and the API program are
    compiled against the glm the reference vendors, linked, and run (host-only: no device work)."""
    import pytest
    if not os.path.isdir(os.path.join(REF_GLM, "glm")):
        pytest.skip("the reference's vendored glm is not mounted on this box")
    cxx = os.environ.get("CXX", "g++")
    host = os.path.join(ROOT, "kuafu_b200", "host")
    srcs = sorted(os.path.join(host, "src", f) for f in os.listdir(os.path.join(host, "src")) if f.endswith(".cpp"))
    libdir = str(tmp_path)
    kfrt = os.path.dirname(build.lib_path("libkfrt.so"))
    so = os.path.join(libdir, "libkuafu_glm.so")
    common = ["-std=c++17", "-O1", "-DKUAFU_USE_SYSTEM_GLM", "-I", REF_GLM, "-I", os.path.join(host, "include"),
              "-I", os.path.join(ROOT, "include")]
    subprocess.check_call([cxx] + common + ["-fPIC", "-shared", "-pthread", "-fvisibility=hidden", "-o", so] + srcs +
                          ["-L", kfrt, "-lkfrt", f"-Wl,-rpath,{kfrt}", "-lz"] + build.CXX_LIBS)
    exe = str(tmp_path / "facade_api_glm")
    subprocess.check_call([cxx] + common + [os.path.join(ROOT, "tests", "cxx", "facade_api.cpp"), "-L", libdir,
                                            "-lkuafu_glm", "-L", kfrt, "-lkfrt", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{kfrt}",
                                            "-o", exe] + build.CXX_LIBS)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "facade api ok" in out.stdout, out.stdout + out.stderr
    # the real-glm build exports the real-glm signatures (what a SAPIEN object file would look for)
    syms = subprocess.run(["nm", "-DC", so], capture_output=True, text=True).stdout
    assert "kuafu::Camera::setPosition(glm::vec<3, float, (glm::" in syms  # glm's own template, not the stand-in struct
    mine = subprocess.run(["nm", "-DC", build.lib_path("libkuafu.so")], capture_output=True, text=True).stdout
    assert "kuafu::Camera::setPosition(glm::vec<3, float" not in mine and "kuafu::Camera::setPosition(" in mine
