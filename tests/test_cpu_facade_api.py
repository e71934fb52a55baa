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
