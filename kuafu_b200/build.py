"""In-tree build of the native libraries (explicit nvcc / g++ invocations; no JIT cache).

    python -m kuafu_b200.build            # everything that is out of date
    python -m kuafu_b200.build --force

Outputs (git-ignored, but they travel to the GPU box with the snapshot):
    kuafu_b200/lib/libkfrt.so     CUDA core + C ABI (include/kf_rt.h), sm_100a
    kuafu_b200/lib/libkuafu.so    C++ host facade (kuafu.hpp API) + C shim, links libkfrt
    oracle/libkf_oracle.so        CPU oracle (test infrastructure only)
    oracle/_ref/libkf_ref.so      the reference's shaders compiled for the CPU (test infrastructure only;
                                  built only where /root/reference exists)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kuafu_b200")
# KFRT_LIB_DIR: tooling hook for A/B runs of kernel variants (a directory holding both libraries)
LIB = os.environ.get("KFRT_LIB_DIR") or os.path.join(PKG, "lib")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-pthread",
             "-fvisibility=hidden"]
# The image's $CXX wrapper (/opt/gcc/bin/g++) resolves -lstdc++ to a static archive: a second copy of
# the C++ runtime inside the library, exported, which interposes on the process's libstdc++.so (seen
# as a crash in iostream code under Python).  Link the shared runtime by name instead.
CXX_LIBS = ["-nostdlib++", "-l:libstdc++.so.6"]


def lib_path(name):
    return os.path.join(LIB, name)


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return sorted(out)


def _run(cmd, cwd=None):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=cwd)


def build_kfrt(force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(PKG, "csrc", "kf_rt.cu")
    deps = _sources(os.path.join(PKG, "csrc"), (".cu", ".cuh")) + [os.path.join(ROOT, "include", "kf_rt.h")]
    out = lib_path("libkfrt.so")
    os.makedirs(LIB, exist_ok=True)
    if force or _newer(out, deps):
        _run([nvcc] + NVCC_FLAGS + ["-o", out, src])
    return out


def build_host(force=False):
    host = os.path.join(PKG, "host")
    srcs = _sources(os.path.join(host, "src"), (".cpp",))
    if not srcs:
        return None
    deps = srcs + _sources(os.path.join(host, "include"), (".hpp", ".h")) + [os.path.join(ROOT, "include", "kf_rt.h")]
    out = lib_path("libkuafu.so")
    if force or _newer(out, deps):
        cxx = os.environ.get("CXX", "g++")
        _run([cxx] + CXX_FLAGS + ["-I", os.path.join(host, "include"), "-I", os.path.join(ROOT, "include"),
                                  "-o", out] + srcs + ["-L", LIB, "-lkfrt", "-Wl,-rpath,$ORIGIN", "-lz"] + CXX_LIBS)
    return out


def build_oracle(force=False):
    d = os.path.join(ROOT, "oracle")
    out = os.path.join(d, "libkf_oracle.so")
    if force or _newer(out, [os.path.join(d, "kf_oracle.cpp"), os.path.join(d, "Makefile")]):
        if force and os.path.exists(out):
            os.remove(out)
        _run(["make", "-C", d])
    return out


REFERENCE = os.environ.get("KUAFU_REFERENCE", "/root/reference")


def build_ref(force=False):
    """oracle/_ref/libkf_ref.so: the reference's shaders compiled for the CPU from where they lie (test
    infrastructure).  Only where the reference tree exists; elsewhere the prebuilt file is used."""
    d = os.path.join(ROOT, "oracle")
    out = os.path.join(d, "_ref", "libkf_ref.so")
    if not os.path.isdir(os.path.join(REFERENCE, "resources", "shaders")):
        return out if os.path.exists(out) else None
    if force and os.path.exists(out):
        os.remove(out)
    _run(["make", "-C", d, "ref", f"REFERENCE={REFERENCE}"])
    return out


def build_all(force=False, oracle=True):
    outs = [build_kfrt(force), build_host(force)]
    if oracle:
        outs.append(build_oracle(force))
        outs.append(build_ref(force))
    return [o for o in outs if o]


if __name__ == "__main__":
    for o in build_all(force="--force" in sys.argv):
        print(o)
