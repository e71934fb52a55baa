"""kuafu_b200 -- B200-native path-tracing core behind the kuafu.hpp API.

The product is the C-ABI library ``lib/libkfrt.so`` (``include/kf_rt.h``) built from
``csrc/*.cu`` for sm_100a, plus the C++ host facade in ``host/``.  This Python package only holds
thin plumbing: the build recipe, ctypes bindings and numpy views of the wire structs.
"""
from . import wire  # noqa: F401
from .build import build_all, lib_path  # noqa: F401

__all__ = ["wire", "build_all", "lib_path"]
