import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Native libraries must exist in-tree (they are built here and travel to the GPU box)."""
    from kuafu_b200 import build
    if not os.path.exists(build.lib_path("libkfrt.so")):
        build.build_all()
    return build
