import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.chdir(ROOT)  # scene files use paths relative to the repository root
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The native product libraries; built in-tree if missing (nvcc cross-compiles without a GPU)."""
    if not (os.path.exists(os.path.join(ROOT, "lisa_b200", "liblisa_rt.so"))
            and os.path.exists(os.path.join(ROOT, "lisa_b200", "liblisa_host.so"))
            and os.path.exists(os.path.join(ROOT, "lisa_b200", "lisa"))):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "lisa_b200"), "all"])
    return True


@pytest.fixture(scope="session")
def orc():
    """The oracle (oracle/cpu_ref.c) — the CHECKER, never the thing under test."""
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def rt(built):
    import lisa_b200.rt as rt
    return rt


@pytest.fixture(scope="session")
def frontend(built):
    import lisa_b200.frontend as fe
    return fe


@pytest.fixture(scope="session")
def cornell(frontend):
    """README Cornell scene parsed by the PRODUCT parser (C++), geometry loaded."""
    return frontend.parse_scene("scenes/cornell_c1.rto")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def resized(sc, w, h=None, spp=None, bounces=None):
    s = dict(sc)
    s["width"], s["height"] = w, (h or w)
    if spp is not None:
        s["num_samples"] = spp
    if bounces is not None:
        s["num_bounces"] = bounces
    return s
