import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    out = {}
    for name in ("weights", "shared_n3000_b3", "primitives", "per_pair", "training_grads"):
        out[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return out


@pytest.fixture(scope="session")
def golden_lift():
    return dict(np.load(os.path.join(GOLDEN, "resblock3d.npz")))


def _ensure_built():
    """Tests may run before the driver's build(): compile the library in-tree if it is missing
    (test convenience only - the product loader itself fails loudly on a missing library)."""
    lib = os.path.join(ROOT, "3dahv_b200", "lib3dahv_b200.so")
    if not os.path.exists(lib):
        import importlib.util

        spec = importlib.util.spec_from_file_location("ahv_build", os.path.join(ROOT, "3dahv_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


@pytest.fixture(scope="session")
def ahv():
    """The product package (loads lib3dahv_b200.so lazily)."""
    _ensure_built()
    return importlib.import_module("3dahv_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import ahv_oracle

    return ahv_oracle
