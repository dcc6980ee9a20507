import importlib
import json
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "zksnark-finalproject_b200"       # hyphenated directory: imported through importlib


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")
    config.addinivalue_line("markers", "slow: longer CPU-side oracle checks")


@pytest.fixture(scope="session")
def b2z():
    """The product package (loads libb200zk.so lazily; building it if absent)."""
    pkg = importlib.import_module(PKG)
    if not os.path.exists(pkg._ffi.LIB_PATH):
        importlib.import_module(PKG + ".build").build_library()
    return pkg


@pytest.fixture(scope="session")
def circuits():
    return importlib.import_module(PKG + ".circuits")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ctx(b2z):
    """A device context; GPU tests only.  Fails (not skips) when CUDA is missing:
    the product has no CPU fallback and a silent skip would hide that."""
    c = b2z.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def cpu_oracle():
    from oracle import cpu_oracle as m
    m.lib()
    m.set_threads(min(16, m.hardware_threads()))
    return m
