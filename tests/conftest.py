import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return ge.load_package()


@pytest.fixture(scope="session")
def orc(pkg):
    o = ge.load_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def cfg(pkg):
    return pkg.synth.euroc_config()


@pytest.fixture(scope="session")
def ctx(pkg, cfg):
    """One CUDA context for the GPU tests; fails loudly when the library or the device is missing."""
    c = pkg.Context(cfg, device=0)
    yield c
    c.close()


def rel_err(got, ref):
    """max |got-ref| / max(|ref|) over the array: the 1e-9-relative bar of BASELINE.json is applied
    block-wise (entries that are structurally tiny are judged against the block's scale)."""
    import numpy as np
    scale = max(float(np.abs(ref).max()) if ref.size else 0.0, 1e-300)
    return float(np.abs(got - ref).max()) / scale if ref.size else 0.0
