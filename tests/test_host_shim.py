"""The C++ host shim (tc-viml_b200/host): reference class interfaces over the C-ABI, checked by its own C++
self-test (host/tests/selftest.cpp), which builds factors / MarginalizationInfo / the line associator exactly like
Estimator::OptimizationWithLine and processImagewithLine do and compares against the oracle."""
import os
import subprocess

import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "tc-viml_b200", "build", "selftest")


def _build(pkg):
    pkg.build.build_all()
    assert os.path.exists(EXE)


def test_host_shim_builds_and_host_only_parts(pkg, orc):
    _build(pkg)
    r = subprocess.run([EXE, "--cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all passed" in r.stdout


def test_host_shim_links_only_the_c_abi(pkg):
    """libviml_host.so depends on libviml_b200.so (the C-ABI), never on the oracle."""
    _build(pkg)
    out = subprocess.run(["ldd", os.path.join(ROOT, "tc-viml_b200", "libviml_host.so")], capture_output=True, text=True).stdout
    assert "libviml_b200.so" in out and "liboracle" not in out


@pytest.mark.gpu
def test_host_shim_on_gpu(pkg, orc):
    _build(pkg)
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all passed" in r.stdout
