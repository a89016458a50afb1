"""The C-ABI library loads on a CPU-only box and exports every symbol include/viml.h declares."""
import ctypes as C
import os
import re

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "viml.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(viml_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(pkg):
    lib = pkg.load_library()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in viml.h but not exported"
    assert sorted(pkg.ABI_SYMBOLS) == syms
    assert lib.viml_abi_version() == 5


def test_struct_sizes_match_header(pkg):
    # sizes implied by the C declarations (LP64): catches drift between viml.h and the ctypes mirror
    abi = pkg._abi
    assert C.sizeof(abi.Config) == 4 * 8 + 2 * 4 + 12 * 8 + 5 * 8
    assert C.sizeof(abi.WindowBatch) == 16 + 3 * 8 + 8 + 4 * 8 + 8 + 3 * 8 + 6 * 8
    assert C.sizeof(abi.LinearizeOut) == 14 * 8
    assert C.sizeof(abi.MargBatch) == 16 + 8 + 16
    assert C.sizeof(abi.MargOut) == 32
    assert C.sizeof(abi.AssocQuery) == 8 + 7 * 8
    assert C.sizeof(abi.AssocOut) == 5 * 8 + 8 + 8
    assert C.sizeof(abi.DenseFactors) == 8 + 8 + 7 * 8
    assert C.sizeof(abi.TriangulateIn) == 8 + 2 * 8 + 8 + 4 * 8
    assert C.sizeof(abi.ReducedOut) == 16 and C.sizeof(abi.GnOptions) == 16 and C.sizeof(abi.GnOut) == 7 * 8


def test_no_cpu_fallback(pkg, cfg):
    """Without a usable device viml_create must fail (VIML_ERR_NO_DEVICE) — there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        return
    lib = pkg.load_library()
    h = C.c_void_p()
    assert lib.viml_create(C.byref(h), C.byref(cfg), 0) == pkg._abi.VIML_ERR_NO_DEVICE
    try:
        pkg.Context(cfg)
        raise AssertionError("Context() must raise without a GPU")
    except pkg.VimlError:
        pass


def test_product_never_imports_oracle():
    """Product sources never reference the oracle; the only mention is build.py linking the C++ SELF-TEST
    (host/tests/, test infrastructure) against it.  The shipped libraries must not depend on it either."""
    import subprocess
    pkgdir = os.path.join(ROOT, "tc-viml_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and "tests" not in dp and f != "build.py":
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in src and "viml_oracle" not in src and "import oracle" not in src, f
    for lib in ("libviml_b200.so", "libviml_host.so"):
        path = os.path.join(pkgdir, lib)
        if os.path.exists(path):
            assert "oracle" not in subprocess.run(["ldd", path], capture_output=True, text=True).stdout
