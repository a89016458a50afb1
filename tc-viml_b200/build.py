"""In-tree build of the native artefacts (no JIT cache: the .so files travel with the repo snapshot).

  libviml_b200.so   csrc/*.cu   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
  libviml_host.so   host/*.cpp  g++   (reference-interface shim over the C-ABI)
  host_selftest     host/tests/selftest.cpp  (C++ tests written like the reference's would be)
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
HOST = os.path.join(_HERE, "host")
INCLUDE = os.path.join(ROOT, "include")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC,
               "--expt-relaxed-constexpr"]
# Translation units whose results must be bit-exact against the oracle's un-fused arithmetic.
EXACT_TUS = {"associate_kernels.cu"}


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError(f"build step failed: {' '.join(cmd[:3])} ...")
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def _newer(srcs, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build_cuda(verbose=False, force=False, ptxas_info=False):
    nvcc = _nvcc()
    cus = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(INCLUDE, "viml.h"))
    objs = []
    os.makedirs(os.path.join(_HERE, "build"), exist_ok=True)
    for cu in cus:
        src = os.path.join(CSRC, cu)
        obj = os.path.join(_HERE, "build", cu[:-3] + ".o")
        objs.append(obj)
        if force or _newer([src] + hdrs, obj):
            extra = ["-fmad=false"] if cu in EXACT_TUS else []
            extra += os.environ.get("VIML_NVCC_EXTRA", "").split()
            if ptxas_info:
                extra += ["-Xptxas", "-v"]
            _run([nvcc] + NVCC_ARCH + NVCC_COMMON + extra + ["-c", src, "-o", obj], verbose)
    so = os.path.join(_HERE, "libviml_b200.so")
    if force or _newer(objs, so):
        _run([nvcc] + NVCC_ARCH + ["-shared", "--cudart", "static", "-o", so] + objs + ["-ldl"], verbose)
    return so


def build_host(verbose=False, force=False):
    cxx = shutil.which("g++") or "g++"
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + [os.path.join(INCLUDE, "viml.h")]
    so = os.path.join(_HERE, "libviml_host.so")
    out = [so]
    if srcs and (force or _newer(srcs + hdrs, so)):
        _run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", INCLUDE, "-I", HOST, "-o", so] + srcs +
             ["-L", _HERE, "-lviml_b200", "-Wl,-rpath,$ORIGIN"], verbose)
    tdir = os.path.join(HOST, "tests")
    if os.path.isdir(tdir):
        for f in sorted(os.listdir(tdir)):
            if f.endswith(".cpp"):
                exe = os.path.join(_HERE, "build", f[:-4])
                src = os.path.join(tdir, f)
                if force or _newer([src] + srcs + hdrs, exe):
                    _run([cxx, "-O2", "-std=c++17", "-Wall", "-I", INCLUDE, "-I", HOST, "-I", ROOT, "-o", exe, src,
                          "-L", _HERE, "-lviml_host", "-lviml_b200", "-L", os.path.join(ROOT, "oracle"), "-loracle",
                          f"-Wl,-rpath,{_HERE}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}", "-pthread"], verbose)
                out.append(exe)
    return out


def build_all(verbose=False, force=False):
    so = build_cuda(verbose=verbose, force=force)
    # the selftest links the oracle as its checker: make sure it exists (building it is not using it)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    build_host(verbose=verbose, force=force)
    return so


if __name__ == "__main__":
    build_all(verbose=True, force="--force" in sys.argv)
