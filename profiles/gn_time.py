"""Times the solver-iteration kernels (K_GN event pairs) of viml_gn_step on 1024 EuRoC-shaped windows with the realistic dense-factor
set, tiled Cholesky vs the column-by-column kernel (VIML_GN_UNBLOCKED=1), and prints the difference of the two steps.
usage (on a GPU box, from the repo root): python profiles/gn_time.py"""
import importlib, sys, time, json, os
sys.path.insert(0, ".")
import numpy as np
pkg = importlib.import_module("tc-viml_b200")
abi, synth = pkg._abi, pkg.synth
cfg = synth.euroc_config()
b = synth.make_windows(1024, seed=9)
dense = synth.make_dense_factors(b, seed=5)
extra = np.zeros((b.W, dense.X))
with pkg.Context(cfg) as c:
    for env in ("", "1"):
        if env: os.environ["VIML_GN_UNBLOCKED"] = env
        o = c.gn_step(b, dense, extra, abi.LOSS_CAUCHY, lam=1e-4)
        c.profile_begin()
        for _ in range(3): o = c.gn_step(b, dense, extra, abi.LOSS_CAUCHY, lam=1e-4)
        pr = c.profile_end()
        print("unblocked" if env else "tiled", "K_GN total per step ms", pr["gn_step"][0]/3, o["solved"].mean())
        if not env: ref = o
    print("max diff dx", float(np.abs(o["dx"]-ref["dx"]).max()), float(np.abs(ref["dx"]).max()), float(np.abs(o["cost"]-ref["cost"]).max()))
