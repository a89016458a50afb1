"""Small invocation of every kernel path, for `compute-sanitizer --tool memcheck|racecheck python tests/sanitize_smoke.py`.
Not a pytest module (sanitizer runs are slow); results are summarised in profiles/sanitizer_r1.md."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

pkg = ge.load_package()
abi, synth = pkg._abi, pkg.synth
cfg = synth.euroc_config()
allf = abi.OUT_RESIDUAL_JACOBIAN | abi.OUT_HB | abi.OUT_SCHUR | abi.LOSS_CAUCHY
with pkg.Context(cfg) as ctx:
    ctx.linearize(synth.make_windows(3, seed=1), allf)                                   # fused path, 1 part
    ctx.linearize(synth.make_windows(2, seed=2, F=160, all_start_zero=True), allf)       # fused path, 2 parts (RED adds)
    ctx.linearize(synth.make_windows(2, seed=3, P=20, F=60, lines_per_frame=2), allf)    # generic atomic path + tiled Schur
    ctx.linearize(synth.make_windows(2, seed=4), abi.OUT_RESIDUAL_JACOBIAN)             # mode A kernels
    rng = np.random.default_rng(0)
    M = rng.standard_normal((50, 40))
    ctx.marginalize(M.T @ M, rng.standard_normal(40), 15)
    lines = synth.make_line_map(3000, seed=5, extent=(120.0, 120.0, 30.0))
    cull, match, ex, l2d = synth.make_assoc_queries(lines, 3, L=20, n_true=8, seed=6, extent=(120.0, 120.0, 30.0))
    ctx.set_map(lines)
    ctx.associate(cull, match, ex, l2d, fov_capacity=256, want_mask=True)                # few poses: 64-thread match CTAs
    cull, match, ex, l2d = synth.make_assoc_queries(lines, 130, L=40, n_true=16, seed=7, extent=(120.0, 120.0, 30.0))
    ctx.associate(cull, None, ex, l2d)                                                   # >= 128 poses: 320-thread match CTAs
print("sanitize_smoke: done")
