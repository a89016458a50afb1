"""Small invocation of every kernel path, for `compute-sanitizer --tool memcheck|racecheck|synccheck python tests/sanitize_smoke.py`.
Not a pytest module (sanitizer runs are slow); results are summarised in profiles/sanitizer_r<round>.md."""
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
    ctx.linearize(synth.make_windows(3, seed=1), allf)                                   # plan + fused assembly + TMA-pipelined Schur
    ctx.linearize(synth.make_windows(2, seed=2, F=300, all_start_zero=True), allf)       # many tasks per window, 10 Schur stages
    ctx.linearize(synth.make_windows(2, seed=8, P=7, F=37), allf)                                      # D < 72, odd landmark count
    ctx.linearize(synth.make_windows(2, seed=3, P=20, F=60, lines_per_frame=2), allf)    # generic atomic path + tiled Schur
    ctx.linearize(synth.make_windows(1, seed=9, P=40, F=200, lines_per_frame=2, max_len=12), abi.OUT_SCHUR | abi.S_PACKED)  # band plan + packing
    ctx.linearize(synth.make_windows(2, seed=4), abi.OUT_RESIDUAL_JACOBIAN)             # mode A kernels
    b = synth.make_windows(600, seed=5, f32_obs=True)                                    # chunked host pipeline + float32 observation table
    ctx.linearize(b, abi.OUT_SCHUR | abi.S_PACKED | abi.LOSS_CAUCHY, obs_table="f32")
    bl, lmap = synth.with_line_map(synth.make_windows(5, seed=10, f32_obs=True), cfg)      # line table: expanded from the map in HBM
    ctx.set_map(lmap)
    ctx.linearize(bl, allf, obs_table="f32", line_table=True)
    irr = synth.make_windows(3, seed=6)
    idx = irr.pf_idx.copy()
    idx[1] = (idx[1] & 0xffff0000) | 0x0101                                              # i == j: the window goes to irregular_kernel
    ctx.linearize(abi.Batch(irr.poses, irr.ex_pose, irr.inv_depth, irr.pf_window_offset, idx, irr.pf_obs, irr.lf_window_offset,
                            irr.lf_frame, irr.lf_geom), allf)
    small = synth.make_windows(4, seed=7)
    dense = synth.make_dense_factors(small, seed=5)
    ctx.gn_step(small, dense, np.zeros((small.W, dense.X)), abi.LOSS_CAUCHY, lam=1e-4)   # reduced system (DMMA tiles), tiled Cholesky, update, cost
    rng = np.random.default_rng(0)
    M = rng.standard_normal((50, 40))
    ctx.marginalize(M.T @ M, rng.standard_normal(40), 15)
    lines = synth.make_line_map(3000, seed=5, extent=(120.0, 120.0, 30.0))
    cull, match, ex, l2d = synth.make_assoc_queries(lines, 3, L=20, n_true=8, seed=6, extent=(120.0, 120.0, 30.0))
    ctx.set_map(lines)
    ctx.associate(cull, match, ex, l2d, fov_capacity=256, want_mask=True)                # few poses: 64-thread match CTAs
    cull, match, ex, l2d = synth.make_assoc_queries(lines, 130, L=40, n_true=16, seed=7, extent=(120.0, 120.0, 30.0))
    ctx.associate(cull, None, ex, l2d)                                                   # >= 128 poses: 320-thread match CTAs
    dense_map = synth.make_line_map(9000, seed=8, extent=(30.0, 30.0, 10.0))             # FoV lists longer than the match stage: chunks
    cull, match, ex, l2d = synth.make_assoc_queries(dense_map, 2, L=40, n_true=16, seed=9, extent=(30.0, 30.0, 10.0))
    ctx.set_map(dense_map)
    r = ctx.associate(cull, None, ex, l2d)
    print("longest FoV list:", int(r["fov_count"].max()))
    ctx.fov_update(0, cull[0], ex[0])
    ctx.associate(None, cull[:1], ex[:1], l2d[:1], cached=True)
    ctx.track_gate(np.array([0, 3, 5], dtype=np.int32), np.array([1, 1, 2, 7, 7], dtype=np.int32))
print("sanitize_smoke: done")
