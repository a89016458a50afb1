"""Writes tests/golden/ref_factors.npz: input / output vectors produced by the REFERENCE's own factor code (oracle/_ref/libref.so
= /root/reference/vins_estimator/src/factor/*.cpp compiled unmodified against oracle/ref_shim/, see oracle/Makefile).
Run in the build container (the reference tree is not on the GPU box):  python tests/golden/make_ref_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from oracle import ref  # noqa: E402

assert ref.build(), "oracle/_ref/libref.so could not be built (reference tree missing?)"
synth = pkg.synth
rng = np.random.default_rng(20261017)
b = synth.make_windows(3, seed=4242)
out = {}
# ProjectionFactor::Evaluate on every factor of three windows + un-normalised quaternions
N = b.NP
win = np.searchsorted(b.pf_window_offset, np.arange(N), side="right") - 1
ii, jj, ll = b.pf_idx & 0xff, (b.pf_idx >> 8) & 0xff, b.pf_idx >> 16
poses = b.poses.copy()
poses[1, :, 3:] *= 1.0 + 0.05 * rng.standard_normal((b.P, 1))     # window 1: clearly non-unit quaternions (projection_factor.cpp:25-31)
sq = 460.0 / 1.5
pr, pJ = np.zeros((N, 2)), [np.zeros((N, 2, 7)) for _ in range(3)] + [np.zeros((N, 2, 1))]
for k in range(N):
    w = win[k]
    r, J = ref.projection_evaluate([b.pf_obs[k, 0], b.pf_obs[k, 1], 1.0], [b.pf_obs[k, 2], b.pf_obs[k, 3], 1.0], sq, poses[w, ii[k]], poses[w, jj[k]],
                                   b.ex_pose[w], b.inv_depth[w, ll[k]])
    pr[k] = r
    for q in range(4):
        pJ[q][k] = J[q]
out.update(pf_poses=poses, pf_ex=b.ex_pose, pf_inv_depth=b.inv_depth, pf_idx=b.pf_idx, pf_obs=b.pf_obs, pf_win=win, pf_sqrt_info=sq,
           pf_r=pr, pf_Ji=pJ[0], pf_Jj=pJ[1], pf_Je=pJ[2], pf_Jl=pJ[3])
# LineProjectionFactor::Evaluate on every line factor (frozen extrinsic = the window's, rotation normalised, estimator.cpp:1777-1781)
NL = b.NL
lwin = np.searchsorted(b.lf_window_offset, np.arange(NL), side="right") - 1
K = np.array([[synth.FX, 0, synth.CX], [0, synth.FY, synth.CY], [0, 0, 1.0]])
lr, lJ = np.zeros((NL, 2)), np.zeros((NL, 2, 7))
for k in range(NL):
    w = lwin[k]
    q = b.ex_pose[w, 3:7] / np.linalg.norm(b.ex_pose[w, 3:7])
    r, J = ref.line_evaluate(b.lf_geom[0:3, k], b.lf_geom[3:6, k], b.lf_geom[6:9, k], K, synth._rot_from_quat(q), b.ex_pose[w, :3],
                             b.poses[w, b.lf_frame[k]])
    lr[k], lJ[k] = r, J
out.update(lf_poses=b.poses, lf_ex=b.ex_pose, lf_frame=b.lf_frame, lf_geom=b.lf_geom, lf_win=lwin, lf_r=lr, lf_J=lJ)
# PoseLocalParameterization::Plus
xs = synth.pose7(rng.standard_normal((64, 3)), synth.rot_from_axis_angle(rng.standard_normal((64, 3))), rng)
ds = 0.05 * rng.standard_normal((64, 6))
out.update(plus_x=xs, plus_d=ds, plus_out=np.array([ref.pose_plus(x, d) for x, d in zip(xs, ds)]))
# MarginalizationInfo (MARGIN_OLD projection set): reduced system of the kept poses + extrinsic
m = synth.make_windows(2, seed=4343, P=6, F=40, all_start_zero=True, lines_per_frame=0)
for w in range(2):
    k0, k1 = m.pf_window_offset[w], m.pf_window_offset[w + 1]
    idx = m.pf_idx[k0:k1]
    for loss, tag in ((None, "noloss"), (1.0, "cauchy")):
        A, bb, mm = ref.marginalize_old(m.poses[w], m.ex_pose[w], m.inv_depth[w], (idx >> 8) & 0xff, idx >> 16, m.pf_obs[k0:k1], sq, loss)
        out[f"marg{w}_{tag}_A"], out[f"marg{w}_{tag}_b"], out[f"marg{w}_m"] = A, bb, mm
out.update(marg_poses=m.poses, marg_ex=m.ex_pose, marg_inv_depth=m.inv_depth, marg_off=m.pf_window_offset, marg_idx=m.pf_idx, marg_obs=m.pf_obs)
# Line2D::Line2D(Vector4d) and Line2D::Point2Flined (float32-rounded pixel segments, some vertical / degenerate)
segs = rng.uniform(0, 750, (400, 4)).astype(np.float32).astype(np.float64)
segs[::40, 2] = segs[::40, 0]                      # vertical: Point2Flined always clamps to an endpoint (feature_manager.cpp:60-66)
qp = rng.uniform(-50, 800, (400, 2))
out.update(l2d_seg=segs, l2d_p=qp, l2d_out=np.array([ref.line2d(s_) for s_ in segs]),
           l2d_foot=np.array([ref.point2flined(s_, p_) for s_, p_ in zip(segs, qp)]))
# FeatureManager::triangulate: the features of window 0 with consecutive observation frames
w = 0
k0, k1 = b.pf_window_offset[w], b.pf_window_offset[w + 1]
idx = b.pf_idx[k0:k1]
feat = idx >> 16
t_start, t_off, t_pts = [], [0], []
for l in np.unique(feat):
    sel = np.nonzero(feat == l)[0]
    i = int(idx[sel[0]] & 0xff)
    js = sorted(int((idx[s_] >> 8) & 0xff) for s_ in sel)
    if js != list(range(i + 1, i + 1 + len(js))):
        continue
    obs = [[b.pf_obs[k0 + sel[0], 0], b.pf_obs[k0 + sel[0], 1], 1.0]]
    for j in js:
        s_ = sel[[int((idx[t] >> 8) & 0xff) for t in sel].index(j)]
        obs.append([b.pf_obs[k0 + s_, 2], b.pf_obs[k0 + s_, 3], 1.0])
    t_start.append(i), t_pts.extend(obs), t_off.append(t_off[-1] + len(obs))
out.update(tri_poses=b.poses[w], tri_ex=b.ex_pose[w], tri_start=np.array(t_start, dtype=np.int32), tri_off=np.array(t_off, dtype=np.int64),
           tri_pts=np.array(t_pts), tri_depth=ref.triangulate(b.poses[w], b.ex_pose[w], t_start, t_off, np.array(t_pts)))
path = os.path.join(ROOT, "tests", "golden", "ref_factors.npz")
np.savez_compressed(path, **out)
print(f"{path}: {N} ProjectionFactor, {NL} LineProjectionFactor, 64 Plus, 4 marginalisations, 400 Line2D / Point2Flined, "
      f"{len(t_start)} triangulations from the reference's own code")
