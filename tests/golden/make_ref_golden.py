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

# ---- association: the reference's own UpdateLinesInFoV / LineCorrespondenceInFrame / CalAngleDist / CalEulerDist (oracle/ref_estimator.cpp)
# on queries built ON the thresholds (sub-segments of the projected map lines whose overlap sits within 1e-9 ... 1e-3 of
# overlap_th, over-long queries, exactly / nearly vertical and (near) zero-length segments, parallel offsets, rotations by about
# angle_th), duplicated map lines (distance ties), drifting match poses, a frame-entry extrinsic, ragged line counts.
ext = (150.0, 150.0, 20.0)
lines = synth.make_line_map(2500, seed=61, extent=ext)
lines = np.concatenate([lines, lines[:100]])
Pq, L = 6, 120
cull, match, ex, l2d = synth.make_assoc_queries(lines, Pq, L=L, n_true=60, seed=62, extent=ext, pose_drift=True)
arng = np.random.Generator(np.random.PCG64(63))
aout = {"lines": lines}
cases = ((0.1745, 0.45, False), (0.18, 0.8, True), (0.3, 0.0, False), (0.1, 1.0, True))
for c, (ang, ov, extras) in enumerate(cases):
    cfg = synth.euroc_config(angle_th=ang, overlap_th=ov)
    base = ref.line_associate(cfg, lines, cull, match, ex, l2d, nthreads=8)
    q = l2d.copy()
    for p in range(Pq):
        segs = base["projected"][p][base["match_index"][p] >= 0]
        if len(segs) == 0:
            continue
        for i in range(L):
            s = segs[i % len(segs)]
            S, E = s[:2], s[2:]
            d = E - S
            Ls = np.hypot(*d)
            if not np.isfinite(Ls) or Ls < 1e-6:
                continue
            u, n = d / Ls, np.array([-d[1], d[0]]) / Ls
            eps = (0.0, 1e-9, -1e-9, 1e-6, -1e-6, 1e-4, -1e-4, 1e-3, -1e-3)[arng.integers(9)]
            f = min(max(ov + eps, 1e-3), 1.0)
            mode = i % 11
            mid = 0.5 * (S + E)
            if mode == 0:
                a, bb = S, E
            elif mode == 1:
                a, bb = S, S + f * Ls * u
            elif mode == 2:
                a = S + (1 - f) * Ls * u
                bb = a + (f * Ls + 40.0) * u * (0.5 if f * Ls + 40.0 > Ls else 1.0)
            elif mode == 3:
                e_ = Ls / f - Ls
                a, bb = S - 0.3 * e_ * u, E + 0.7 * e_ * u
            elif mode == 4:
                a, bb = mid - [0, 0.4 * Ls], mid + [0, 0.4 * Ls]
            elif mode == 5:
                a, bb = mid - [5e-5, 0.4 * Ls], mid + [5e-5, 0.4 * Ls]
            elif mode == 6:
                a, bb = mid, mid + (1e-5 * u if i % 22 == 6 else 0.0)
            elif mode == 7:
                a, bb = S + 3 * n, E + 3 * n
            elif mode == 8:
                a, bb = E.astype(np.float32), S.astype(np.float32)
            elif mode == 9:
                cs, si = np.cos(ang * (1 + eps)), np.sin(ang * (1 + eps))
                a, bb = S, S + np.array([cs * d[0] - si * d[1], si * d[0] + cs * d[1]])
            else:
                continue                                              # keep the synthetic query
            q[p, i] = [a[0], a[1], bb[0], bb[1]]
    kw = {}
    if extras:
        cex = ex.copy()
        cex[:, :3] += 0.01 * arng.standard_normal((Pq, 3))
        kw = dict(cull_ex_pose=cex, n_lines2d=arng.integers(L // 2, L + 1, Pq).astype(np.int32))
        aout[f"c{c}_cull_ex"], aout[f"c{c}_n_lines2d"] = cex, kw["n_lines2d"]
    cap = 1024
    res = ref.line_associate(cfg, lines, cull, match, ex, q, fov_capacity=cap, nthreads=8, **kw)
    assert res["fov_count"].max() <= cap
    if "n_lines2d" in kw:                                             # rows beyond a pose's count are not outputs: fixed filler
        inval = np.arange(L)[None, :] >= kw["n_lines2d"][:, None]
        res["match_index"][inval], res["err"][inval], res["projected"][inval] = -2, 0, 0
    aout.update({f"c{c}_angle_th": ang, f"c{c}_overlap_th": ov, f"c{c}_cull": cull, f"c{c}_match": match, f"c{c}_ex": ex,
                 f"c{c}_lines2d": q, f"c{c}_fov_capacity": cap})
    aout.update({f"c{c}_{k}": v for k, v in res.items()})
    print(f"ref_assoc case {c}: angle_th {ang}, overlap_th {ov}: matched {(res['match_index'] >= 0).mean():.2f}, fov median {int(np.median(res['fov_count']))}")
aout["n_cases"] = len(cases)
apath = os.path.join(ROOT, "tests", "golden", "ref_assoc.npz")
np.savez_compressed(apath, **aout)
print("wrote", apath, os.path.getsize(apath), "bytes")
