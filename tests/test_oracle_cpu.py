"""CPU tests that pin the oracle: the reference's own numeric-diff convention (ProjectionFactor::check),
an independent 50-digit mpmath derivation, an independent pure-Python association, numpy linear algebra,
and the committed golden fixtures derived from the reference's real maps/poses."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err

GOLD = os.path.join(ROOT, "tests", "golden")


def qmul(a, b):  # (x,y,z,w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def factor_inputs(b, k):
    w = int(np.searchsorted(b.pf_window_offset, k, side="right") - 1)
    pk = int(b.pf_idx[k])
    i, j, f = pk & 255, (pk >> 8) & 255, pk >> 16
    return (np.array([b.pf_obs[k, 0], b.pf_obs[k, 1], 1.0]), np.array([b.pf_obs[k, 2], b.pf_obs[k, 3], 1.0]),
            b.poses[w, i].copy(), b.poses[w, j].copy(), b.ex_pose[w].copy(), float(b.inv_depth[w, f]))


def test_projection_numeric_diff_like_reference_check(pkg, orc, cfg):
    """ProjectionFactor::check (projection_factor.cpp:126-228): forward differences, eps 1e-6,
    Q <- Q * deltaQ(d), column order [Pi,Qi,Pj,Qj,tic,qic,inv_dep]."""
    b = pkg.synth.make_windows(3, seed=21)
    eps = 1e-6
    for k in range(0, b.NP, 211):
        pts_i, pts_j, pi, pj, ex, lam = factor_inputs(b, k)
        r, Js = orc.projection_evaluate(pts_i, pts_j, cfg.sqrt_info, pi, pj, ex, lam)
        ana = np.concatenate([Js[0][:, :6], Js[1][:, :6], Js[2][:, :6], Js[3]], 1)
        assert np.all(Js[0][:, 6] == 0) and np.all(Js[1][:, 6] == 0) and np.all(Js[2][:, 6] == 0)
        num = np.zeros((2, 19))
        for c in range(19):
            P, ll = [pi.copy(), pj.copy(), ex.copy()], lam
            a, bb = divmod(c, 6)
            if a < 3:
                if bb < 3:
                    P[a][bb] += eps
                else:
                    d = np.zeros(3)
                    d[bb - 3] = eps
                    P[a][3:7] = qmul(P[a][3:7], np.array([d[0] / 2, d[1] / 2, d[2] / 2, 1.0]))
            else:
                ll = lam + eps
            r2, _ = orc.projection_evaluate(pts_i, pts_j, cfg.sqrt_info, P[0], P[1], P[2], ll, want_jac=False)
            num[:, c] = (r2 - r) / eps
        assert np.abs(num - ana).max() < 2e-5 * max(1.0, np.abs(ana).max())


def test_factors_against_mpmath(pkg, orc, cfg):
    from oracle import oracle_mp
    synth = pkg.synth
    b = synth.make_windows(2, seed=22)
    for k in range(0, b.NP, 149):
        pts_i, pts_j, pi, pj, ex, lam = factor_inputs(b, k)
        r, Js = orc.projection_evaluate(pts_i, pts_j, cfg.sqrt_info, pi, pj, ex, lam)
        rm, Jm = oracle_mp.projection(pts_i, pts_j, cfg.sqrt_info, pi, pj, ex, lam)
        ana = np.concatenate([Js[0][:, :6], Js[1][:, :6], Js[2][:, :6], Js[3]], 1)
        Jn = np.array([[float(Jm[a, c]) for c in range(19)] for a in range(2)])
        assert rel_err(ana, Jn) < 1e-11
        assert rel_err(r, np.array([float(x) for x in rm])) < 1e-10
    K = [cfg.fx, 0, cfg.cx, 0, cfg.fy, cfg.cy, 0, 0, 1]
    for k in range(0, b.NL, 17):
        w = k // (b.NL // b.W)
        g = b.lf_geom[:, k]
        qe = b.ex_pose[w, 3:7] / np.linalg.norm(b.ex_pose[w, 3:7])
        bcR = synth._rot_from_quat(qe).reshape(-1)
        pose = b.poses[w, b.lf_frame[k]]
        r, J = orc.line_evaluate(g[0:3], g[3:6], g[6:9], K, bcR, b.ex_pose[w, :3], pose)
        rm, Jm = oracle_mp.line(g[0:3], g[3:6], g[6:9], K, bcR, b.ex_pose[w, :3], pose)
        assert rel_err(J[:, :6], np.array([[float(Jm[a, c]) for c in range(6)] for a in range(2)])) < 1e-10
        assert rel_err(r, np.array([float(x) for x in rm])) < 1e-10
        assert np.all(J[:, 6] == 0)
        # the residual is the point-to-line distance |Au+Bv+C|/sqrt(A^2+B^2) >= 0
        assert np.all(r >= 0)


def test_cauchy_correction_is_uniform_scaling(pkg, orc, cfg):
    """CauchyLoss(1): rho'' < 0 always, so r and J are both scaled by 1/sqrt(1+|r|^2)."""
    abi = pkg._abi
    b = pkg.synth.make_windows(2, seed=23)
    raw = orc.linearize_batch(cfg, b, abi.OUT_RESIDUAL_JACOBIAN)
    cor = orc.linearize_batch(cfg, b, abi.OUT_RESIDUAL_JACOBIAN | abi.LOSS_CAUCHY)
    s = 1.0 / np.sqrt(1.0 + (raw["pf_residual"] ** 2).sum(1))
    assert rel_err(cor["pf_residual"], raw["pf_residual"] * s[:, None]) < 1e-15
    assert rel_err(cor["pf_jac_pose_j"], raw["pf_jac_pose_j"] * s[:, None]) < 1e-15
    s = 1.0 / np.sqrt(1.0 + (raw["lf_residual"] ** 2).sum(1))
    assert rel_err(cor["lf_jac_pose"], raw["lf_jac_pose"] * s[:, None]) < 1e-15


def dense_from_blocks(o, w, D, F):
    A = np.zeros((D + F, D + F))
    A[:D, :D] = o["H_pp"][w]
    A[D:, :D] = o["H_lp"][w]
    A[:D, D:] = o["H_lp"][w].T
    A[D:, D:] = np.diag(o["H_ll"][w])
    return A, np.concatenate([o["b_p"][w], o["b_l"][w]])


def test_hb_is_jtj_and_schur_matches_dense_marginalize(pkg, orc, cfg):
    abi = pkg._abi
    b = pkg.synth.make_windows(2, seed=24, F=40)
    flags = abi.OUT_RESIDUAL_JACOBIAN | abi.OUT_HB | abi.OUT_SCHUR | abi.LOSS_CAUCHY
    o = orc.linearize_batch(cfg, b, flags)
    D, F = b.D, b.F
    for w in range(b.W):
        A, bb = orc.window_dense(cfg, b, w, flags)
        A2, b2 = dense_from_blocks(o, w, D, F)
        assert np.array_equal(A, A2) and np.array_equal(bb, b2)
        assert np.abs(A - A.T).max() == 0.0
        # independent J^T J from the mode-A outputs
        J = np.zeros((2 * (b.NP + b.NL), D + F))
        r = np.zeros(2 * (b.NP + b.NL))
        row = 0
        for k in range(b.pf_window_offset[w], b.pf_window_offset[w + 1]):
            pk = int(b.pf_idx[k])
            i, j, f = pk & 255, (pk >> 8) & 255, pk >> 16
            J[row:row + 2, 6 * i:6 * i + 6] = o["pf_jac_pose_i"][k].reshape(2, 7)[:, :6]
            J[row:row + 2, 6 * j:6 * j + 6] = o["pf_jac_pose_j"][k].reshape(2, 7)[:, :6]
            J[row:row + 2, 6 * b.P:6 * b.P + 6] = o["pf_jac_ex"][k].reshape(2, 7)[:, :6]
            J[row:row + 2, D + f] = o["pf_jac_feat"][k]
            r[row:row + 2] = o["pf_residual"][k]
            row += 2
        for k in range(b.lf_window_offset[w], b.lf_window_offset[w + 1]):
            fr = b.lf_frame[k]
            J[row:row + 2, 6 * fr:6 * fr + 6] = o["lf_jac_pose"][k].reshape(2, 7)[:, :6]
            r[row:row + 2] = o["lf_residual"][k]
            row += 2
        assert rel_err(A, J.T @ J) < 1e-13 and rel_err(bb, J.T @ r) < 1e-13   # b = +J^T r
        # landmark Schur == the reference's dense pseudo-inverse marginalisation with landmarks first
        perm = np.concatenate([np.arange(D, D + F), np.arange(D)])
        As, bs, _, _ = orc.marginalize_dense(A[np.ix_(perm, perm)], bb[perm], F)
        assert rel_err(o["S"][w], As) < 1e-10 and rel_err(o["g"][w], bs) < 1e-10


def test_sym_eig_and_marginalize_against_numpy(orc):
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 40, 91):
        M = rng.standard_normal((n, n))
        S = M @ M.T
        w, v = orc.sym_eig(S)
        wn = np.linalg.eigvalsh(S)
        assert rel_err(w, wn) < 1e-12
        assert rel_err(v @ np.diag(w) @ v.T, S) < 1e-12 and rel_err(v.T @ v, np.eye(n)) < 1e-12
    pos, m = 60, 21
    M = rng.standard_normal((pos + 5, pos))
    A = M.T @ M
    b = rng.standard_normal(pos)
    As, bs, lj, lr = orc.marginalize_dense(A, b, m)
    Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
    ref = A[m:, m:] - A[m:, :m] @ np.linalg.pinv(Amm, hermitian=True) @ A[:m, m:]
    assert rel_err(As, ref) < 1e-10
    assert rel_err(lj.T @ lj, As) < 1e-10          # linearized_jacobians^T J == A
    assert rel_err(lj.T @ lr, bs) < 1e-9           # J^T r == b
    # rank-deficient marginalised block: eigenvalues <= eps are dropped (marginalization_factor.cpp:272)
    A2 = A.copy()
    A2[0, :] = 0
    A2[:, 0] = 0
    As2, _, _, _ = orc.marginalize_dense(A2, b, m)
    Amm2 = 0.5 * (A2[:m, :m] + A2[:m, :m].T)
    ref2 = A2[m:, m:] - A2[m:, :m] @ np.linalg.pinv(Amm2, rcond=1e-14, hermitian=True) @ A2[:m, m:]
    assert rel_err(As2, ref2) < 1e-9


def test_marginalization_factor(orc):
    rng = np.random.default_rng(6)
    sizes, m = [7, 9, 7, 1], 15
    n = sum(6 if s == 7 else s for s in sizes)
    idx, pos = [], m
    for s in sizes:
        idx.append(pos)
        pos += 6 if s == 7 else s
    x0 = [rng.standard_normal(s) for s in sizes]
    for k, s in enumerate(sizes):
        if s == 7:
            x0[k][3:7] /= np.linalg.norm(x0[k][3:7])
    lj, lr = rng.standard_normal((n, n)), rng.standard_normal(n)
    r, Js = orc.marginalization_factor_evaluate(n, m, sizes, idx, x0, lj, lr, x0)
    assert rel_err(r, lr) < 1e-15                  # dx = 0 at the linearisation point
    for k, s in enumerate(sizes):
        loc = 6 if s == 7 else s
        assert np.array_equal(Js[k][:, :loc], lj[:, idx[k] - m: idx[k] - m + loc])
        if s == 7:
            assert np.all(Js[k][:, 6] == 0)
    # small perturbation: r = r0 + J dx with dx_rot = 2 vec(q0^-1 q)
    x = [v.copy() for v in x0]
    x[0][:3] += 1e-3
    dq = np.array([1e-3, -2e-3, 5e-4, 1.0])
    x[0][3:7] = qmul(x0[0][3:7], dq)
    r2, _ = orc.marginalization_factor_evaluate(n, m, sizes, idx, x0, lj, lr, x, want_jac=False)
    dx = np.zeros(n)
    dx[0:3] = 1e-3
    dx[3:6] = 2 * dq[:3]
    assert rel_err(r2, lr + lj @ dx) < 1e-12
    # w < 0 branch flips the sign (marginalization_factor.cpp:361-364)
    x[0][3:7] = -x[0][3:7]
    r3, _ = orc.marginalization_factor_evaluate(n, m, sizes, idx, x0, lj, lr, x, want_jac=False)
    assert rel_err(r3, r2) < 1e-12


def test_association_against_pure_python(pkg, orc, cfg):
    from oracle import oracle_py
    synth = pkg.synth
    lines = synth.make_line_map(6000, seed=31, extent=(200.0, 200.0, 30.0))
    cull, match, ex, l2d = synth.make_assoc_queries(lines, 3, L=24, n_true=10, seed=32, extent=(200.0, 200.0, 30.0))
    res = orc.line_associate(cfg, lines, cull, match, ex, l2d, fov_capacity=4096)
    ll = lines.tolist()
    n_match = n_clip = 0
    for p in range(3):
        fl = oracle_py.fov(cfg, cull[p].tolist(), ex[p].tolist(), ll)
        assert fl == res["fov_index"][p, :res["fov_count"][p]].tolist()
        for l in range(24):
            idx, err, pl = oracle_py.correspondence(cfg, match[p].tolist(), ex[p].tolist(), ll, fl, l2d[p, l])
            assert idx == res["match_index"][p, l]
            assert np.array_equal(np.array(err, dtype=np.float32), res["err"][p, l])
            if idx >= 0:
                n_match += 1
                assert np.array_equal(np.array(pl), res["projected"][p, l])
                n_clip += any(float(v) != float(np.float32(v)) for v in pl)   # a clipped endpoint stays double
    assert n_match >= 20


def test_association_edge_cases(pkg, orc, cfg):
    synth = pkg.synth
    lines = synth.make_line_map(500, seed=33, extent=(60.0, 60.0, 30.0))
    cull, match, ex, l2d = synth.make_assoc_queries(lines, 2, L=8, n_true=4, seed=34, extent=(60.0, 60.0, 30.0))
    # empty map / empty FoV list: every query unmatched with (-1,-1,-1) (estimator.cpp:703-713)
    res = orc.line_associate(cfg, np.zeros((0, 6)), cull, match, ex, l2d)
    assert np.all(res["match_index"] == -1) and np.all(res["err"] == -1) and np.all(res["fov_count"] == 0)
    # degenerate (zero-length) detected line: NaN direction -> acos NaN -> PI -> rejected
    l2d[0, 0] = [100, 100, 100, 100]
    res = orc.line_associate(cfg, lines, cull, match, ex, l2d)
    assert res["match_index"][0, 0] == -1
    # ragged counts: queries beyond n_lines2d are left untouched
    res = orc.line_associate(cfg, lines, cull, match, ex, l2d, n_lines2d=[3, 0])
    assert np.all(res["match_index"][0, 3:] == -2) and np.all(res["match_index"][1] == -2)


def test_cos_threshold_is_exact_decision_boundary(orc):
    import math
    for th in (0.1745, 0.1, 0.18, 0.3491, 1e-3, 1.2):
        c = orc.cos_threshold(th)
        assert math.acos(c) <= th and math.acos(np.nextafter(c, 0.0)) > th
    assert orc.cos_threshold(2.0) == 0.0 and orc.cos_threshold(-1.0) == math.inf


def test_track_gate(orc):
    """removeLineOutlier (feature_manager.cpp:494-541): diff > 0.1 m marks the observation; the track is dropped
    only when (count / n) >= 0.5 in INTEGER division, i.e. when every observation is incredible."""
    v = np.array([[1.0, 0, 0], [1.05, 0, 0], [2.0, 0, 0], [1.0, 0.2, 0]])
    ok, cred = orc.track_gate(v)
    assert ok and cred.tolist() == [True, True, False, False]
    ok, cred = orc.track_gate(np.array([[1.0, 0, 0]]))
    assert ok and cred.tolist() == [True]          # first observation compares with itself


@pytest.mark.parametrize("name", ["assoc_euroc_v1.npz", "assoc_euroc_v2.npz"])
def test_golden_association_on_reference_maps(pkg, orc, name):
    """The shipped prior line maps + GT poses of the reference (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, name))
    c = g["cfg"]
    cfg = pkg._abi.make_config(fx=c[0], fy=c[1], cx=c[2], cy=c[3], width=int(c[4]), height=int(c[5]), Rbw=g["Rbw"],
                               Tbw=g["Tbw"], overlap_th=c[6], dist_th=c[7], angle_th=c[8])
    res = orc.line_associate(cfg, g["lines"], g["cull"], g["match"], g["ex"], g["lines2d"], fov_capacity=512)
    assert np.array_equal(res["match_index"], g["match_index"])
    assert np.array_equal(res["err"], g["err"])
    assert np.array_equal(res["fov_count"], g["fov_count"]) and np.array_equal(res["fov_index"], g["fov_index"])
    m = g["match_index"] >= 0
    assert np.array_equal(res["projected"][m], g["projected"][m])
    assert 30 <= np.median(g["fov_count"]) <= 260      # list size 38..256 on the real map


def test_golden_linearize(pkg, orc, cfg):
    g = np.load(os.path.join(GOLD, "linearize_cfg1.npz"))
    abi = pkg._abi
    b = abi.Batch(g["in_poses"], g["in_ex_pose"], g["in_inv_depth"], g["in_pf_window_offset"], g["in_pf_idx"],
                  g["in_pf_obs"], g["in_lf_window_offset"], g["in_lf_frame"], g["in_lf_geom"])
    flags = abi.OUT_RESIDUAL_JACOBIAN | abi.OUT_HB | abi.OUT_SCHUR | abi.LOSS_CAUCHY
    o = orc.linearize_batch(cfg, b, flags)
    for k, v in o.items():
        assert np.array_equal(v, g["out_" + k]), k
