"""CPU oracle of viml_reduced_system / viml_gn_step (TEST INFRASTRUCTURE ONLY — tests/, smoke(), bench.py cpu_baseline).

The per-factor arithmetic is the C++ oracle's (oracle/viml_oracle.cpp via oracle.linearize_batch: Evaluate, loss correction,
ThreadsConstructA rule, landmark Schur complement); the steps on top are restated here with numpy:
  * dense-block factors (prior, IMU) added by the ThreadsConstructA rule, marginalization_factor.cpp:141-172: A += J^T J, b += J^T r
  * one solver iteration as ceres::Solve performs it on the reduced camera system (estimator.cpp:1888-1905): solve
    (S + lambda diag S) dx = -g, back-substitute the landmarks, PoseLocalParameterization::Plus
    (factor/pose_local_parameterization.cpp:3-19: p + dp, (q * deltaQ(dtheta)).normalized(); Utility::deltaQ utility/utility.h:16-28)
  * cost = 1/2 sum rho(|r|^2), Ceres CauchyLoss rho(s) = a^2 log(1 + s / a^2).
Parity unpinned (the reference holds no vectors for ceres::Solve internals).
"""
import numpy as np

from . import oracle as orc


def _abi():
    return orc._abi()


def reduced_system(cfg, batch, dense, flags, nthreads=8):
    abi = _abi()
    f = (flags & abi.LOSS_CAUCHY) | abi.OUT_HB | abi.OUT_SCHUR
    ref = orc.linearize_batch(cfg, batch, f, nthreads=nthreads)
    W, D = batch.W, batch.D
    X = dense.X if dense is not None else 0
    Dx = D + X
    Sx, gx = np.zeros((W, Dx, Dx)), np.zeros((W, Dx))
    Sx[:, :D, :D] = ref["S"]
    gx[:, :D] = ref["g"]
    if dense is not None:
        for (w, r, J, ci) in dense.factors:
            r, J, ci = np.asarray(r, dtype=np.float64), np.asarray(J, dtype=np.float64), np.asarray(ci)
            Sx[w][np.ix_(ci, ci)] += J.T @ J
            gx[w][ci] += J.T @ r
    return Sx, gx, ref


def pose_plus(x, d):
    p = x[:3] + d[:3]
    ax, ay, az, aw = x[3:7]
    bx, by, bz, bw = 0.5 * d[3], 0.5 * d[4], 0.5 * d[5], 1.0
    w = aw * bw - ax * bx - ay * by - az * bz
    q = np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                  aw * bz + az * bw + ax * by - ay * bx, w])
    return np.concatenate([p, q / np.sqrt((q * q).sum())])


def window_cost(cfg, batch, flags, nthreads=8):
    """1/2 sum rho(|r|^2) of the point and line factors per window, from the RAW residuals."""
    abi = _abi()
    o = orc.linearize_batch(cfg, batch, abi.OUT_RESIDUAL_JACOBIAN, nthreads=nthreads)
    a2 = cfg.cauchy_a * cfg.cauchy_a
    cost = np.zeros(batch.W)
    for res, off in ((o["pf_residual"], batch.pf_window_offset), (o["lf_residual"], batch.lf_window_offset)):
        s = (res * res).sum(axis=1)
        rho = a2 * np.log1p(s / a2) if (flags & abi.LOSS_CAUCHY) else s
        cs = np.concatenate([[0.0], np.cumsum(rho)])
        cost += cs[off[1:]] - cs[off[:-1]]
    return 0.5 * cost


def gn_step(cfg, batch, dense, extra, flags, lam=0.0, nthreads=8):
    abi = _abi()
    Sx, gx, ref = reduced_system(cfg, batch, dense, flags, nthreads)
    W, P, F, D = batch.W, batch.P, batch.F, batch.D
    X = dense.X if dense is not None else 0
    Dx = D + X
    out = {"poses": batch.poses.copy(), "ex_pose": batch.ex_pose.copy(), "inv_depth": batch.inv_depth.copy(),
           "extra": np.zeros((W, X)) if extra is None else np.array(extra, dtype=np.float64).reshape(W, X).copy(),
           "dx": np.zeros((W, Dx)), "cost": np.zeros((W, 3)), "solved": np.zeros(W, dtype=np.int32)}
    out["cost"][:, 0] = window_cost(cfg, batch, flags, nthreads)
    for w in range(W):
        A = 0.5 * (Sx[w] + Sx[w].T)
        Ad = A + lam * np.diag(np.diag(A))
        try:
            Lc = np.linalg.cholesky(Ad)
        except np.linalg.LinAlgError:
            continue
        y = np.linalg.solve(Lc, -gx[w])
        dx = np.linalg.solve(Lc.T, y)
        out["solved"][w] = 1
        out["dx"][w] = dx
        out["cost"][w, 2] = -gx[w] @ dx - 0.5 * dx @ (Sx[w] @ dx)
        for p in range(P):
            out["poses"][w, p] = pose_plus(batch.poses[w, p], dx[6 * p:6 * p + 6])
        out["ex_pose"][w] = pose_plus(batch.ex_pose[w], dx[6 * P:6 * P + 6])
        hll = ref["H_ll"][w]
        dl = np.where(hll > 1e-8, -(ref["b_l"][w] + ref["H_lp"][w] @ dx[:D]) / np.where(hll > 1e-8, hll, 1.0), 0.0)
        out["inv_depth"][w] = batch.inv_depth[w] + dl
        out["extra"][w] += dx[D:]
    nb = abi.Batch(out["poses"], out["ex_pose"], out["inv_depth"], batch.pf_window_offset, batch.pf_idx, batch.pf_obs,
                   batch.lf_window_offset, batch.lf_frame, batch.lf_geom, batch.pf_pts_i_z)
    out["cost"][:, 1] = window_cost(cfg, nb, flags, nthreads)
    if dense is not None:
        for (w, r, J, ci) in dense.factors:
            r, J, ci = np.asarray(r, dtype=np.float64), np.asarray(J, dtype=np.float64), np.asarray(ci)
            out["cost"][w, 0] += 0.5 * (r @ r)
            v = r + J @ out["dx"][w][ci]
            out["cost"][w, 1] += 0.5 * (v @ v)
    return out
