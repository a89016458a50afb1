"""Generates the committed golden fixtures.  Run HERE (build container) only: it reads the reference's real
data fixtures under /root/reference, which do not exist on the GPU box.

  assoc_euroc_<room>.npz   the shipped prior line map (benchmark_publisher/config/V?_0?/line_3d.txt), the
                           thresholds and map->VIO transform of that sequence's sensor.yaml, ground-truth body
                           poses from data.csv (1 Hz subsample), synthetic LSD-style 2D lines, and the
                           association results of the CPU oracle — cross-checked bit for bit against the
                           independent pure-Python restatement (oracle_py) while generating.
  linearize_cfg1.npz       one EuRoC-shaped window batch (BASELINE.json configs[0]) with the oracle's
                           residuals/Jacobians/H/b/Schur, cross-checked against 50-digit mpmath (oracle_mp).

The reference ships no golden vectors of its own; these pin the oracle against regressions
and give the GPU tests fixed inputs that were derived from the reference's real geometry.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

REF = "/root/reference/benchmark_publisher/config"


def yaml_matrix(txt, key, n):
    m = re.search(key + r":.*?\n((?:\s*#.*\n|\s+\w+:.*\n)*?)\s*data:\s*\[(.*?)\]", txt, flags=re.S)
    # take the LAST uncommented data: after the key
    seg = txt[txt.index(key + ":"):]
    datas = [d for d in re.finditer(r"^\s*data:\s*\[(.*?)\]", seg, flags=re.S | re.M)]
    vals = [float(v) for v in datas[0].group(1).replace("\n", " ").split(",")]
    assert len(vals) == n, (key, len(vals))
    return np.array(vals)


def yaml_scalar(txt, key):
    m = re.search(r"^" + key + r":\s*([-+0-9.eE]+)", txt, flags=re.M)
    return float(m.group(1))


def make_assoc(seq, out_name, n_poses=48, L=60, n_true=20, seed=11):
    pkg, orc = ge.load_package(), ge.load_oracle()
    from oracle import oracle_py
    synth, abi = pkg.synth, pkg._abi
    txt = open(os.path.join(REF, seq, "sensor.yaml")).read()
    Rbw = yaml_matrix(txt, "initialRotation", 9).reshape(3, 3)
    Tbw = yaml_matrix(txt, "initialTranslation", 3)
    Ric = yaml_matrix(txt, "extrinsicRotation", 9).reshape(3, 3)
    Tic = yaml_matrix(txt, "extrinsicTranslation", 3)
    cfg = abi.make_config(fx=yaml_scalar(txt, "fx"), fy=yaml_scalar(txt, "fy"), cx=yaml_scalar(txt, "cx"),
                          cy=yaml_scalar(txt, "cy"), width=int(yaml_scalar(txt, "width")),
                          height=int(yaml_scalar(txt, "height")), Rbw=Rbw, Tbw=Tbw,
                          overlap_th=yaml_scalar(txt, "overlap_th"), dist_th=yaml_scalar(txt, "dist_th"),
                          angle_th=yaml_scalar(txt, "angle_th"))
    lines = np.loadtxt(os.path.join(REF, seq, "line_3d.txt")).reshape(-1, 6)
    gt = np.loadtxt(os.path.join(REF, seq, "data.csv"), delimiter=",", comments="#")
    step = max(len(gt) // n_poses, 1)
    gt = gt[::step][:n_poses]
    rng = np.random.Generator(np.random.PCG64(seed))
    # GT body pose in the map frame (p, q = w,x,y,z) -> VIO world via Rbw/Tbw (estimator.cpp:402-403)
    pm, qw = gt[:, 1:4], gt[:, 4:8]
    Rm = np.stack([synth._rot_from_quat(np.array([q[1], q[2], q[3], q[0]]) / np.linalg.norm(q)) for q in qw])
    Rv = Rbw @ Rm
    pv = pm @ Rbw.T + Tbw
    cull = synth.pose7(pv, Rv, rng)
    match = synth.pose7(pv + 0.01 * rng.standard_normal(pv.shape),
                        Rv @ synth.rot_from_axis_angle(rng.standard_normal(pv.shape) * 0.002), rng)
    ex = synth.pose7(np.broadcast_to(Tic, pv.shape).copy(), np.broadcast_to(Ric, Rv.shape).copy(), rng)
    W, H = cfg.width, cfg.height
    l2d = np.empty((n_poses, L, 4))
    for p in range(n_poses):
        q = cull[p, 3:7] / np.linalg.norm(cull[p, 3:7])
        qe = ex[p, 3:7] / np.linalg.norm(ex[p, 3:7])
        Rbi, Rc = synth._rot_from_quat(q), synth._rot_from_quat(qe)
        R = Rc.T @ Rbi.T @ Rbw
        T = Rc.T @ (Rbi.T @ (Tbw - cull[p, :3]) - ex[p, :3])
        ps, pe = lines[:, :3] @ R.T + T, lines[:, 3:] @ R.T + T
        with np.errstate(all="ignore"):
            uv = np.stack([cfg.fx * ps[:, 0] / ps[:, 2] + cfg.cx, cfg.fy * ps[:, 1] / ps[:, 2] + cfg.cy,
                           cfg.fx * pe[:, 0] / pe[:, 2] + cfg.cx, cfg.fy * pe[:, 1] / pe[:, 2] + cfg.cy], -1)
        ok = (ps[:, 2] > 0) & (pe[:, 2] > 0) & (uv[:, 0] > 1) & (uv[:, 0] < W - 2) & (uv[:, 2] > 1) & (uv[:, 2] < W - 2) & \
            (uv[:, 1] > 1) & (uv[:, 1] < H - 2) & (uv[:, 3] > 1) & (uv[:, 3] < H - 2)
        ok &= np.hypot(uv[:, 2] - uv[:, 0], uv[:, 3] - uv[:, 1]) > 30
        vis = np.nonzero(ok)[0]
        k = min(n_true, len(vis))
        out = np.empty((L, 4))
        if k:
            seg = uv[rng.choice(vis, size=k, replace=False)]
            a0, a1 = rng.uniform(0, 0.2, (k, 1)), rng.uniform(0.8, 1.0, (k, 1))
            out[:k, :2] = seg[:, :2] + a0 * (seg[:, 2:] - seg[:, :2])
            out[:k, 2:] = seg[:, :2] + a1 * (seg[:, 2:] - seg[:, :2])
            out[:k] += rng.standard_normal((k, 4))
        u0, v0 = rng.uniform(0, W, L - k), rng.uniform(0, H, L - k)
        ang, ln = rng.uniform(0, 2 * np.pi, L - k), rng.uniform(100, 400, L - k)
        out[k:] = np.stack([u0, v0, np.clip(u0 + ln * np.cos(ang), 0, W - 1), np.clip(v0 + ln * np.sin(ang), 0, H - 1)], -1)
        l2d[p] = out[rng.permutation(L)]
    # a few degenerate queries: zero-length, exactly vertical/horizontal, duplicate of another
    l2d[0, 0] = [100, 100, 100, 100]
    l2d[0, 1] = [200, 50, 200, 300]
    l2d[0, 2] = [50, 240, 600, 240]
    l2d = l2d.astype(np.float32).astype(np.float64)
    res = orc.line_associate(cfg, lines, cull, match, ex, l2d, fov_capacity=512, want_mask=True)
    # cross-check against the pure-Python restatement, bit for bit
    nchk = 0
    for p in range(n_poses):
        fl = oracle_py.fov(cfg, cull[p].tolist(), ex[p].tolist(), lines.tolist())
        assert fl == res["fov_index"][p, :res["fov_count"][p]].tolist(), ("fov", p)
        for l in range(L):
            idx, err, pl = oracle_py.correspondence(cfg, match[p].tolist(), ex[p].tolist(), lines.tolist(), fl, l2d[p, l])
            assert idx == res["match_index"][p, l], ("idx", p, l, idx, res["match_index"][p, l])
            assert np.array_equal(np.array(err, dtype=np.float32), res["err"][p, l]), ("err", p, l, err, res["err"][p, l])
            if idx >= 0:
                assert np.array_equal(np.array(pl), res["projected"][p, l]), ("proj", p, l)
                nchk += 1
    print(f"{out_name}: {len(lines)} map lines, fov median {int(np.median(res['fov_count']))}, "
          f"{nchk}/{n_poses * L} queries matched; oracle == oracle_py bit for bit")
    cfgv = np.array([cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.width, cfg.height, cfg.overlap_th, cfg.dist_th, cfg.angle_th])
    np.savez_compressed(os.path.join(HERE, out_name), cfg=cfgv, Rbw=Rbw, Tbw=Tbw, lines=lines, cull=cull, match=match,
                        ex=ex, lines2d=l2d, match_index=res["match_index"], err=res["err"],
                        projected=np.nan_to_num(res["projected"], nan=0.0), fov_count=res["fov_count"],
                        fov_index=res["fov_index"], source=np.array(f"{seq}: line_3d.txt, sensor.yaml, data.csv"))


def make_linearize(out_name="linearize_cfg1.npz"):
    pkg, orc = ge.load_package(), ge.load_oracle()
    from oracle import oracle_mp
    import mpmath as mp
    synth, abi = pkg.synth, pkg._abi
    cfg = synth.euroc_config()
    b = synth.make_windows(2, seed=0x5EED + 1)
    flags = abi.OUT_RESIDUAL_JACOBIAN | abi.OUT_HB | abi.OUT_SCHUR | abi.LOSS_CAUCHY
    out = orc.linearize_batch(cfg, b, flags)
    raw = orc.linearize_batch(cfg, b, abi.OUT_RESIDUAL_JACOBIAN)
    # mpmath cross-check of a sample of factors (raw and loss-corrected)
    worst = 0.0
    for k in list(range(0, b.NP, 97)):
        w = int(np.searchsorted(b.pf_window_offset, k, side="right") - 1)
        pk = int(b.pf_idx[k])
        i, j, f = pk & 255, (pk >> 8) & 255, pk >> 16
        r, J = oracle_mp.projection([b.pf_obs[k, 0], b.pf_obs[k, 1], 1.0], [b.pf_obs[k, 2], b.pf_obs[k, 3], 1.0], cfg.sqrt_info,
                                    b.poses[w, i], b.poses[w, j], b.ex_pose[w], b.inv_depth[w, f])
        Jo = np.concatenate([raw["pf_jac_pose_i"][k].reshape(2, 7)[:, :6], raw["pf_jac_pose_j"][k].reshape(2, 7)[:, :6],
                             raw["pf_jac_ex"][k].reshape(2, 7)[:, :6], raw["pf_jac_feat"][k].reshape(2, 1)], 1)
        Jm = np.array([[float(J[a, c]) for c in range(19)] for a in range(2)])
        rm = np.array([float(x) for x in r])
        worst = max(worst, np.abs(Jo - Jm).max() / np.abs(Jm).max(), np.abs(raw["pf_residual"][k] - rm).max() / np.abs(rm).max())
        sc = float(oracle_mp.cauchy_scale(r, cfg.cauchy_a))
        worst = max(worst, np.abs(out["pf_residual"][k] - sc * rm).max() / np.abs(rm).max())
    for k in range(0, b.NL, 13):
        w = k // (b.NL // b.W)
        g = b.lf_geom[:, k]
        K = [cfg.fx, 0, cfg.cx, 0, cfg.fy, cfg.cy, 0, 0, 1]
        qe = b.ex_pose[w, 3:7] / np.linalg.norm(b.ex_pose[w, 3:7])
        r, J = oracle_mp.line(g[0:3], g[3:6], g[6:9], K, synth._rot_from_quat(qe).reshape(-1), b.ex_pose[w, :3],
                              b.poses[w, b.lf_frame[k]])
        Jm = np.array([[float(J[a, c]) for c in range(6)] for a in range(2)])
        rm = np.array([float(x) for x in r])
        worst = max(worst, np.abs(raw["lf_jac_pose"][k].reshape(2, 7)[:, :6] - Jm).max() / np.abs(Jm).max(),
                    np.abs(raw["lf_residual"][k] - rm).max() / np.abs(rm).max())
    print(f"{out_name}: {b.NP} point + {b.NL} line factors; oracle vs 50-digit mpmath worst rel err {worst:.2e}")
    assert worst < 1e-10
    np.savez_compressed(os.path.join(HERE, out_name), **{"in_" + k: v for k, v in b.arrays().items() if v is not None},
                        **{"out_" + k: v for k, v in out.items()})


if __name__ == "__main__":
    make_assoc("V1_02_medium", "assoc_euroc_v1.npz")
    make_assoc("V2_01_easy", "assoc_euroc_v2.npz", n_poses=32, seed=12)
    make_linearize()
