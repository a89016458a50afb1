"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (tc-viml_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("viml_oracle.cpp", "viml_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src if os.path.exists(s))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_cos_threshold.restype = C.c_double
        _LIB.orc_cos_threshold.argtypes = [C.c_double]
    return _LIB


def _abi():
    import sys
    return sys.modules["tc_viml_b200"]._abi if "tc_viml_b200" in sys.modules else __import__("__graft_entry__").load_package()._abi


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def projection_evaluate(pts_i, pts_j, sqrt_info, pose_i, pose_j, ex, inv_dep, want_jac=True):
    """Returns (r[2], [J_i 2x7, J_j 2x7, J_ex 2x7, J_l 2x1])."""
    pts_i, pts_j = np.asarray(pts_i, dtype=np.float64), np.asarray(pts_j, dtype=np.float64)
    params = [np.ascontiguousarray(x, dtype=np.float64) for x in (pose_i, pose_j, ex, np.atleast_1d(inv_dep))]
    pp = (C.POINTER(C.c_double) * 4)(*[_dp(p) for p in params])
    r = np.zeros(2)
    Js = [np.zeros((2, 7)), np.zeros((2, 7)), np.zeros((2, 7)), np.zeros((2, 1))]
    jp = (C.POINTER(C.c_double) * 4)(*[_dp(j) for j in Js]) if want_jac else None
    lib().orc_projection_evaluate(_dp(pts_i), _dp(pts_j), C.c_double(sqrt_info), pp, _dp(r), jp)
    return r, Js


def line_evaluate(ps, pe, abc, K, bcR, bcT, pose, want_jac=True):
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (ps, pe, abc, np.asarray(K).reshape(-1), np.asarray(bcR).reshape(-1), bcT)]
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    pp = (C.POINTER(C.c_double) * 1)(_dp(pose))
    r, J = np.zeros(2), np.zeros((2, 7))
    jp = (C.POINTER(C.c_double) * 1)(_dp(J)) if want_jac else None
    lib().orc_line_evaluate(*[_dp(a) for a in arrs], pp, _dp(r), jp)
    return r, J


def linearize_batch(cfg, batch, flags, nthreads=1, out=None):
    """Oracle counterpart of Context.linearize: returns dict of numpy outputs (`out`: reuse buffers of an earlier call, so
    that a timed loop does not pay for the allocation)."""
    bufs = batch.alloc_out(flags, fill=0.0) if out is None else out
    s = batch.struct()
    o = _abi().out_struct(bufs)
    rc = lib().orc_linearize_batch(C.byref(cfg), C.byref(s), C.byref(o), C.c_uint32(flags), C.c_int(nthreads))
    assert rc == 0, rc
    return bufs


def window_dense(cfg, batch, w, flags):
    pos = batch.D + batch.F
    A, b = np.zeros((pos, pos)), np.zeros(pos)
    s = batch.struct()
    rc = lib().orc_window_dense(C.byref(cfg), C.byref(s), C.c_int(w), C.c_uint32(flags), _dp(A), _dp(b))
    assert rc == 0, rc
    return A, b


def marginalize_dense(A, b, m, eps=1e-8):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    pos = A.shape[0]
    n = pos - m
    As, bs, lj, lr = np.zeros((n, n)), np.zeros(n), np.zeros((n, n)), np.zeros(n)
    rc = lib().orc_marginalize_dense(_dp(A), _dp(b), C.c_int(pos), C.c_int(m), C.c_double(eps), _dp(As), _dp(bs), _dp(lj), _dp(lr))
    assert rc == 0, rc
    return As, bs, lj, lr


def sym_eig(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    w, v = np.zeros(n), np.zeros((n, n))
    lib().orc_sym_eig(_dp(a), C.c_int(n), _dp(w), _dp(v))
    return w, v


def marginalization_factor_evaluate(n, m, keep_size, keep_idx, keep_data, lin_jac, lin_res, params, want_jac=True):
    nblk = len(keep_size)
    ks = np.asarray(keep_size, dtype=np.int32)
    ki = np.asarray(keep_idx, dtype=np.int32)
    kd = [np.ascontiguousarray(x, dtype=np.float64) for x in keep_data]
    pr = [np.ascontiguousarray(x, dtype=np.float64) for x in params]
    kdp = (C.POINTER(C.c_double) * nblk)(*[_dp(x) for x in kd])
    prp = (C.POINTER(C.c_double) * nblk)(*[_dp(x) for x in pr])
    lin_jac = np.ascontiguousarray(lin_jac, dtype=np.float64)
    lin_res = np.ascontiguousarray(lin_res, dtype=np.float64)
    r = np.zeros(n)
    Js = [np.full((n, int(s)), np.nan) for s in ks]
    jp = (C.POINTER(C.c_double) * nblk)(*[_dp(j) for j in Js]) if want_jac else None
    lib().orc_marginalization_factor_evaluate(C.c_int(n), C.c_int(m), C.c_int(nblk), ks.ctypes.data_as(C.POINTER(C.c_int)),
                                              ki.ctypes.data_as(C.POINTER(C.c_int)), kdp, _dp(lin_jac), _dp(lin_res), prp, _dp(r), jp)
    return r, Js


def update_lines_in_fov(cfg, pose, ex, lines):
    lines = np.ascontiguousarray(lines, dtype=np.float64)
    pose, ex = np.ascontiguousarray(pose, dtype=np.float64), np.ascontiguousarray(ex, dtype=np.float64)
    out = np.zeros(len(lines), dtype=np.int32)
    n = lib().orc_update_lines_in_fov(C.byref(cfg), _dp(pose), _dp(ex), _dp(lines), C.c_int64(len(lines)),
                                      out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out[:n].copy()


def line_correspondence(cfg, pose, ex, lines, fov, line2d):
    lines = np.ascontiguousarray(lines, dtype=np.float64)
    fov = np.ascontiguousarray(fov, dtype=np.int32)
    pose, ex = np.ascontiguousarray(pose, dtype=np.float64), np.ascontiguousarray(ex, dtype=np.float64)
    l2 = np.ascontiguousarray(line2d, dtype=np.float64)
    err, proj = np.zeros(3, dtype=np.float32), np.zeros(4)
    idx = lib().orc_line_correspondence(C.byref(cfg), _dp(pose), _dp(ex), _dp(lines), fov.ctypes.data_as(C.POINTER(C.c_int32)),
                                        C.c_int(len(fov)), _dp(l2), err.ctypes.data_as(C.POINTER(C.c_float)), _dp(proj))
    return idx, err, proj


def line_associate(cfg, lines, cull_poses, match_poses, ex_pose, lines2d, n_lines2d=None, fov_capacity=0, want_mask=False, nthreads=1,
                   cull_ex_pose=None):
    """Oracle counterpart of Context.associate: dict(match_index, err, projected, fov_count, fov_index, fov_mask)."""
    abi = _abi()
    lines = np.ascontiguousarray(lines, dtype=np.float64)
    Pq, L = lines2d.shape[0], lines2d.shape[1]
    q = abi.AssocQuery()
    q.n_poses, q.lines_per_pose = Pq, L
    keep = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None for x in (cull_poses, match_poses, ex_pose, lines2d)]
    q.cull_poses, q.match_poses, q.ex_pose, q.lines2d = [abi.ptr(k) for k in keep]
    nl = None if n_lines2d is None else np.ascontiguousarray(n_lines2d, dtype=np.int32)
    q.n_lines2d = abi.ptr(nl)
    cex = None if cull_ex_pose is None else np.ascontiguousarray(cull_ex_pose, dtype=np.float64)
    q.cull_ex_pose = abi.ptr(cex)
    N = len(lines)
    res = {"match_index": np.full((Pq, L), -2, dtype=np.int32), "err": np.full((Pq, L, 3), np.nan, dtype=np.float32),
           "projected": np.full((Pq, L, 4), np.nan), "fov_count": np.zeros(Pq, dtype=np.int32)}
    if fov_capacity:
        res["fov_index"] = np.full((Pq, fov_capacity), -1, dtype=np.int32)
    if want_mask:
        res["fov_mask"] = np.zeros((Pq, (N + 31) // 32), dtype=np.uint32)
    o = abi.AssocOut()
    o.match_index, o.err, o.projected, o.fov_count = [abi.ptr(res[k]) for k in ("match_index", "err", "projected", "fov_count")]
    o.fov_index, o.fov_capacity, o.fov_mask = abi.ptr(res.get("fov_index")), fov_capacity, abi.ptr(res.get("fov_mask"))
    rc = lib().orc_line_associate(C.byref(cfg), _dp(lines), C.c_int64(N), C.byref(q), C.byref(o), C.c_int(nthreads))
    assert rc == 0, rc
    return res


def track_gate(line_vecs):
    lv = np.ascontiguousarray(line_vecs, dtype=np.float64).reshape(-1, 3)
    cred = np.zeros(len(lv), dtype=np.uint8)
    cm = lib().orc_track_gate(C.c_int(len(lv)), _dp(lv), cred.ctypes.data_as(C.POINTER(C.c_uint8)))
    return bool(cm), cred.astype(bool)


def cos_threshold(angle_th):
    return float(lib().orc_cos_threshold(C.c_double(angle_th)))


def hardware_threads():
    return int(lib().orc_hardware_threads())
