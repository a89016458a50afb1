"""ctypes binding of oracle/_ref/libref.so — the REFERENCE's own factor translation units (TEST INFRASTRUCTURE ONLY).

Built by `make -C oracle ref` from the sources under /root/reference (see oracle/Makefile, oracle/ref_wrap.cpp); present in this
container and, as a prebuilt file, on the GPU box.  available() is False where it was never built."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref.so")
_LIB = None


def build():
    """Compile _ref/libref.so when the reference tree is here; a no-op otherwise (the prebuilt file is used as it is)."""
    if os.path.isdir("/root/reference/vins_estimator/src"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(_PATH)


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(_PATH)
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def projection_evaluate(pts_i, pts_j, sqrt_info, pose_i, pose_j, ex, inv_dep):
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (pose_i, pose_j, ex, np.atleast_1d(inv_dep))]
    pp = (C.POINTER(C.c_double) * 4)(*[_dp(a) for a in arrs])
    pi, pj = np.ascontiguousarray(pts_i, dtype=np.float64), np.ascontiguousarray(pts_j, dtype=np.float64)
    r = np.zeros(2)
    J = [np.zeros((2, 7)), np.zeros((2, 7)), np.zeros((2, 7)), np.zeros((2, 1))]
    jp = (C.POINTER(C.c_double) * 4)(*[_dp(j) for j in J])
    lib().ref_projection_evaluate(_dp(pi), _dp(pj), C.c_double(sqrt_info), pp, _dp(r), jp)
    return r, J


def line_evaluate(ps, pe, abc, K, bcR, bcT, pose):
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (ps, pe, abc, np.asarray(K).reshape(9), np.asarray(bcR).reshape(9), bcT)]
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    pp = (C.POINTER(C.c_double) * 1)(_dp(pose))
    r, J = np.zeros(2), np.zeros((2, 7))
    jp = (C.POINTER(C.c_double) * 1)(_dp(J))
    lib().ref_line_evaluate(*[_dp(a) for a in arrs], pp, _dp(r), jp)
    return r, J


def pose_plus(x, delta):
    x, d = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(delta, dtype=np.float64)
    o = np.zeros(7)
    lib().ref_pose_plus(_dp(x), _dp(d), _dp(o))
    return o


def marginalize_old(poses, ex, inv_depth, fj, fl, obs, sqrt_info, cauchy_a=None):
    """MarginalizationInfo on projection factors (0, j, l) with drop set {0, 3}.  Returns (A [n,n], b [n]) = (J^T J, J^T r) of the
    reference's linearized_jacobians / linearized_residuals, re-ordered to [pose 1 .. pose P-1, extrinsic] (6 tangent columns
    each), and m."""
    poses = np.ascontiguousarray(poses, dtype=np.float64).copy()
    ex = np.ascontiguousarray(ex, dtype=np.float64).copy()
    dep = np.ascontiguousarray(inv_depth, dtype=np.float64).copy()
    fj, fl = np.ascontiguousarray(fj, dtype=np.int32), np.ascontiguousarray(fl, dtype=np.int32)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    P = poses.shape[0]
    n = 6 * P
    ko, ki = np.zeros(P, dtype=np.int32), np.zeros(P, dtype=np.int32)
    lj, lr = np.zeros((n, n)), np.zeros(n)
    m = C.c_int()
    code = lib().ref_marginalize_old(P, len(dep), _dp(poses), _dp(ex), _dp(dep), len(fj), fj.ctypes.data_as(C.POINTER(C.c_int)),
                                     fl.ctypes.data_as(C.POINTER(C.c_int)), _dp(obs), C.c_double(sqrt_info), 0 if cauchy_a is None else 1,
                                     C.c_double(cauchy_a or 1.0), ko.ctypes.data_as(C.POINTER(C.c_int)), ki.ctypes.data_as(C.POINTER(C.c_int)),
                                     _dp(lj), _dp(lr), C.byref(m))
    nn, nkeep = code // 1000, code % 1000
    assert nn == n and nkeep == P, (nn, nkeep)
    lj = lj.reshape(-1)[:n * n].reshape(n, n)
    A, b = lj.T @ lj, lj.T @ lr
    perm = np.zeros(n, dtype=np.int64)   # canonical column (pose p-1 block, or last block for the extrinsic) -> reference column
    for blk in range(P):
        tgt = (ko[blk] - 1) if ko[blk] < P else P - 1
        perm[6 * tgt:6 * tgt + 6] = ki[blk] + np.arange(6)
    return A[np.ix_(perm, perm)], b[perm], int(m.value)


def line2d(seg):
    seg = np.ascontiguousarray(seg, dtype=np.float64)
    o = np.zeros(7)
    lib().ref_line2d(_dp(seg), _dp(o))
    return o


def point2flined(seg, p):
    seg, p = np.ascontiguousarray(seg, dtype=np.float64), np.ascontiguousarray(p, dtype=np.float64)
    o = np.zeros(2)
    lib().ref_point2flined(_dp(seg), _dp(p), _dp(o))
    return o


def triangulate(poses, ex, start, off, pts):
    """FeatureManager::triangulate on one window (the JacobiSVD behind it is the stand-in's one-sided Jacobi: depths agree with
    any accurate SVD to rounding times conditioning, not bit for bit)."""
    poses, ex = np.ascontiguousarray(poses, dtype=np.float64), np.ascontiguousarray(ex, dtype=np.float64)
    start = np.ascontiguousarray(start, dtype=np.int32)
    off = np.ascontiguousarray(off, dtype=np.int64)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    depth = np.zeros(len(start))
    lib().ref_triangulate(poses.shape[0], _dp(poses), _dp(ex), len(start), start.ctypes.data_as(C.POINTER(C.c_int)),
                          off.ctypes.data_as(C.POINTER(C.c_longlong)), _dp(pts), _dp(depth))
    return depth


def line_associate(cfg, lines, cull_poses, match_poses, ex_pose, lines2d, n_lines2d=None, fov_capacity=0, nthreads=1, cull_ex_pose=None):
    """The association sweep through the REFERENCE's own UpdateLinesInFoV / LineCorrespondenceInFrame / CalAngleDist / CalEulerDist
    (oracle/ref_estimator.cpp): same arguments and result keys as oracle.line_associate (no mask)."""
    lines = np.ascontiguousarray(lines, dtype=np.float64)
    Pq, L = lines2d.shape[0], lines2d.shape[1]
    keep = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None
            for x in (cull_poses, match_poses, ex_pose, cull_ex_pose, lines2d)]
    nl = None if n_lines2d is None else np.ascontiguousarray(n_lines2d, dtype=np.int32)
    res = {"match_index": np.full((Pq, L), -2, dtype=np.int32), "err": np.full((Pq, L, 3), np.nan, dtype=np.float32),
           "projected": np.full((Pq, L, 4), np.nan), "fov_count": np.zeros(Pq, dtype=np.int32)}
    if fov_capacity:
        res["fov_index"] = np.full((Pq, fov_capacity), -1, dtype=np.int32)

    def p(a, t=C.c_double):
        return None if a is None else a.ctypes.data_as(C.POINTER(t))
    f = lib().ref_line_associate
    f.restype = C.c_int
    rc = f(C.byref(cfg), p(lines), C.c_int64(len(lines)), C.c_int(Pq), C.c_int(L), p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]),
           p(keep[4]), p(nl, C.c_int32), p(res["fov_count"], C.c_int32), p(res.get("fov_index"), C.c_int32), C.c_int(fov_capacity),
           p(res["match_index"], C.c_int32), p(res["err"], C.c_float), p(res["projected"]), C.c_int(nthreads))
    assert rc == 0, "a FoV list or a chosen line could not be mapped back to map indices"
    return res
