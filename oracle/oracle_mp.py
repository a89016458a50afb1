"""Independent 50-digit re-derivation (mpmath) of the factor formulas.  TEST INFRASTRUCTURE ONLY.

Purpose: validate the C++ restatement (viml_oracle.cpp) — and measure the true error of the CPU and GPU
paths — without sharing any code or evaluation order with it.  Written from the reference's formulas
(projection_factor.cpp:21-124, line_projection_factor.cpp:19-120, marginalization_factor.cpp:37-68) with
matrix algebra, not statement by statement.
"""
import mpmath as mp

mp.mp.dps = 50


def _R(q):
    """Eigen toRotationMatrix of an UN-normalised quaternion given as pose-layout (x,y,z,w)."""
    x, y, z, w = [mp.mpf(float(v)) for v in q]
    return mp.matrix([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _Rinv(q):
    """Rotation map of q.inverse() = conj(q)/|q|^2."""
    x, y, z, w = [mp.mpf(float(v)) for v in q]
    n2 = x * x + y * y + z * z + w * w
    return _Rq(-x / n2, -y / n2, -z / n2, w / n2)


def _Rq(x, y, z, w):
    return mp.matrix([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _Rn(q):
    x, y, z, w = [mp.mpf(float(v)) for v in q]
    n = mp.sqrt(x * x + y * y + z * z + w * w)
    return _Rq(x / n, y / n, z / n, w / n)


def _v(a):
    return mp.matrix([mp.mpf(float(t)) for t in a])


def _skew(v):
    return mp.matrix([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def projection(pts_i, pts_j, sqrt_info, pose_i, pose_j, ex, inv_dep):
    """Returns (r [2], J 2x19 in local columns [pose_i 6 | pose_j 6 | ex 6 | inv_dep 1])."""
    Pi, Pj, tic = _v(pose_i[:3]), _v(pose_j[:3]), _v(ex[:3])
    Ri, Rj, ric = _R(pose_i[3:7]), _R(pose_j[3:7]), _R(ex[3:7])
    lam = mp.mpf(float(inv_dep))
    s = mp.mpf(float(sqrt_info))
    pi3 = _v(pts_i)
    pc_i = pi3 / lam
    p_imu_i = ric * pc_i + tic
    p_w = Ri * p_imu_i + Pi
    p_imu_j = _Rinv(pose_j[3:7]) * (p_w - Pj)
    pc_j = _Rinv(ex[3:7]) * (p_imu_j - tic)
    dep = pc_j[2]
    r = [s * (pc_j[0] / dep - mp.mpf(float(pts_j[0]))), s * (pc_j[1] / dep - mp.mpf(float(pts_j[1])))]
    reduce = s * mp.matrix([[1 / dep, 0, -pc_j[0] / dep ** 2], [0, 1 / dep, -pc_j[1] / dep ** 2]])
    J = mp.zeros(2, 19)
    ji = mp.zeros(3, 6)
    ji[:, 0:3] = ric.T * Rj.T
    ji[:, 3:6] = ric.T * Rj.T * Ri * (-_skew(p_imu_i))
    J[:, 0:6] = reduce * ji
    jj = mp.zeros(3, 6)
    jj[:, 0:3] = ric.T * (-Rj.T)
    jj[:, 3:6] = ric.T * _skew(p_imu_j)
    J[:, 6:12] = reduce * jj
    je = mp.zeros(3, 6)
    je[:, 0:3] = ric.T * (Rj.T * Ri - mp.eye(3))
    T = ric.T * Rj.T * Ri * ric
    je[:, 3:6] = -T * _skew(pc_i) + _skew(T * pc_i) + _skew(ric.T * (Rj.T * (Ri * tic + Pi - Pj) - tic))
    J[:, 12:18] = reduce * je
    J[:, 18] = reduce * T * pi3 * (-1 / lam ** 2)
    return r, J


def line(ps, pe, abc, K, bcR, bcT, pose):
    """Returns (r [2], J 2x6) of LineProjectionFactor with the Jacobian AS CODED."""
    Kp = mp.matrix([[mp.mpf(float(K[3 * i + j])) for j in range(3)] for i in range(3)])
    Rbc = mp.matrix([[mp.mpf(float(bcR[3 * i + j])) for j in range(3)] for i in range(3)])
    R = Rbc.T * _Rn(pose[3:7]).T
    t = -R * _v(pose[:3]) - Rbc.T * _v(bcT)
    a, b, c = [mp.mpf(float(v)) for v in abc]
    d = a * a + b * b
    r, J = [], mp.zeros(2, 6)
    for k, P in enumerate((ps, pe)):
        pc = R * _v(P) + t
        im = Kp * pc
        u, v = im[0] / im[2], im[1] / im[2]
        mu = (b * b * u - a * b * v - a * c) / d
        mv = (a * a * v - a * b * u - b * c) / d
        r.append(mp.sqrt((mu - u) ** 2 + (mv - v) ** 2))
        ep = mp.matrix([[-2 / d * ((mu - u) * a * a + a * b * (mv - v)), -2 / d * ((mu - u) * a * b + b * b * (mv - v))]])
        fx, fy = Kp[0, 0], Kp[1, 1]
        pp = mp.matrix([[fx / pc[2], 0, -fx * pc[0] / pc[2] ** 2], [0, fy / pc[2], -fy * pc[1] / pc[2] ** 2]])
        jac = mp.zeros(3, 6)
        jac[:, 0:3] = mp.eye(3)
        jac[:, 3:6] = _skew(pc)
        J[k, :] = ep * pp * jac
    return r, J


def cauchy_scale(r, a=1.0):
    """sqrt(rho'(s)) of CauchyLoss(a): with rho'' < 0 both r and J are scaled by it (marg.cpp:48-52)."""
    s = sum(x * x for x in r)
    return mp.sqrt(1 / (1 + s / mp.mpf(a) ** 2))
