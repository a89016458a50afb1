"""Second, pure-Python restatement of the line association (small cases only).  TEST INFRASTRUCTURE ONLY.

Python floats are IEEE binary64 with no fused multiply-add, numpy.float32 gives the reference's float
narrowing, math.acos is the host libm — so this must agree BIT FOR BIT with viml_oracle.cpp.  It is written
with Python lists/tuples in a different style on purpose, to catch transcription slips in either version.
Follows estimator.cpp:385-447, :601-669, :671-885 and feature_manager.cpp:4-15, :46-71 of the reference.
"""
import math

import numpy as np

PI = 3.1415926  # feature_manager.h:26
f32 = np.float32


def _rot(q):  # normalized().toRotationMatrix(); q = (x,y,z,w) pose layout
    x, y, z, w = q
    n = math.sqrt(((x * x + y * y) + z * z) + w * w)
    w, x, y, z = w / n, x / n, y / n, z / n
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return [[1.0 - (tyy + tzz), txy - twz, txz + twy],
            [txy + twz, 1.0 - (txx + tzz), tyz - twx],
            [txz - twy, tyz + twx, 1.0 - (txx + tyy)]]


def _mm(A, B):
    return [[(A[r][0] * B[0][c] + A[r][1] * B[1][c]) + A[r][2] * B[2][c] for c in range(3)] for r in range(3)]


def _mv(A, v):
    return [(A[r][0] * v[0] + A[r][1] * v[1]) + A[r][2] * v[2] for r in range(3)]


def _T(A):
    return [[A[c][r] for c in range(3)] for r in range(3)]


def camera(cfg, pose, ex):
    Ric, Rbi = _rot(ex[3:7]), _rot(pose[3:7])
    Rbw = [[cfg.Rbw[3 * r + c] for c in range(3)] for r in range(3)]
    R = _mm(_mm(_T(Ric), _T(Rbi)), Rbw)
    d = [cfg.Tbw[k] - pose[k] for k in range(3)]
    v = _mv(_T(Rbi), d)
    v = [v[k] - ex[k] for k in range(3)]
    return R, _mv(_T(Ric), v)


def _proj(R, T, p):
    v = _mv(R, p)
    return [v[0] + T[0], v[1] + T[1], v[2] + T[2]]


def fov(cfg, pose, ex, lines):
    R, T = camera(cfg, pose, ex)
    wl, wr, hu, hd = -20, 20 + cfg.width, -20, 20 + cfg.height
    out = []
    for j, ln in enumerate(lines):
        s, e = _proj(R, T, ln[:3]), _proj(R, T, ln[3:])
        sf = ef = False
        if s[2] > 0 and e[2] > 0:
            xx = cfg.fx * s[0] / s[2] + cfg.cx
            yy = cfg.fy * s[1] / s[2] + cfg.cy
            xx_ = cfg.fx * e[0] / e[2] + cfg.cx
            yy_ = cfg.fy * e[1] / e[2] + cfg.cy
            sf = xx > wl and xx < (wr - 1) and yy > hu and yy < hd
            ef = xx_ > wl and xx_ < (wr - 1) and yy_ > hu and yy_ < hd
        if sf or ef:
            out.append(j)
    return out


class L2:
    def __init__(self, sx, sy, ex, ey):
        self.S, self.E = (sx, sy), (ex, ey)
        lv = (ex - sx, ey - sy)
        self.Length = math.sqrt(lv[0] * lv[0] + lv[1] * lv[1])
        with np.errstate(all="ignore"):
            self.D = (float(np.float64(lv[0]) / np.float64(self.Length)), float(np.float64(lv[1]) / np.float64(self.Length)))
        self.A = ey - sy
        self.B = sx - ex
        self.C = ex * sy - sx * ey
        self.A2B2 = math.sqrt(self.A * self.A + self.B * self.B)

    def foot(self, p):
        d1 = math.sqrt(_sq(p[0] - self.S[0]) + _sq(p[1] - self.S[1]))
        d2 = math.sqrt(_sq(p[0] - self.E[0]) + _sq(p[1] - self.E[1]))
        A_, B_ = self.B, -self.A
        C_ = -1 * (A_ * p[0] + B_ * p[1])
        with np.errstate(all="ignore"):
            det = np.float64(self.A * B_ - A_ * self.B)
            invdet = float(np.float64(1.0) / det)
        i00, i01, i10, i11 = B_ * invdet, -self.B * invdet, -A_ * invdet, self.A * invdet
        rx, ry = -self.C, -C_
        ix, iy = i00 * rx + i01 * ry, i10 * rx + i11 * ry
        if (ix - self.S[0]) * (ix - self.E[0]) >= 0:
            return self.S if d1 < d2 else self.E
        return (ix, iy)


def _sq(x):
    return x * x


def angle_dist(proj, det):
    c = abs(det.D[0] * proj.D[0] + det.D[1] * proj.D[1])
    try:
        beta = math.acos(c)
    except ValueError:
        beta = float("nan")
    return PI if math.isnan(beta) else beta


def euler_dist(proj, det):
    if det.Length <= proj.Length:
        l1, l2 = det, proj
    else:
        l1, l2 = proj, det
    a, b = l2.foot(l1.S), l2.foot(l1.E)
    with np.errstate(all="ignore"):
        ov = float(np.float64(math.sqrt(_sq(a[0] - b[0]) + _sq(a[1] - b[1]))) / np.float64(l2.Length))
        sx_, sy_ = (l1.S[0] - l1.E[0]) / 10, (l1.S[1] - l1.E[1]) / 10
        dist = 0.0
        for i in range(10):
            x, y = l1.S[0] + i * sx_, l1.S[1] + i * sy_
            dist = dist + float(np.float64(abs(l2.A * x + l2.B * y + l2.C)) / np.float64(l2.A2B2))
        dist = dist + float(np.float64(1 * abs(l2.A * l1.S[0] + l2.B * l1.S[1] + l2.C)) / np.float64(l2.A2B2))
        dist = dist + float(np.float64(1 * abs(l2.A * l1.E[0] + l2.B * l1.E[1] + l2.C)) / np.float64(l2.A2B2))
    dist = dist / 12
    if math.isnan(dist) or math.isnan(ov):
        return 10000.0, 0.0
    return dist, ov


def correspondence(cfg, pose, ex, lines, fov_list, l2d):
    """Returns (map index or -1, (errA, errD, overlap) float32, projected 4-tuple or None)."""
    R, T = camera(cfg, pose, ex)
    det = L2(*[float(v) for v in l2d])
    W, H = cfg.width, cfg.height
    best, choose, err, pl = f32(10000.0), -1, None, None
    for i, j in enumerate(fov_list):
        s, e = _proj(R, T, lines[j][:3]), _proj(R, T, lines[j][3:])
        sf = ef = False
        xx = yy = xx_ = yy_ = f32(0)
        if s[2] > 0 and e[2] > 0:
            xx, yy = f32(cfg.fx * s[0] / s[2] + cfg.cx), f32(cfg.fy * s[1] / s[2] + cfg.cy)
            xx_, yy_ = f32(cfg.fx * e[0] / e[2] + cfg.cx), f32(cfg.fy * e[1] / e[2] + cfg.cy)
            sf = xx > 0 and xx < f32(W - 1) and yy > 0 and yy < f32(H - 1)
            ef = xx_ > 0 and xx_ < f32(W - 1) and yy_ > 0 and yy_ < f32(H - 1)
        tmp = None
        if sf and ef:
            tmp = L2(float(xx), float(yy), float(xx_), float(yy_))
        elif sf != ef:
            base, other = (s, e) if sf else (e, s)
            dv = [other[k] - base[k] for k in range(3)]
            t, found, x, y = 0.9, False, 0.0, 0.0
            while t > 0:
                p = [base[k] + t * dv[k] for k in range(3)]
                if p[2] > 0:
                    x = cfg.fx * p[0] / p[2] + cfg.cx
                    y = cfg.fy * p[1] / p[2] + cfg.cy
                    if x > 0 and x < (W - 1) and y > 0 and y < (H - 1):
                        found = True
                        break
                t = t - 0.1
            if found:
                tmp = L2(float(xx), float(yy), x, y) if sf else L2(x, y, float(xx_), float(yy_))
        if tmp is None:
            continue
        ang = angle_dist(tmp, det)
        if ang > cfg.angle_th:
            continue
        d, o = euler_dist(tmp, det)
        distance, overlap = f32(d), f32(o)
        if float(overlap) < cfg.overlap_th:
            continue
        if distance < best:
            best, choose, pl = distance, i, tmp
            err = (f32(ang), best, overlap)
    if choose < 0:
        return -1, (f32(-1), f32(-1), f32(-1)), None
    return fov_list[choose], err, (pl.S[0], pl.S[1], pl.E[0], pl.E[1])
