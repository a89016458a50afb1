"""Per-unit relative error norms for the 1e-9 parity bar (tests/, smoke(), bench.py; numpy only, no product path).

BASELINE.json asks for residuals, Jacobians, H, b and the Schur complement "within 1e-9 relative".  A single
array-wide scale hides a wrong factor behind the largest one of the batch, so every output is judged against
the scale of its own UNIT:

  pf_* / lf_*   one factor (one row)
  H_pp, S       one 6x6 block of one window
  b_p, g        one 6-vector block of one window
  H_lp          one landmark row of one window
  H_ll, b_l     one window (these are one scalar per landmark; a landmark's own value can be an exact
                cancellation, the window's landmark vector is the unit)

err = max over units of  max|got - ref| / max|ref|  inside the unit.  A unit whose reference is all zeros
(structural zero block) must be reproduced exactly.
"""
import numpy as np

PER_FACTOR = ("pf_residual", "pf_jac_pose_i", "pf_jac_pose_j", "pf_jac_ex", "pf_jac_feat", "lf_residual", "lf_jac_pose")


def _units(name, a):
    """Reshape `a` to [n_units, unit_size] for output `name` (leading dimension = windows or factors)."""
    if a.size == 0:
        return a.reshape(0, 1)
    if name in PER_FACTOR:
        return a.reshape(a.shape[0], -1)
    if name in ("H_pp", "S"):
        W, D, _ = a.shape
        nb = D // 6
        return a.reshape(W, nb, 6, nb, 6).transpose(0, 1, 3, 2, 4).reshape(W * nb * nb, 36)
    if name in ("b_p", "g"):
        return a.reshape(-1, 6)
    if name == "H_lp":
        return a.reshape(-1, a.shape[-1])
    if name in ("H_ll", "b_l"):
        return a.reshape(a.shape[0], -1)
    return a.reshape(1, -1)


def unit_errors(name, got, ref):
    """Per-unit relative errors (1-D array, one entry per unit)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    g, r = _units(name, got), _units(name, ref)
    if r.shape[0] == 0:
        return np.zeros(0)
    scale = np.abs(r).max(axis=1)
    diff = np.abs(g - r).max(axis=1)
    diff = np.where(np.isnan(diff), np.inf, diff)
    out = np.where(scale > 0, diff / np.where(scale > 0, scale, 1.0), np.where(diff == 0, 0.0, np.inf))
    return out


def unit_err(name, got, ref):
    e = unit_errors(name, got, ref)
    return float(e.max()) if e.size else 0.0


def check_outputs(got, ref, tol=1e-9, names=None):
    """Assert every output of `ref` is matched by `got` to `tol` per unit; returns {name: err}."""
    errs = {}
    for k in (names or ref.keys()):
        errs[k] = unit_err(k, got[k], ref[k])
        assert errs[k] < tol, (k, errs[k])
    return errs
