// viml_oracle.cpp — CPU ORACLE (test infrastructure only; see viml_oracle.h for the parity status).
//
// Dependency-free restatement of the TC-VIML hot path.  Every function cites the reference lines it
// follows (paths relative to /root/reference/vins_estimator/src unless noted).  Build with
// -ffp-contract=off: the arithmetic contract has no fused multiply-add.
#include "viml_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------------------------
// Tiny fixed-size algebra with the summation order contract of SURVEY.md A.2.
// ----------------------------------------------------------------------------------------------
struct V3 {
  double x, y, z;
};
struct M3 {
  double m[9];  // row-major
  double operator()(int r, int c) const { return m[3 * r + c]; }
  double& operator()(int r, int c) { return m[3 * r + c]; }
};
struct Quat {
  double w, x, y, z;
};

inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 scale(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
  return (a0 * b0 + a1 * b1) + a2 * b2;
}
inline V3 mul(const M3& A, V3 v) {
  return {dot3(A.m[0], A.m[1], A.m[2], v.x, v.y, v.z), dot3(A.m[3], A.m[4], A.m[5], v.x, v.y, v.z),
          dot3(A.m[6], A.m[7], A.m[8], v.x, v.y, v.z)};
}
inline M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      C(r, c) = dot3(A(r, 0), A(r, 1), A(r, 2), B(0, c), B(1, c), B(2, c));
  return C;
}
inline M3 transpose(const M3& A) {
  M3 T;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T(r, c) = A(c, r);
  return T;
}
inline M3 neg(const M3& A) {
  M3 T;
  for (int k = 0; k < 9; ++k) T.m[k] = -A.m[k];
  return T;
}
inline M3 addm(const M3& A, const M3& B) {
  M3 T;
  for (int k = 0; k < 9; ++k) T.m[k] = A.m[k] + B.m[k];
  return T;
}
inline M3 subm(const M3& A, const M3& B) {
  M3 T;
  for (int k = 0; k < 9; ++k) T.m[k] = A.m[k] - B.m[k];
  return T;
}
inline M3 identity() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }

// Utility::skewSymmetric (utility/utility.h:31-38)
inline M3 skew(V3 q) { return {{0, -q.z, q.y, q.z, 0, -q.x, -q.y, q.x, 0}}; }

// Eigen::Quaterniond(w,x,y,z) built as Quaterniond(p[6],p[3],p[4],p[5]) (projection_factor.cpp:25)
inline Quat quat_from_pose(const double* p) { return {p[6], p[3], p[4], p[5]}; }

// Eigen QuaternionBase::toRotationMatrix — no normalisation.
inline M3 to_rotation(Quat q) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 R;
  R(0, 0) = 1.0 - (tyy + tzz);
  R(0, 1) = txy - twz;
  R(0, 2) = txz + twy;
  R(1, 0) = txy + twz;
  R(1, 1) = 1.0 - (txx + tzz);
  R(1, 2) = tyz - twx;
  R(2, 0) = txz - twy;
  R(2, 1) = tyz + twx;
  R(2, 2) = 1.0 - (txx + tyy);
  return R;
}
// Eigen QuaternionBase::_transformVector:  uv = q.vec x v; uv += uv; v + w*uv + q.vec x uv
inline V3 rotate(Quat q, V3 v) {
  V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = add(uv, uv);
  return add(add(v, scale(q.w, uv)), cross(qv, uv));
}
// Eigen QuaternionBase::inverse: conjugate / squaredNorm
inline Quat inverse(Quat q) {
  const double n2 = ((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w;  // coeffs order x,y,z,w
  if (n2 > 0.0) return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return {0, 0, 0, 0};
}
inline Quat normalized(Quat q) {
  const double n = std::sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
// Eigen quaternion product
inline Quat qmul(Quat a, Quat b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

// 2x3 helper
struct M23 {
  double m[6];
  double operator()(int r, int c) const { return m[3 * r + c]; }
  double& operator()(int r, int c) { return m[3 * r + c]; }
};
inline M23 mul(const M23& A, const M3& B) {
  M23 C;
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 3; ++c)
      C(r, c) = dot3(A(r, 0), A(r, 1), A(r, 2), B(0, c), B(1, c), B(2, c));
  return C;
}
inline void mul(const M23& A, V3 v, double out[2]) {
  out[0] = dot3(A(0, 0), A(0, 1), A(0, 2), v.x, v.y, v.z);
  out[1] = dot3(A(1, 0), A(1, 1), A(1, 2), v.x, v.y, v.z);
}

template <class F>
void parallel_for(int n, int nthreads, F f) {
  nthreads = std::max(1, std::min(nthreads, n));
  if (nthreads == 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([=]() {
      // contiguous ranges, like a static schedule
      const int64_t lo = (int64_t)n * t / nthreads, hi = (int64_t)n * (t + 1) / nthreads;
      for (int i = (int)lo; i < (int)hi; ++i) f(i);
    });
  for (auto& t : th) t.join();
}

}  // namespace

// ================================================================================================
// a2  ProjectionFactor::Evaluate — factor/projection_factor.cpp:21-124 (UNIT_SPHERE_ERROR off,
//     parameters.h:23).  sqrt_info = s*I2 (estimator.cpp:85): (s*I)*v has the extra "+ 0*v1" term of a
//     2x2 product, which is exact, so it is written as s*v.
// ================================================================================================
extern "C" int orc_projection_evaluate(const double* pts_i_, const double* pts_j_, double sqrt_info,
                                       double const* const* parameters, double* residuals,
                                       double** jacobians) {
  const V3 pts_i{pts_i_[0], pts_i_[1], pts_i_[2]}, pts_j{pts_j_[0], pts_j_[1], pts_j_[2]};
  const V3 Pi{parameters[0][0], parameters[0][1], parameters[0][2]};  // :24
  const Quat Qi = quat_from_pose(parameters[0]);                      // :25
  const V3 Pj{parameters[1][0], parameters[1][1], parameters[1][2]};  // :27
  const Quat Qj = quat_from_pose(parameters[1]);                      // :28
  const V3 tic{parameters[2][0], parameters[2][1], parameters[2][2]}; // :30
  const Quat qic = quat_from_pose(parameters[2]);                     // :31
  const double inv_dep_i = parameters[3][0];                          // :33

  const V3 pts_camera_i{pts_i.x / inv_dep_i, pts_i.y / inv_dep_i, pts_i.z / inv_dep_i};  // :35
  const V3 pts_imu_i = add(rotate(qic, pts_camera_i), tic);                                // :36
  const V3 pts_w = add(rotate(Qi, pts_imu_i), Pi);                                         // :37
  const V3 pts_imu_j = rotate(inverse(Qj), sub(pts_w, Pj));                                // :38
  const V3 pts_camera_j = rotate(inverse(qic), sub(pts_imu_j, tic));                       // :39

  const double dep_j = pts_camera_j.z;                                                     // :45
  double r0 = pts_camera_j.x / dep_j - pts_j.x;                                            // :46
  double r1 = pts_camera_j.y / dep_j - pts_j.y;
  residuals[0] = sqrt_info * r0;                                                           // :49
  residuals[1] = sqrt_info * r1;

  if (jacobians) {                                                                         // :54
    const M3 Ri = to_rotation(Qi), Rj = to_rotation(Qj), ric = to_rotation(qic);           // :56-58
    M23 reduce;                                                                            // :72-75
    reduce(0, 0) = 1. / dep_j;
    reduce(0, 1) = 0;
    reduce(0, 2) = -pts_camera_j.x / (dep_j * dep_j);
    reduce(1, 0) = 0;
    reduce(1, 1) = 1. / dep_j;
    reduce(1, 2) = -pts_camera_j.y / (dep_j * dep_j);
    for (int k = 0; k < 6; ++k) reduce.m[k] = sqrt_info * reduce.m[k];                     // :75

    const M3 ricT = transpose(ric), RjT = transpose(Rj);
    if (jacobians[0]) {                                                                    // :77-87
      const M3 left = mul(ricT, RjT);
      const M3 right = mul(mul(mul(ricT, RjT), Ri), neg(skew(pts_imu_i)));
      const M23 jl = mul(reduce, left), jr = mul(reduce, right);
      double* J = jacobians[0];
      for (int r = 0; r < 2; ++r) {
        for (int c = 0; c < 3; ++c) {
          J[7 * r + c] = jl(r, c);
          J[7 * r + 3 + c] = jr(r, c);
        }
        J[7 * r + 6] = 0.0;
      }
    }
    if (jacobians[1]) {                                                                    // :89-99
      const M3 left = mul(ricT, neg(RjT));
      const M3 right = mul(ricT, skew(pts_imu_j));
      const M23 jl = mul(reduce, left), jr = mul(reduce, right);
      double* J = jacobians[1];
      for (int r = 0; r < 2; ++r) {
        for (int c = 0; c < 3; ++c) {
          J[7 * r + c] = jl(r, c);
          J[7 * r + 3 + c] = jr(r, c);
        }
        J[7 * r + 6] = 0.0;
      }
    }
    if (jacobians[2]) {                                                                    // :100-110
      const M3 left = mul(ricT, subm(mul(RjT, Ri), identity()));
      const M3 tmp_r = mul(mul(mul(ricT, RjT), Ri), ric);
      const V3 v3 = mul(ricT, sub(mul(RjT, sub(add(mul(Ri, tic), Pi), Pj)), tic));
      const M3 right = addm(addm(mul(neg(tmp_r), skew(pts_camera_i)), skew(mul(tmp_r, pts_camera_i))),
                            skew(v3));
      const M23 jl = mul(reduce, left), jr = mul(reduce, right);
      double* J = jacobians[2];
      for (int r = 0; r < 2; ++r) {
        for (int c = 0; c < 3; ++c) {
          J[7 * r + c] = jl(r, c);
          J[7 * r + 3 + c] = jr(r, c);
        }
        J[7 * r + 6] = 0.0;
      }
    }
    if (jacobians[3]) {                                                                    // :111-119
      // reduce * ric^T * Rj^T * Ri * ric * pts_i * -1.0 / (inv_dep_i * inv_dep_i), left to right
      const M23 chain = mul(mul(mul(mul(reduce, ricT), RjT), Ri), ric);
      double v[2];
      mul(chain, pts_i, v);
      jacobians[3][0] = v[0] * -1.0 / (inv_dep_i * inv_dep_i);
      jacobians[3][1] = v[1] * -1.0 / (inv_dep_i * inv_dep_i);
    }
  }
  return 1;  // Evaluate returns true always (:123)
}

// ================================================================================================
// a4  LineProjectionFactor::Evaluate — factor/line_projection_factor.cpp:19-120.  The Jacobian is
//     replicated AS CODED (it is not d r / d pose).
// ================================================================================================
extern "C" int orc_line_evaluate(const double* ps, const double* pe, const double* lp,
                                 const double* K_, const double* bcR_, const double* bcT_,
                                 double const* const* parameters, double* residuals,
                                 double** jacobians) {
  M3 K, b_c_R;
  std::memcpy(K.m, K_, sizeof(K.m));
  std::memcpy(b_c_R.m, bcR_, sizeof(b_c_R.m));
  const V3 b_c_T{bcT_[0], bcT_[1], bcT_[2]};
  const V3 pts_start{ps[0], ps[1], ps[2]}, pts_end{pe[0], pe[1], pe[2]};

  const V3 T_w{parameters[0][0], parameters[0][1], parameters[0][2]};                       // :29
  const M3 R_w = to_rotation(normalized(quat_from_pose(parameters[0])));                    // :31-33
  const M3 R = mul(transpose(b_c_R), transpose(R_w));                                       // :39
  const V3 t = sub(mul(neg(R), T_w), mul(transpose(b_c_R), b_c_T));                         // :40
  const V3 pcs = add(mul(R, pts_start), t);                                                 // :42
  const V3 pce = add(mul(R, pts_end), t);                                                   // :43
  const V3 psi = mul(K, pcs), pei = mul(K, pce);                                            // :45-46
  const double u_start = psi.x / psi.z, v_start = psi.y / psi.z;                            // :48-49
  const double u_end = pei.x / pei.z, v_end = pei.y / pei.z;                                // :50-51
  const double a = lp[0], b = lp[1], c = lp[2];
  const double d = a * a + b * b;                                                           // :56
  const double mea_u_start = (b * b * u_start - a * b * v_start - a * c) / (d);             // :58
  const double mea_v_start = (a * a * v_start - a * b * u_start - b * c) / (d);
  const double mea_u_end = (b * b * u_end - a * b * v_end - a * c) / (d);
  const double mea_v_end = (a * a * v_end - a * b * u_end - b * c) / (d);
  const double rho_line = 1.0;
  residuals[0] = rho_line * std::sqrt((mea_u_start - u_start) * (mea_u_start - u_start) +
                                      (mea_v_start - v_start) * (mea_v_start - v_start));   // :68
  residuals[1] = rho_line * std::sqrt((mea_u_end - u_end) * (mea_u_end - u_end) +
                                      (mea_v_end - v_end) * (mea_v_end - v_end));           // :69
  const double lambda = 1;
  if (jacobians) {                                                                          // :73
    const double ep11 = -2 / d * ((mea_u_start - u_start) * a * a + a * b * (mea_v_start - v_start)) * lambda;
    const double ep12 = -2 / d * ((mea_u_start - u_start) * a * b + b * b * (mea_v_start - v_start)) * lambda;
    const double ep11_ = -2 / d * ((mea_u_end - u_end) * a * a + a * b * (mea_v_end - v_end)) * lambda;
    const double ep12_ = -2 / d * ((mea_u_end - u_end) * a * b + b * b * (mea_v_end - v_end)) * lambda;
    const double fx = K(0, 0), fy = K(1, 1);
    M23 pps, ppe;                                                                           // :93-100
    pps(0, 0) = fx / pcs.z; pps(0, 1) = 0; pps(0, 2) = -fx * pcs.x / (pcs.z * pcs.z);
    pps(1, 0) = 0; pps(1, 1) = fy / pcs.z; pps(1, 2) = -fy * pcs.y / (pcs.z * pcs.z);
    ppe(0, 0) = fx / pce.z; ppe(0, 1) = 0; ppe(0, 2) = -fx * pce.x / (pce.z * pce.z);
    ppe(1, 0) = 0; ppe(1, 1) = fy / pce.z; ppe(1, 2) = -fy * pce.y / (pce.z * pce.z);
    // (_e_p * _p_p_s) is 1x3; times jaco = [I3 | skew(pc)] is 1x6      (:104-113)
    auto row = [](double e1, double e2, const M23& pp, V3 pc, double* out) {
      double w[3];
      for (int cidx = 0; cidx < 3; ++cidx) w[cidx] = e1 * pp(0, cidx) + e2 * pp(1, cidx);
      const M3 S = skew(pc);
      for (int cidx = 0; cidx < 3; ++cidx) {
        // [I | S]: column c of I picks w[c] with two exact zero terms
        out[cidx] = w[cidx];
        out[3 + cidx] = dot3(w[0], w[1], w[2], S(0, cidx), S(1, cidx), S(2, cidx));
      }
      out[6] = 0.0;                                                                         // :114
    };
    row(ep11, ep12, pps, pcs, jacobians[0]);        // written without a NULL check (:102)
    row(ep11_, ep12_, ppe, pce, jacobians[0] + 7);
  }
  return 1;
}

// ================================================================================================
// A.4  ceres::CauchyLoss::Evaluate and the corrector as re-implemented by the reference.
// ================================================================================================
extern "C" void orc_cauchy_loss(double a, double s, double rho[3]) {
  const double b = a * a, c = 1 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho[0] = b * std::log(sum);
  rho[1] = std::max(std::numeric_limits<double>::min(), inv);
  rho[2] = -c * (inv * inv);
}

// a5  ResidualBlockInfo::Evaluate loss part — factor/marginalization_factor.cpp:37-68
extern "C" void orc_loss_correct(double cauchy_a, int nres, double* residuals, int nblk,
                                 const int* sizes, double** jacobians) {
  double sq_norm = 0.0, rho[3];
  for (int r = 0; r < nres; ++r) sq_norm += residuals[r] * residuals[r];                    // :42
  orc_cauchy_loss(cauchy_a, sq_norm, rho);                                                  // :43
  const double sqrt_rho1_ = std::sqrt(rho[1]);                                              // :46
  double residual_scaling_, alpha_sq_norm_;
  if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {                                                // :48
    residual_scaling_ = sqrt_rho1_;
    alpha_sq_norm_ = 0.0;
  } else {
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling_ = sqrt_rho1_ / (1 - alpha);
    alpha_sq_norm_ = alpha / sq_norm;
  }
  for (int k = 0; k < nblk; ++k) {                                                          // :62-65
    double* J = jacobians[k];
    if (!J) continue;
    const int nc = sizes[k];
    std::vector<double> rtJ(nc, 0.0);
    for (int c = 0; c < nc; ++c) {
      double s = 0.0;
      for (int r = 0; r < nres; ++r) s += residuals[r] * J[r * nc + c];
      rtJ[c] = s;
    }
    for (int r = 0; r < nres; ++r)
      for (int c = 0; c < nc; ++c)
        J[r * nc + c] = sqrt_rho1_ * (J[r * nc + c] - alpha_sq_norm_ * residuals[r] * rtJ[c]);
  }
  for (int r = 0; r < nres; ++r) residuals[r] *= residual_scaling_;                         // :67
}

// ================================================================================================
// Batch restatement (a2, a4, a5, a7 + header-defined landmark Schur).
// ================================================================================================
namespace {

struct FactorEval {          // one evaluated residual block, like ResidualBlockInfo
  int nblk;
  int idx[4];                // canonical dense index of each block
  int lsize[4];              // local size (7 -> 6, marginalization_factor.h:31-34)
  double r[2];
  double J[4][14];           // row-major 2 x gsize
  int gsize[4];
};

void eval_point_factor(const viml_config* cfg, const viml_window_batch* in, int w, int64_t k,
                       uint32_t flags, FactorEval* fe, double* Jout[4], double* rout) {
  const int P = in->poses_per_window, F = in->feats_per_window;
  const uint32_t pk = in->pf_idx[k];
  const int i = pk & 0xff, j = (pk >> 8) & 0xff, feat = pk >> 16;
  const double* poses = in->poses + (size_t)w * P * 7;
  const double* params[4] = {poses + 7 * i, poses + 7 * j, in->ex_pose + (size_t)w * 7,
                             in->inv_depth + (size_t)w * F + feat};
  const double pts_i[3] = {in->pf_obs[4 * k + 0], in->pf_obs[4 * k + 1],
                           in->pf_pts_i_z ? in->pf_pts_i_z[k] : 1.0};
  const double pts_j[3] = {in->pf_obs[4 * k + 2], in->pf_obs[4 * k + 3], 1.0};
  double* Jp[4] = {fe->J[0], fe->J[1], fe->J[2], fe->J[3]};
  orc_projection_evaluate(pts_i, pts_j, cfg->sqrt_info, params, fe->r, Jp);
  fe->nblk = 4;
  const int gs[4] = {7, 7, 7, 1};
  for (int b = 0; b < 4; ++b) fe->gsize[b] = gs[b], fe->lsize[b] = gs[b] == 7 ? 6 : gs[b];
  fe->idx[0] = 6 * i;
  fe->idx[1] = 6 * j;
  fe->idx[2] = 6 * P;
  fe->idx[3] = 6 * (P + 1) + feat;
  if (flags & VIML_LOSS_CAUCHY) orc_loss_correct(cfg->cauchy_a, 2, fe->r, 4, gs, Jp);
  if (rout) rout[0] = fe->r[0], rout[1] = fe->r[1];
  for (int b = 0; b < 4; ++b)
    if (Jout && Jout[b]) std::memcpy(Jout[b], fe->J[b], sizeof(double) * 2 * gs[b]);
}

void eval_line_factor(const viml_config* cfg, const viml_window_batch* in, int w, int64_t k,
                      uint32_t flags, FactorEval* fe, double* Jout, double* rout) {
  const int P = in->poses_per_window;
  const int64_t NL = in->n_line_factors;
  const int frame = in->lf_frame[k];
  const double* pose = in->poses + ((size_t)w * P + frame) * 7;
  const double* ex = in->ex_pose + (size_t)w * 7;
  double g[9];
  for (int c = 0; c < 9; ++c) g[c] = in->lf_geom[(size_t)c * NL + k];
  // _Ric = Quaterniond(ex).normalized().toRotationMatrix(), _Tic = ex[0..2]  (estimator.cpp:1777-1781)
  const M3 Ric = to_rotation(normalized(quat_from_pose(ex)));
  const double Kmat[9] = {cfg->fx, 0, cfg->cx, 0, cfg->fy, cfg->cy, 0, 0, 1};
  const double* params[1] = {pose};
  double* Jp[1] = {fe->J[0]};
  orc_line_evaluate(g, g + 3, g + 6, Kmat, Ric.m, ex, params, fe->r, Jp);
  fe->nblk = 1;
  fe->gsize[0] = 7;
  fe->lsize[0] = 6;
  fe->idx[0] = 6 * frame;
  const int gs[1] = {7};
  if (flags & VIML_LOSS_CAUCHY) orc_loss_correct(cfg->cauchy_a, 2, fe->r, 1, gs, Jp);
  if (rout) rout[0] = fe->r[0], rout[1] = fe->r[1];
  if (Jout) std::memcpy(Jout, fe->J[0], sizeof(double) * 14);
}

// ThreadsConstructA — factor/marginalization_factor.cpp:141-172, on one partial (A, b).
void construct_A(const FactorEval& f, double* A, double* b, int pos) {
  for (int i = 0; i < f.nblk; ++i) {
    const int idx_i = f.idx[i], size_i = f.lsize[i], gi = f.gsize[i];
    for (int j = i; j < f.nblk; ++j) {
      const int idx_j = f.idx[j], size_j = f.lsize[j], gj = f.gsize[j];
      for (int r = 0; r < size_i; ++r)
        for (int c = 0; c < size_j; ++c) {
          const double v = f.J[i][0 * gi + r] * f.J[j][0 * gj + c] + f.J[i][1 * gi + r] * f.J[j][1 * gj + c];
          A[(size_t)(idx_i + r) * pos + idx_j + c] += v;                                    // :160/:163
        }
      if (i != j)                                                                           // :165 (assign)
        for (int r = 0; r < size_i; ++r)
          for (int c = 0; c < size_j; ++c)
            A[(size_t)(idx_j + c) * pos + idx_i + r] = A[(size_t)(idx_i + r) * pos + idx_j + c];
    }
    for (int r = 0; r < size_i; ++r)
      b[idx_i + r] += f.J[i][0 * gi + r] * f.r[0] + f.J[i][1 * gi + r] * f.r[1];            // :168
  }
}

constexpr int kNumThreads = 4;  // NUM_THREADS, factor/marginalization_factor.h:13

// Dense A, b of window w: round-robin deal to 4 partials (:232-241), summed 3,2,1,0 (:256-261).
void window_dense(const viml_config* cfg, const viml_window_batch* in, int w, uint32_t flags,
                  double* A, double* b, const viml_linearize_out* out) {
  const int P = in->poses_per_window, F = in->feats_per_window;
  const int pos = 6 * (P + 1) + F;
  const bool want_hb = A != nullptr;
  std::vector<double> pA, pb;
  if (want_hb) {
    pA.assign((size_t)kNumThreads * pos * pos, 0.0);
    pb.assign((size_t)kNumThreads * pos, 0.0);
  }
  int dealt = 0;
  FactorEval fe;
  for (int64_t k = in->pf_window_offset[w]; k < in->pf_window_offset[w + 1]; ++k) {
    double* Jout[4] = {nullptr, nullptr, nullptr, nullptr};
    double* rout = nullptr;
    if (out && (flags & VIML_OUT_RESIDUAL_JACOBIAN)) {
      if (out->pf_residual) rout = out->pf_residual + 2 * k;
      if (out->pf_jac_pose_i) Jout[0] = out->pf_jac_pose_i + 14 * k;
      if (out->pf_jac_pose_j) Jout[1] = out->pf_jac_pose_j + 14 * k;
      if (out->pf_jac_ex) Jout[2] = out->pf_jac_ex + 14 * k;
      if (out->pf_jac_feat) Jout[3] = out->pf_jac_feat + 2 * k;
    }
    eval_point_factor(cfg, in, w, k, flags, &fe, Jout, rout);
    if (want_hb) {
      const int t = dealt++ % kNumThreads;
      construct_A(fe, pA.data() + (size_t)t * pos * pos, pb.data() + (size_t)t * pos, pos);
    }
  }
  if (in->n_line_factors > 0 && in->lf_window_offset)
    for (int64_t k = in->lf_window_offset[w]; k < in->lf_window_offset[w + 1]; ++k) {
      double* Jout = nullptr;
      double* rout = nullptr;
      if (out && (flags & VIML_OUT_RESIDUAL_JACOBIAN)) {
        if (out->lf_residual) rout = out->lf_residual + 2 * k;
        if (out->lf_jac_pose) Jout = out->lf_jac_pose + 14 * k;
      }
      eval_line_factor(cfg, in, w, k, flags, &fe, Jout, rout);
      if (want_hb) {
        const int t = dealt++ % kNumThreads;
        construct_A(fe, pA.data() + (size_t)t * pos * pos, pb.data() + (size_t)t * pos, pos);
      }
    }
  if (want_hb) {
    std::fill(A, A + (size_t)pos * pos, 0.0);
    std::fill(b, b + pos, 0.0);
    for (int t = kNumThreads - 1; t >= 0; --t) {
      const double* a = pA.data() + (size_t)t * pos * pos;
      const double* bb = pb.data() + (size_t)t * pos;
      for (size_t e = 0; e < (size_t)pos * pos; ++e) A[e] += a[e];
      for (int e = 0; e < pos; ++e) b[e] += bb[e];
    }
  }
}

}  // namespace

extern "C" int orc_window_dense(const viml_config* cfg, const viml_window_batch* in, int w,
                                uint32_t flags, double* A, double* b) {
  if (!cfg || !in || w < 0 || w >= in->n_windows || !A || !b) return VIML_ERR_INVALID;
  window_dense(cfg, in, w, flags, A, b, nullptr);
  return VIML_OK;
}

extern "C" int orc_linearize_batch(const viml_config* cfg, const viml_window_batch* in,
                                   const viml_linearize_out* out, uint32_t flags, int nthreads) {
  if (!cfg || !in || !out) return VIML_ERR_INVALID;
  const int W = in->n_windows, P = in->poses_per_window, F = in->feats_per_window;
  const int D = 6 * (P + 1), pos = D + F;
  const bool want_hb = (flags & (VIML_OUT_HB | VIML_OUT_SCHUR)) != 0;
  parallel_for(W, nthreads, [&](int w) {
    std::vector<double> A, b;
    if (want_hb) A.resize((size_t)pos * pos), b.resize(pos);
    window_dense(cfg, in, w, flags, want_hb ? A.data() : nullptr, want_hb ? b.data() : nullptr, out);
    if (!want_hb) return;
    if (flags & VIML_OUT_HB) {
      if (out->H_pp)
        for (int r = 0; r < D; ++r)
          std::memcpy(out->H_pp + ((size_t)w * D + r) * D, A.data() + (size_t)r * pos, sizeof(double) * D);
      if (out->H_lp)
        for (int l = 0; l < F; ++l)
          std::memcpy(out->H_lp + ((size_t)w * F + l) * D, A.data() + (size_t)(D + l) * pos, sizeof(double) * D);
      if (out->H_ll)
        for (int l = 0; l < F; ++l) out->H_ll[(size_t)w * F + l] = A[(size_t)(D + l) * pos + D + l];
      if (out->b_p) std::memcpy(out->b_p + (size_t)w * D, b.data(), sizeof(double) * D);
      if (out->b_l) std::memcpy(out->b_l + (size_t)w * F, b.data() + D, sizeof(double) * F);
    }
    if (flags & VIML_OUT_SCHUR) {
      // S = H_pp - sum_l W_l^T W_l / L_l ; g = b_p - sum_l W_l^T b_l / L_l ; L_l <= eps dropped.
      std::vector<double> S((size_t)D * D), g(D);
      for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) S[(size_t)r * D + c] = A[(size_t)r * pos + c];
      for (int r = 0; r < D; ++r) g[r] = b[r];
      for (int l = 0; l < F; ++l) {
        const double L = A[(size_t)(D + l) * pos + D + l];
        if (!(L > 1e-8)) continue;
        const double inv = 1.0 / L;
        const double* Wl = A.data() + (size_t)(D + l) * pos;
        const double bl = b[D + l];
        for (int r = 0; r < D; ++r) {
          const double wr = Wl[r];
          if (wr == 0.0) continue;
          const double wri = wr * inv;
          for (int c = 0; c < D; ++c) S[(size_t)r * D + c] -= wri * Wl[c];
          g[r] -= wri * bl;
        }
      }
      if (out->S) std::memcpy(out->S + (size_t)w * D * D, S.data(), sizeof(double) * D * D);
      if (out->g) std::memcpy(out->g + (size_t)w * D, g.data(), sizeof(double) * D);
    }
  });
  return VIML_OK;
}

// ================================================================================================
// A.3  Symmetric eigen-decomposition (cyclic Jacobi) standing in for Eigen::SelfAdjointEigenSolver.
// ================================================================================================
extern "C" int orc_sym_eig(const double* a_in, int n, double* w, double* v) {
  std::vector<double> a((size_t)n * n);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) a[(size_t)r * n + c] = r >= c ? a_in[(size_t)r * n + c] : a_in[(size_t)c * n + r];
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) v[(size_t)r * n + c] = r == c ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int r = 0; r < n; ++r) {
      diag += a[(size_t)r * n + r] * a[(size_t)r * n + r];
      for (int c = r + 1; c < n; ++c) off += a[(size_t)r * n + c] * a[(size_t)r * n + c];
    }
    if (off <= 1e-34 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = a[(size_t)p * n + p], aqq = a[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns p,q of a
          const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - s * akq;
          a[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {  // rows p,q of a
          const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - s * aqk;
          a[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
          v[(size_t)k * n + p] = c * vkp - s * vkq;
          v[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  // ascending order like SelfAdjointEigenSolver
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(),
            [&](int x, int y) { return a[(size_t)x * n + x] < a[(size_t)y * n + y]; });
  std::vector<double> vs((size_t)n * n);
  for (int i = 0; i < n; ++i) {
    w[i] = a[(size_t)order[i] * n + order[i]];
    for (int k = 0; k < n; ++k) vs[(size_t)k * n + i] = v[(size_t)k * n + order[i]];
  }
  std::memcpy(v, vs.data(), sizeof(double) * n * n);
  return VIML_OK;
}

// ================================================================================================
// a8  MarginalizationInfo::marginalize numeric core — factor/marginalization_factor.cpp:264-293
// ================================================================================================
extern "C" int orc_marginalize_dense(const double* A, const double* b, int pos, int m, double eps,
                                     double* A_schur, double* b_schur, double* lin_jac,
                                     double* lin_res) {
  if (!A || !b || m < 0 || m > pos) return VIML_ERR_INVALID;
  const int n = pos - m;
  std::vector<double> Amm((size_t)m * m), Amm_inv((size_t)m * m, 0.0);
  for (int r = 0; r < m; ++r)                                                               // :267
    for (int c = 0; c < m; ++c)
      Amm[(size_t)r * m + c] = 0.5 * (A[(size_t)r * pos + c] + A[(size_t)c * pos + r]);
  if (m > 0) {
    std::vector<double> w(m), V((size_t)m * m);
    orc_sym_eig(Amm.data(), m, w.data(), V.data());                                         // :268
    // V * diag(lambda > eps ? 1/lambda : 0) * V^T                                         // :272
    std::vector<double> VD((size_t)m * m);
    for (int r = 0; r < m; ++r)
      for (int k = 0; k < m; ++k) VD[(size_t)r * m + k] = V[(size_t)r * m + k] * (w[k] > eps ? 1.0 / w[k] : 0.0);
    for (int r = 0; r < m; ++r)
      for (int c = 0; c < m; ++c) {
        double s = 0.0;
        for (int k = 0; k < m; ++k) s += VD[(size_t)r * m + k] * V[(size_t)c * m + k];
        Amm_inv[(size_t)r * m + c] = s;
      }
  }
  // A = Arr - Arm * Amm_inv * Amr ; b = brr - Arm * Amm_inv * bmm                          // :275-282
  std::vector<double> T((size_t)n * m, 0.0);  // Arm * Amm_inv
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < m; ++c) {
      double s = 0.0;
      for (int k = 0; k < m; ++k) s += A[(size_t)(m + r) * pos + k] * Amm_inv[(size_t)k * m + c];
      T[(size_t)r * m + c] = s;
    }
  std::vector<double> Ar((size_t)n * n), br(n);
  for (int r = 0; r < n; ++r) {
    for (int c = 0; c < n; ++c) {
      double s = 0.0;
      for (int k = 0; k < m; ++k) s += T[(size_t)r * m + k] * A[(size_t)k * pos + m + c];
      Ar[(size_t)r * n + c] = A[(size_t)(m + r) * pos + m + c] - s;
    }
    double s = 0.0;
    for (int k = 0; k < m; ++k) s += T[(size_t)r * m + k] * b[k];
    br[r] = b[m + r] - s;
  }
  if (A_schur) std::memcpy(A_schur, Ar.data(), sizeof(double) * n * n);
  if (b_schur) std::memcpy(b_schur, br.data(), sizeof(double) * n);
  if (lin_jac || lin_res) {
    std::vector<double> w(n), V((size_t)n * n);
    orc_sym_eig(Ar.data(), n, w.data(), V.data());                                          // :284
    for (int k = 0; k < n; ++k) {
      const double S = w[k] > eps ? w[k] : 0.0;                                             // :285
      const double S_inv = w[k] > eps ? 1.0 / w[k] : 0.0;                                   // :286
      const double S_sqrt = std::sqrt(S), S_inv_sqrt = std::sqrt(S_inv);                    // :288-289
      if (lin_jac)                                                                          // :292
        for (int c = 0; c < n; ++c) lin_jac[(size_t)k * n + c] = S_sqrt * V[(size_t)c * n + k];
      if (lin_res) {                                                                        // :293
        double s = 0.0;
        for (int c = 0; c < n; ++c) s += (S_inv_sqrt * V[(size_t)c * n + k]) * br[c];
        lin_res[k] = s;
      }
    }
  }
  return VIML_OK;
}

// ================================================================================================
// a9  MarginalizationFactor::Evaluate — factor/marginalization_factor.cpp:335-384
// ================================================================================================
extern "C" int orc_marginalization_factor_evaluate(int n, int m, int nblk, const int* keep_block_size,
                                                   const int* keep_block_idx,
                                                   double const* const* keep_block_data,
                                                   const double* lin_jac, const double* lin_res,
                                                   double const* const* parameters, double* residuals,
                                                   double** jacobians) {
  std::vector<double> dx(n, 0.0);
  for (int i = 0; i < nblk; ++i) {
    const int size = keep_block_size[i], idx = keep_block_idx[i] - m;
    const double* x = parameters[i];
    const double* x0 = keep_block_data[i];
    if (size != 7) {
      for (int k = 0; k < size; ++k) dx[idx + k] = x[k] - x0[k];                            // :356
    } else {
      for (int k = 0; k < 3; ++k) dx[idx + k] = x[k] - x0[k];                               // :359
      const Quat dq = qmul(inverse(quat_from_pose(x0)), quat_from_pose(x));                 // :360 (positify is identity, utility.h:41-48)
      dx[idx + 3] = 2.0 * dq.x;
      dx[idx + 4] = 2.0 * dq.y;
      dx[idx + 5] = 2.0 * dq.z;
      if (!(dq.w >= 0)) {                                                                   // :361-364
        dx[idx + 3] = 2.0 * -dq.x;
        dx[idx + 4] = 2.0 * -dq.y;
        dx[idx + 5] = 2.0 * -dq.z;
      }
    }
  }
  for (int r = 0; r < n; ++r) {                                                             // :366
    double s = 0.0;
    for (int c = 0; c < n; ++c) s += lin_jac[(size_t)r * n + c] * dx[c];
    residuals[r] = lin_res[r] + s;
  }
  if (jacobians)
    for (int i = 0; i < nblk; ++i)
      if (jacobians[i]) {                                                                   // :371-380
        const int size = keep_block_size[i], local = size == 7 ? 6 : size;
        const int idx = keep_block_idx[i] - m;
        for (int r = 0; r < n; ++r)
          for (int c = 0; c < size; ++c)
            jacobians[i][(size_t)r * size + c] = c < local ? lin_jac[(size_t)r * n + idx + c] : 0.0;
      }
  return 1;
}

// ================================================================================================
// Line association (a11-a17).
// ================================================================================================
namespace {

#define ORC_PI 3.1415926  // feature_manager.h:26

struct Line2D {  // feature_manager.h:30-50, single-argument ctor feature_manager.cpp:4-15
  double Sx, Sy, Ex, Ey;
  double Length, Dx, Dy;
  double A, B, C, A2B2;
};
Line2D make_line2d(double sx, double sy, double ex, double ey) {
  Line2D L;
  L.Sx = sx, L.Sy = sy, L.Ex = ex, L.Ey = ey;
  const double lvx = ex - sx, lvy = ey - sy;           // LineVec = PtrEnd - PtrStart
  L.Length = std::sqrt(lvx * lvx + lvy * lvy);         // .norm()
  L.Dx = lvx / L.Length, L.Dy = lvy / L.Length;        // Direction = LineVec / Length
  L.A = ey - sy;                                       // :11
  L.B = sx - ex;                                       // :12
  L.C = ex * sy - sx * ey;                             // :13
  L.A2B2 = std::sqrt(L.A * L.A + L.B * L.B);           // :14
  return L;
}

// Line2D::Point2Flined — feature_manager.cpp:46-71
void point2flined(const Line2D& L, double px, double py, double* ox, double* oy) {
  const double t1x = px - L.Sx, t1y = py - L.Sy;
  const double d1 = std::sqrt(t1x * t1x + t1y * t1y);
  const double t2x = px - L.Ex, t2y = py - L.Ey;
  const double d2 = std::sqrt(t2x * t2x + t2y * t2y);
  const double A_ = L.B, B_ = -L.A;
  const double C_ = -1 * (A_ * px + B_ * py);
  // Cof << A, B, A_, B_;  Cof.inverse() (Eigen 2x2: adjugate * (1/det))
  const double det = L.A * B_ - A_ * L.B;
  const double invdet = 1.0 / det;
  const double i00 = B_ * invdet, i01 = -L.B * invdet, i10 = -A_ * invdet, i11 = L.A * invdet;
  const double rx = -L.C, ry = -C_;
  const double ix = i00 * rx + i01 * ry, iy = i10 * rx + i11 * ry;
  if ((ix - L.Sx) * (ix - L.Ex) >= 0) {
    if (d1 < d2)
      *ox = L.Sx, *oy = L.Sy;
    else
      *ox = L.Ex, *oy = L.Ey;
  } else {
    *ox = ix, *oy = iy;
  }
}

// Estimator::CalAngleDist — estimator.cpp:601-613
double cal_angle_dist(const Line2D& projectedL, const Line2D& detectedL) {
  double beta = std::acos(std::fabs(detectedL.Dx * projectedL.Dx + detectedL.Dy * projectedL.Dy));
  if (std::isnan(beta)) beta = ORC_PI;
  return beta;
}

// Estimator::CalEulerDist — estimator.cpp:615-669
void cal_euler_dist(const Line2D& projectedL, const Line2D& detectedL, double* dist_out,
                    double* ovl_out) {
  const int sampleNum = 10;
  const double lengthM = detectedL.Length, lengthP = projectedL.Length;
  const Line2D& line1 = (lengthM <= lengthP) ? detectedL : projectedL;                      // :632-643
  const Line2D& line2 = (lengthM <= lengthP) ? projectedL : detectedL;
  double ax, ay, bx, by;
  point2flined(line2, line1.Sx, line1.Sy, &ax, &ay);                                        // :644
  point2flined(line2, line1.Ex, line1.Ey, &bx, &by);
  const double dx = ax - bx, dy = ay - by;
  const double overlap_ratio = std::sqrt(dx * dx + dy * dy) / line2.Length;                 // :645
  const double point_x = line1.Sx, point_y = line1.Sy;
  const double len_x = line1.Sx - line1.Ex, len_y = line1.Sy - line1.Ey;                    // :649-650
  const double step_x = len_x / sampleNum, step_y = len_y / sampleNum;
  double distance = 0.0;
  for (int i = 0; i < sampleNum; ++i) {                                                     // :655-660
    const double x = point_x + i * step_x, y = point_y + i * step_y;
    distance = distance + std::fabs(line2.A * x + line2.B * y + line2.C) / line2.A2B2;
  }
  distance = distance + 1 * std::fabs(line2.A * line1.Sx + line2.B * line1.Sy + line2.C) / line2.A2B2;
  distance = distance + 1 * std::fabs(line2.A * line1.Ex + line2.B * line1.Ey + line2.C) / line2.A2B2;
  distance = distance / (sampleNum + 2);                                                    // :663
  if (std::isnan(distance) || std::isnan(overlap_ratio)) {                                  // :665
    *dist_out = 10000.0;
    *ovl_out = 0.0;
  } else {
    *dist_out = distance;
    *ovl_out = overlap_ratio;
  }
}

struct CamPose {
  M3 R;
  V3 T;
};
// R = Ric^T * Rbi^T * Rbw ; T = Ric^T * (Rbi^T * (Tbw - Tbi) - Tic)   estimator.cpp:391-403 / :679-692
CamPose camera_pose(const viml_config* cfg, const double* pose, const double* ex) {
  const V3 Tic{ex[0], ex[1], ex[2]};
  const M3 Ric = to_rotation(normalized(quat_from_pose(ex)));   // initialLineFoVWindow :487-491 / :680-683
  const V3 Tbi{pose[0], pose[1], pose[2]};
  const M3 Rbi = to_rotation(normalized(quat_from_pose(pose)));
  M3 Rbw;
  std::memcpy(Rbw.m, cfg->Rbw, sizeof(Rbw.m));
  const V3 Tbw{cfg->Tbw[0], cfg->Tbw[1], cfg->Tbw[2]};
  CamPose cp;
  cp.R = mul(mul(transpose(Ric), transpose(Rbi)), Rbw);
  cp.T = mul(transpose(Ric), sub(mul(transpose(Rbi), sub(Tbw, Tbi)), Tic));
  return cp;
}

int fov_cull(const viml_config* cfg, const CamPose& cp, const double* map, int64_t n, int32_t* out,
             uint32_t* mask) {
  const int WINDOW_SIZE = 10;
  const int height_up = -2 * WINDOW_SIZE, height_down = 2 * WINDOW_SIZE + cfg->height;      // :405-408
  const int width_left = -2 * WINDOW_SIZE, width_right = 2 * WINDOW_SIZE + cfg->width;
  int cnt = 0;
  for (int64_t j = 0; j < n; ++j) {
    bool start_flag = false, end_flag = false;
    const V3 s{map[6 * j], map[6 * j + 1], map[6 * j + 2]}, e{map[6 * j + 3], map[6 * j + 4], map[6 * j + 5]};
    const V3 ts = add(mul(cp.R, s), cp.T), te = add(mul(cp.R, e), cp.T);                    // :419-420
    if ((ts.z > 0) && (te.z > 0)) {                                                         // :423
      const double xx = cfg->fx * ts.x / ts.z + cfg->cx, yy = cfg->fy * ts.y / ts.z + cfg->cy;
      const double xx_ = cfg->fx * te.x / te.z + cfg->cx, yy_ = cfg->fy * te.y / te.z + cfg->cy;
      if (xx > width_left && xx < (width_right - 1) && yy > height_up && yy < (height_down)) start_flag = true;
      if (xx_ > width_left && xx_ < (width_right - 1) && yy_ > height_up && yy_ < (height_down)) end_flag = true;
    }
    if (start_flag || end_flag) {                                                           // :440
      if (out) out[cnt] = (int32_t)j;
      if (mask) mask[j >> 5] |= 1u << (j & 31);
      ++cnt;
    }
  }
  return cnt;
}

int correspondence(const viml_config* cfg, const CamPose& cp, const double* map, const int32_t* fov,
                   int fov_n, const double* l2d, float* err, double* projected) {
  const Line2D detectLine = make_line2d(l2d[0], l2d[1], l2d[2], l2d[3]);
  const int width = cfg->width, height = cfg->height;
  int choose_index = -1;
  float error[3] = {0.f, 0.f, 0.f};
  float min_dist = 10000.0;                                                                 // :701
  Line2D projectedLine{};
  if (fov_n == 0) {                                                                         // :703-713
    err[0] = err[1] = err[2] = -1;
    if (projected)                                                                          // :709 returns detectLine itself
      projected[0] = l2d[0], projected[1] = l2d[1], projected[2] = l2d[2], projected[3] = l2d[3];
    return -1;
  }
  for (int i = 0; i < fov_n; ++i) {                                                         // :715
    float overlap = 0.0, distance = 10000.0;
    const int64_t j = fov[i];
    const V3 s{map[6 * j], map[6 * j + 1], map[6 * j + 2]}, e{map[6 * j + 3], map[6 * j + 4], map[6 * j + 5]};
    bool start_flag = false, end_flag = false;
    const V3 ts = add(mul(cp.R, s), cp.T), te = add(mul(cp.R, e), cp.T);                    // :727-728
    float xx = 0, yy = 0, xx_ = 0, yy_ = 0;                                                 // :730
    if (ts.z > 0 && te.z > 0) {
      xx = cfg->fx * ts.x / ts.z + cfg->cx;                                                 // double -> float
      yy = cfg->fy * ts.y / ts.z + cfg->cy;
      xx_ = cfg->fx * te.x / te.z + cfg->cx;
      yy_ = cfg->fy * te.y / te.z + cfg->cy;
      if (xx > 0 && xx < width - 1 && yy > 0 && yy < height - 1) start_flag = true;         // :739
      if (xx_ > 0 && xx_ < width - 1 && yy_ > 0 && yy_ < height - 1) end_flag = true;
    }
    Line2D temp_line{};
    bool have = false;
    if (start_flag && end_flag) {                                                           // :745
      temp_line = make_line2d(xx, yy, xx_, yy_);
      have = true;
    } else if (start_flag && (!end_flag)) {                                                 // :768
      const V3 dirvec = sub(te, ts);
      double t = 0.9;
      bool found = false;
      double x = 0.0, y = 0.0;
      while (t > 0) {
        const V3 p{ts.x + t * dirvec.x, ts.y + t * dirvec.y, ts.z + t * dirvec.z};
        if (p.z > 0) {
          x = cfg->fx * p.x / p.z + cfg->cx;
          y = cfg->fy * p.y / p.z + cfg->cy;
          if (x > 0 && x < (width - 1) && y > 0 && y < (height - 1)) {
            found = true;
            break;
          } else
            t = t - 0.1;
        } else
          t = t - 0.1;
      }
      if (found) {
        temp_line = make_line2d(xx, yy, x, y);                                              // :796
        have = true;
      }
    } else if (end_flag && (!start_flag)) {                                                 // :817
      const V3 dirvec = sub(ts, te);
      double t = 0.9;
      bool found = false;
      double x = 0.0, y = 0.0;
      while (t > 0) {
        const V3 p{te.x + t * dirvec.x, te.y + t * dirvec.y, te.z + t * dirvec.z};
        if (p.z > 0) {
          x = cfg->fx * p.x / p.z + cfg->cx;
          y = cfg->fy * p.y / p.z + cfg->cy;
          if (x > 0 && x < (width - 1) && y > 0 && y < (height - 1)) {
            found = true;
            break;
          } else
            t = t - 0.1;
        } else
          t = t - 0.1;
      }
      if (found) {
        temp_line = make_line2d(x, y, xx_, yy_);                                            // :845
        have = true;
      }
    }
    if (!have) continue;
    const double angle = cal_angle_dist(temp_line, detectLine);                             // :749
    if (angle > cfg->angle_th) continue;
    double d, o;
    cal_euler_dist(temp_line, detectLine, &d, &o);                                          // :752
    distance = d;
    overlap = o;
    if (overlap < cfg->overlap_th) continue;                                                // :756
    if (distance < min_dist) {                                                              // :758
      min_dist = distance;
      choose_index = i;
      projectedLine = temp_line;
      error[0] = angle;
      error[1] = min_dist;
      error[2] = overlap;
    }
  }
  if (choose_index == -1) {                                                                 // :869
    err[0] = err[1] = err[2] = -1;
    if (projected)                                                                          // :874 returns detectLine itself
      projected[0] = l2d[0], projected[1] = l2d[1], projected[2] = l2d[2], projected[3] = l2d[3];
    return -1;
  }
  err[0] = error[0], err[1] = error[1], err[2] = error[2];
  if (projected) {
    projected[0] = projectedLine.Sx, projected[1] = projectedLine.Sy;
    projected[2] = projectedLine.Ex, projected[3] = projectedLine.Ey;
  }
  return fov[choose_index];
}

}  // namespace

extern "C" int orc_update_lines_in_fov(const viml_config* cfg, const double* pose, const double* ex,
                                       const double* map, int64_t n, int32_t* out_index) {
  const CamPose cp = camera_pose(cfg, pose, ex);
  return fov_cull(cfg, cp, map, n, out_index, nullptr);
}

extern "C" int orc_line_correspondence(const viml_config* cfg, const double* pose, const double* ex,
                                       const double* map, const int32_t* fov_index, int fov_count,
                                       const double* line2d, float* err, double* projected) {
  const CamPose cp = camera_pose(cfg, pose, ex);
  return correspondence(cfg, cp, map, fov_index, fov_count, line2d, err, projected);
}

extern "C" int orc_line_associate(const viml_config* cfg, const double* map, int64_t n,
                                  const viml_assoc_query* q, const viml_assoc_out* out, int nthreads) {
  if (!cfg || !map || !q || !out) return VIML_ERR_INVALID;
  const int L = q->lines_per_pose;
  const int64_t words = (n + 31) / 32;
  parallel_for(q->n_poses, nthreads, [&](int p) {
    std::vector<int32_t> fov(n);
    const double* ex = q->ex_pose + (size_t)p * 7;
    const double* cex = q->cull_ex_pose ? q->cull_ex_pose + (size_t)p * 7 : ex;
    const CamPose cull = camera_pose(cfg, q->cull_poses + (size_t)p * 7, cex);
    uint32_t* mask = out->fov_mask ? out->fov_mask + (size_t)p * words : nullptr;
    if (mask) std::fill(mask, mask + words, 0u);
    const int cnt = fov_cull(cfg, cull, map, n, fov.data(), mask);
    if (out->fov_count) out->fov_count[p] = cnt;
    if (out->fov_index)
      for (int k = 0; k < std::min(cnt, out->fov_capacity); ++k)
        out->fov_index[(size_t)p * out->fov_capacity + k] = fov[k];
    const CamPose mp = (q->match_poses || q->cull_ex_pose)
                           ? camera_pose(cfg, (q->match_poses ? q->match_poses : q->cull_poses) + (size_t)p * 7, ex)
                           : cull;
    const int nl = q->n_lines2d ? q->n_lines2d[p] : L;
    for (int l = 0; l < nl; ++l) {
      float err[3];
      double proj[4] = {0, 0, 0, 0};
      const int idx = correspondence(cfg, mp, map, fov.data(), cnt, q->lines2d + ((size_t)p * L + l) * 4, err, proj);
      const size_t o = (size_t)p * L + l;
      if (out->match_index) out->match_index[o] = idx;
      if (out->err) out->err[3 * o] = err[0], out->err[3 * o + 1] = err[1], out->err[3 * o + 2] = err[2];
      if (out->projected) std::memcpy(out->projected + 4 * o, proj, sizeof(proj));
    }
  });
  return VIML_OK;
}

// a17  FeatureManager::removeLineOutlier / lineDiff — feature_manager.cpp:494-541
extern "C" void orc_line2d(const double* seg, double* out) {
  const Line2D l = make_line2d(seg[0], seg[1], seg[2], seg[3]);
  out[0] = l.A, out[1] = l.B, out[2] = l.C, out[3] = l.A2B2, out[4] = l.Length, out[5] = l.Dx, out[6] = l.Dy;
}
void orc_point2flined(const double* seg, const double* p, double* out) {
  const Line2D l = make_line2d(seg[0], seg[1], seg[2], seg[3]);
  point2flined(l, p[0], p[1], &out[0], &out[1]);
}

int orc_track_gate(int n_obs, const double* line_vec, uint8_t* credible_line) {
  if (n_obs < 1) return 1;
  int count = 0;
  for (int k = 0; k < n_obs; ++k) {
    const double dx = line_vec[0] - line_vec[3 * k], dy = line_vec[1] - line_vec[3 * k + 1],
                 dz = line_vec[2] - line_vec[3 * k + 2];
    const float diff_ = (float)std::sqrt((dx * dx + dy * dy) + dz * dz);                    // :536-540
    if (diff_ > 0.1) {                                                                      // :515
      count++;
      credible_line[k] = 0;
    } else
      credible_line[k] = 1;
  }
  return ((count / n_obs) >= 0.5) ? 0 : 1;                                                  // :524 integer division
}

extern "C" double orc_cos_threshold(double angle_th) {
  // smallest double c in [0,1] with acos(c) <= angle_th (acos is monotone decreasing)
  if (!(std::acos(1.0) <= angle_th)) return INFINITY;
  if (std::acos(0.0) <= angle_th) return 0.0;
  uint64_t lo, hi;  // lo: rejected, hi: accepted; positive doubles order like their bit patterns
  double dlo = 0.0, dhi = 1.0;
  std::memcpy(&lo, &dlo, 8);
  std::memcpy(&hi, &dhi, 8);
  while (hi - lo > 1) {
    const uint64_t mid = lo + (hi - lo) / 2;
    double dm;
    std::memcpy(&dm, &mid, 8);
    if (std::acos(dm) <= angle_th)
      hi = mid;
    else
      lo = mid;
  }
  double r;
  std::memcpy(&r, &hi, 8);
  return r;
}

extern "C" int orc_hardware_threads(void) {
  const unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}
