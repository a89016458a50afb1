// ref_wrap.cpp — C entry points around the REFERENCE's own factor classes (TEST INFRASTRUCTURE ONLY).
//
// oracle/_ref/libref.so = this file + /root/reference/vins_estimator/src/factor/{projection_factor, line_projection_factor,
// pose_local_parameterization, marginalization_factor}.cpp compiled UNMODIFIED, where they lie, against the interface stand-ins
// under oracle/ref_shim/ (Eigen, Ceres and ROS are not installed in this image; see the headers there for what they restate).
// No reference source is copied into this repository.  Used by tests/ to check oracle/viml_oracle.cpp against the reference's
// statement sequence, and by tests/golden/make_ref_golden.py to write reference-generated vectors.
#include <cstring>
#include <unordered_map>
#include <vector>

#include "factor/line_projection_factor.h"
#include "factor/marginalization_factor.h"
#include "factor/pose_local_parameterization.h"
#include "factor/projection_factor.h"
#include "feature_manager.h"

// globals the reference declares extern in parameters.h and defines in parameters.cpp (not compiled here)
double INIT_DEPTH = 5.0;
double MIN_PARALLAX = 10.0 / 460.0;

extern "C" {

// ProjectionFactor::Evaluate (projection_factor.cpp:21-124); jac[k] may be NULL like in Ceres
void ref_projection_evaluate(const double* pts_i, const double* pts_j, double sqrt_info, const double* const* params, double* residuals,
                             double** jac) {
  ProjectionFactor::sqrt_info = sqrt_info * Eigen::Matrix2d::Identity();   // estimator.cpp:48, :85
  ProjectionFactor f(Eigen::Vector3d(pts_i[0], pts_i[1], pts_i[2]), Eigen::Vector3d(pts_j[0], pts_j[1], pts_j[2]));
  f.Evaluate(params, residuals, jac);
}

// LineProjectionFactor::Evaluate (line_projection_factor.cpp:19-120); K, bcR row-major 3x3
void ref_line_evaluate(const double* ps, const double* pe, const double* abc, const double* K, const double* bcR, const double* bcT,
                       const double* const* params, double* residuals, double** jac) {
  Eigen::Matrix3d Km, Rm;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Km(r, c) = K[3 * r + c], Rm(r, c) = bcR[3 * r + c];
  LineProjectionFactor f(Eigen::Vector3d(ps[0], ps[1], ps[2]), Eigen::Vector3d(pe[0], pe[1], pe[2]), Eigen::Vector3d(abc[0], abc[1], abc[2]), Km,
                         Rm, Eigen::Vector3d(bcT[0], bcT[1], bcT[2]));
  f.Evaluate(params, residuals, jac);
}

// PoseLocalParameterization::Plus (pose_local_parameterization.cpp:3-19); Plus is private in the reference class, the
// ceres::LocalParameterization interface it overrides is public
void ref_pose_plus(const double* x, const double* delta, double* out) {
  PoseLocalParameterization p;
  static_cast<ceres::LocalParameterization&>(p).Plus(x, delta, out);
}

// MarginalizationInfo (marginalization_factor.cpp:89-299) on the MARGIN_OLD factor set of projection factors
// (estimator.cpp:1961-1990): every factor (i = 0, j, feature l) with drop set {0, 3}; CauchyLoss(cauchy_a) when use_loss.
//   poses [P][7], ex [7], inv_depth [F]; factors: fj[k], fl[k], obs[k] = {pts_i.xy, pts_j.xy}, pts z = 1.
// Outputs: n, keep_order[P] = for each kept block in the reference's keep order the pose index (1..P-1) or P for the extrinsic,
// keep_idx[P] = its offset inside the kept part, lin_jac [n][n] row-major, lin_res [n].  Returns n.
int ref_marginalize_old(int P, int F, double* poses, double* ex, double* inv_depth, int NF, const int* fj, const int* fl, const double* obs,
                        double sqrt_info, int use_loss, double cauchy_a, int* keep_order, int* keep_idx, double* lin_jac, double* lin_res,
                        int* m_out) {
  ProjectionFactor::sqrt_info = sqrt_info * Eigen::Matrix2d::Identity();
  ceres::LossFunction* loss = use_loss ? new ceres::CauchyLoss(cauchy_a) : nullptr;
  MarginalizationInfo* info = new MarginalizationInfo();
  for (int k = 0; k < NF; ++k) {
    ProjectionFactor* f = new ProjectionFactor(Eigen::Vector3d(obs[4 * k], obs[4 * k + 1], 1.0), Eigen::Vector3d(obs[4 * k + 2], obs[4 * k + 3], 1.0));
    ResidualBlockInfo* rbi = new ResidualBlockInfo(f, loss, std::vector<double*>{poses, poses + 7 * fj[k], ex, inv_depth + fl[k]},
                                                   std::vector<int>{0, 3});
    info->addResidualBlockInfo(rbi);
  }
  info->preMarginalize();
  info->marginalize();
  std::unordered_map<long, double*> addr_shift;
  for (int p = 1; p < P; ++p) addr_shift[reinterpret_cast<long>(poses + 7 * p)] = poses + 7 * p;
  addr_shift[reinterpret_cast<long>(ex)] = ex;
  std::vector<double*> keep = info->getParameterBlocks(addr_shift);
  const int n = info->n;
  for (size_t b = 0; b < keep.size(); ++b) {
    keep_order[b] = keep[b] == ex ? P : (int)((keep[b] - poses) / 7);
    keep_idx[b] = info->keep_block_idx[b] - info->m;
  }
  for (int r = 0; r < n; ++r) {
    lin_res[r] = info->linearized_residuals(r);
    for (int c = 0; c < n; ++c) lin_jac[(size_t)r * n + c] = info->linearized_jacobians(r, c);
  }
  *m_out = info->m;
  const int nkeep = (int)keep.size();
  delete info;   // frees the factors and their cost functions (marginalization_factor.cpp:71-87)
  delete loss;
  (void)F;
  return n * 1000 + nkeep;
}

// Line2D::Line2D(Vector4d) (feature_manager.cpp:4-15): out = {A, B, C, A2B2, Length, Direction.x, Direction.y}
void ref_line2d(const double* seg, double* out) {
  Line2D l(Eigen::Vector4d(seg[0], seg[1], seg[2], seg[3]));
  out[0] = l.A, out[1] = l.B, out[2] = l.C, out[3] = l.A2B2, out[4] = l.Length, out[5] = l.Direction(0), out[6] = l.Direction(1);
}

// Line2D::Point2Flined (feature_manager.cpp:46-71)
void ref_point2flined(const double* seg, const double* p, double* out) {
  Line2D l(Eigen::Vector4d(seg[0], seg[1], seg[2], seg[3]));
  const Eigen::Vector2d r = l.Point2Flined(Eigen::Vector2d(p[0], p[1]));
  out[0] = r(0), out[1] = r(1);
}

// FeatureManager::triangulate (feature_manager.cpp:440-492) for NF features of one window: feature l starts in frame start[l] and
// has observations off[l] .. off[l+1]-1 (normalised points xyz).  poses [P][7] (the estimator's Ps / Rs = normalised quaternion),
// ex [7].  depth[l] = estimated_depth after the call.
void ref_triangulate(int P, const double* poses, const double* ex, int NF, const int* start, const long long* off, const double* pts,
                     double* depth) {
  std::vector<Eigen::Matrix3d> Rs(P + 1);
  std::vector<Eigen::Vector3d> Ps(P + 1);
  for (int p = 0; p < P; ++p) {
    const double* q = poses + 7 * p;
    Ps[p] = Eigen::Vector3d(q[0], q[1], q[2]);
    Rs[p] = Eigen::Quaterniond(q[6], q[3], q[4], q[5]).normalized().toRotationMatrix();   // estimator.cpp double2vector
  }
  Eigen::Vector3d tic[1] = {Eigen::Vector3d(ex[0], ex[1], ex[2])};
  Eigen::Matrix3d ric[1] = {Eigen::Quaterniond(ex[6], ex[3], ex[4], ex[5]).normalized().toRotationMatrix()};
  FeatureManager fm(Rs.data());
  for (int l = 0; l < NF; ++l) {
    fm.feature.push_back(FeaturePerId(l, start[l]));
    for (long long k = off[l]; k < off[l + 1]; ++k) {
      Eigen::Matrix<double, 7, 1> v;
      v << pts[3 * k], pts[3 * k + 1], pts[3 * k + 2], 0, 0, 0, 0;
      fm.feature.back().feature_per_frame.push_back(FeaturePerFrame(v, 0.0));
    }
  }
  fm.triangulate(Ps.data(), tic, ric);
  int l = 0;
  for (auto& it : fm.feature) depth[l++] = it.estimated_depth;
}

}  // extern "C"
