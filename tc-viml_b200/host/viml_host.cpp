// viml_host.cpp — see viml_host.h.  Reference lines are cited as file:line relative to
// /root/reference/vins_estimator/src.
#include "viml_host.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace {

[[noreturn]] void die(const char* what) {
  // the reference's failure mode on this path is ROS_BREAK() (marginalization_factor.cpp:250-254)
  std::fprintf(stderr, "viml_host: %s\n", what);
  std::abort();
}

viml_ctx* need_ctx() {
  viml_ctx* c = viml::Runtime::instance().ctx();
  if (!c) die("viml::Runtime::configure() has not been called or failed; there is no CPU fallback for the hot path");
  return c;
}

void check_rc(int rc, const char* where) {
  if (rc != VIML_OK) {
    std::fprintf(stderr, "viml_host: %s failed (%d): %s\n", where, rc, viml_last_error(viml::Runtime::instance().ctx()));
    std::abort();
  }
}

double scalar_sqrt_info() {
  const viml::Matrix2d& s = ProjectionFactor::sqrt_info;
  if (s(0, 0) != s(1, 1) || s(0, 1) != 0.0 || s(1, 0) != 0.0)
    die("ProjectionFactor::sqrt_info must be s*I (estimator.cpp:85 sets FOCAL_LENGTH/1.5*I)");
  if (s(0, 0) != viml::Runtime::instance().config().sqrt_info)
    die("ProjectionFactor::sqrt_info differs from viml_config::sqrt_info; call Runtime::configure after changing it");
  return s(0, 0);
}

inline int local_size(int size) { return size == 7 ? 6 : size; }

}  // namespace

// ---------------------------------------------------------------------------------------------- Runtime
namespace viml {
Runtime& Runtime::instance() {
  static Runtime r;
  return r;
}
int Runtime::configure(const viml_config& cfg, int device) {
  shutdown();
  cfg_ = cfg;
  const int rc = viml_create(&ctx_, &cfg_, device);
  if (rc != VIML_OK) ctx_ = nullptr;
  ProjectionFactor::sqrt_info = cfg.sqrt_info * Matrix2d::Identity();
  return rc;
}
void Runtime::shutdown() {
  if (ctx_) viml_destroy(ctx_);
  ctx_ = nullptr;
}
Runtime::~Runtime() { shutdown(); }
}  // namespace viml

// ---------------------------------------------------------------------------------------------- factors
viml::Matrix2d ProjectionFactor::sqrt_info;
double ProjectionFactor::sum_t;
double LineProjectionFactor::sum_t;

ProjectionFactor::ProjectionFactor(const viml::Vector3d& _pts_i, const viml::Vector3d& _pts_j) : pts_i(_pts_i), pts_j(_pts_j) {}

bool ProjectionFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  if (batch_ && batch_->fetch(this, parameters, residuals, jacobians)) return true;
  // batch of one (needed by ResidualBlockInfo::Evaluate and by any caller outside a prepared batch)
  viml_ctx* ctx = need_ctx();
  scalar_sqrt_info();
  double poses[14], ex[7], lam[1] = {parameters[3][0]};
  std::memcpy(poses, parameters[0], 56);
  std::memcpy(poses + 7, parameters[1], 56);
  std::memcpy(ex, parameters[2], 56);
  const int32_t off[2] = {0, 1};
  const uint32_t idx[1] = {0u | (1u << 8)};
  const double obs[4] = {pts_i.x(), pts_i.y(), pts_j.x(), pts_j.y()};
  const double zi[1] = {pts_i.z()};
  viml_window_batch in{};
  in.n_windows = 1, in.poses_per_window = 2, in.feats_per_window = 1;
  in.poses = poses, in.ex_pose = ex, in.inv_depth = lam;
  in.n_point_factors = 1, in.pf_window_offset = off, in.pf_idx = idx, in.pf_obs = obs, in.pf_pts_i_z = zi;
  double r[2], J[4][14];
  viml_linearize_out out{};
  out.pf_residual = r;
  if (jacobians) out.pf_jac_pose_i = J[0], out.pf_jac_pose_j = J[1], out.pf_jac_ex = J[2], out.pf_jac_feat = J[3];
  check_rc(viml_linearize_batch(ctx, &in, &out, VIML_OUT_RESIDUAL_JACOBIAN), "viml_linearize_batch");
  residuals[0] = r[0], residuals[1] = r[1];
  if (jacobians) {
    const int sz[4] = {14, 14, 14, 2};
    for (int k = 0; k < 4; ++k)
      if (jacobians[k]) std::memcpy(jacobians[k], J[k], sizeof(double) * sz[k]);  // NULL blocks are skipped (:77,:89,:100,:111)
  }
  return true;  // projection_factor.cpp:123
}

double ProjectionFactor::check(double** parameters) {
  // projection_factor.cpp:126-228: eps 1e-6, Q <- Q*deltaQ(d), columns [Pi,Qi,Pj,Qj,tic,qic,inv_dep]
  double r0[2], J[4][14];
  double* jp[4] = {J[0], J[1], J[2], J[3]};
  Evaluate(parameters, r0, jp);
  const double eps = 1e-6;
  double worst = 0.0, scale = 1e-300;
  for (int k = 0; k < 19; ++k) {
    double P[3][7], lam = parameters[3][0];
    for (int b = 0; b < 3; ++b) std::memcpy(P[b], parameters[b], 56);
    const int a = k / 6, bb = k % 6;
    if (a < 3) {
      if (bb < 3) {
        P[a][bb] += eps;
      } else {
        double d[3] = {0, 0, 0};
        d[bb - 3] = eps;
        const double qx = P[a][3], qy = P[a][4], qz = P[a][5], qw = P[a][6];
        const double dx = d[0] / 2, dy = d[1] / 2, dz = d[2] / 2, dw = 1.0;  // Utility::deltaQ
        P[a][6] = qw * dw - qx * dx - qy * dy - qz * dz;
        P[a][3] = qw * dx + qx * dw + qy * dz - qz * dy;
        P[a][4] = qw * dy + qy * dw + qz * dx - qx * dz;
        P[a][5] = qw * dz + qz * dw + qx * dy - qy * dx;
      }
    } else {
      lam += eps;
    }
    const double* pp[4] = {P[0], P[1], P[2], &lam};
    double r1[2];
    Evaluate(pp, r1, nullptr);
    for (int row = 0; row < 2; ++row) {
      const double num = (r1[row] - r0[row]) / eps;
      const double ana = a < 3 ? J[a][7 * row + bb] : J[3][row];
      worst = std::max(worst, std::fabs(num - ana));
      scale = std::max(scale, std::fabs(ana));
    }
  }
  return worst / scale;
}

LineProjectionFactor::LineProjectionFactor(const viml::Vector3d& _pts_start, const viml::Vector3d& _pts_end,
                                           const viml::Vector3d& _line_param, const viml::Matrix3d _K,
                                           const viml::Matrix3d _b_c_R, const viml::Vector3d _b_c_T)
    : pts_start(_pts_start), pts_end(_pts_end), line_param(_line_param), K(_K), b_c_R(_b_c_R), b_c_T(_b_c_T) {}

namespace {
// b_c_R is a rotation matrix frozen at problem-build time (estimator.cpp:1777-1781); the C-ABI takes the
// extrinsic as a pose, so convert back to a unit quaternion (w >= 0).
void rot_to_quat(const viml::Matrix3d& R, double* q /*x,y,z,w*/) {
  const double tr = R(0, 0) + R(1, 1) + R(2, 2);
  double x, y, z, w;
  if (tr > 0) {
    const double s = std::sqrt(tr + 1.0) * 2;
    w = 0.25 * s, x = (R(2, 1) - R(1, 2)) / s, y = (R(0, 2) - R(2, 0)) / s, z = (R(1, 0) - R(0, 1)) / s;
  } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
    const double s = std::sqrt(1.0 + R(0, 0) - R(1, 1) - R(2, 2)) * 2;
    w = (R(2, 1) - R(1, 2)) / s, x = 0.25 * s, y = (R(0, 1) + R(1, 0)) / s, z = (R(0, 2) + R(2, 0)) / s;
  } else if (R(1, 1) > R(2, 2)) {
    const double s = std::sqrt(1.0 + R(1, 1) - R(0, 0) - R(2, 2)) * 2;
    w = (R(0, 2) - R(2, 0)) / s, x = (R(0, 1) + R(1, 0)) / s, y = 0.25 * s, z = (R(1, 2) + R(2, 1)) / s;
  } else {
    const double s = std::sqrt(1.0 + R(2, 2) - R(0, 0) - R(1, 1)) * 2;
    w = (R(1, 0) - R(0, 1)) / s, x = (R(0, 2) + R(2, 0)) / s, y = (R(1, 2) + R(2, 1)) / s, z = 0.25 * s;
  }
  q[0] = x, q[1] = y, q[2] = z, q[3] = w;
}
void check_K(const viml::Matrix3d& K) {
  const viml_config& c = viml::Runtime::instance().config();
  if (K(0, 0) != c.fx || K(1, 1) != c.fy || K(0, 2) != c.cx || K(1, 2) != c.cy)
    die("LineProjectionFactor::K differs from viml_config intrinsics");
}
}  // namespace

bool LineProjectionFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  if (batch_ && batch_->fetch(this, parameters, residuals, jacobians)) return true;
  viml_ctx* ctx = need_ctx();
  check_K(K);
  double pose[7], ex[7], lam[1] = {1.0};
  std::memcpy(pose, parameters[0], 56);
  ex[0] = b_c_T.x(), ex[1] = b_c_T.y(), ex[2] = b_c_T.z();
  rot_to_quat(b_c_R, ex + 3);
  const int32_t poff[2] = {0, 0}, loff[2] = {0, 1}, frame[1] = {0};
  const double geom[9] = {pts_start.x(), pts_start.y(), pts_start.z(), pts_end.x(), pts_end.y(), pts_end.z(),
                          line_param.x(), line_param.y(), line_param.z()};
  viml_window_batch in{};
  in.n_windows = 1, in.poses_per_window = 1, in.feats_per_window = 1;
  in.poses = pose, in.ex_pose = ex, in.inv_depth = lam;
  in.n_point_factors = 0, in.pf_window_offset = poff;
  in.n_line_factors = 1, in.lf_window_offset = loff, in.lf_frame = frame, in.lf_geom = geom;
  double r[2], J[14];
  viml_linearize_out out{};
  out.lf_residual = r;
  if (jacobians) out.lf_jac_pose = J;
  check_rc(viml_linearize_batch(ctx, &in, &out, VIML_OUT_RESIDUAL_JACOBIAN), "viml_linearize_batch");
  residuals[0] = r[0], residuals[1] = r[1];
  if (jacobians) std::memcpy(jacobians[0], J, sizeof(J));  // written without a NULL check, like :102
  return true;  // line_projection_factor.cpp:119
}

// ---------------------------------------------------------------------------------------------- batch
LinearizationBatch::~LinearizationBatch() {
  for (auto& e : pf_) e.f->batch_ = nullptr;
  for (auto& e : lf_) e.f->batch_ = nullptr;
  free_pinned();
}
void LinearizationBatch::free_pinned() {
  for (double** p : {&r_pf_, &ji_, &jj_, &je_, &jl_, &r_lf_, &jp_}) {
    if (*p) viml_host_free(*p);
    *p = nullptr;
  }
  cap_pf_ = cap_lf_ = 0;
}
int LinearizationBatch::pose_index(double* p) {
  auto it = pose_id_.find(p);
  if (it != pose_id_.end()) return it->second;
  const int id = (int)poses_.size();
  poses_.push_back(p);
  pose_id_[p] = id;
  return id;
}
void LinearizationBatch::AddResidualBlock(ProjectionFactor* f, double* pose_i, double* pose_j, double* ex_pose, double* feature) {
  if (ex_ && ex_ != ex_pose) die("LinearizationBatch: one extrinsic block per batch (NUM_OF_CAM = 1, parameters.h:21)");
  ex_ = ex_pose;
  int l;
  auto it = feat_id_.find(feature);
  if (it == feat_id_.end()) {
    l = (int)feats_.size();
    feats_.push_back(feature);
    feat_id_[feature] = l;
  } else {
    l = it->second;
  }
  f->batch_ = this;
  f->slot_ = (int)pf_.size();
  pf_.push_back({f, pose_i, pose_j, ex_pose, feature, pose_index(pose_i), pose_index(pose_j), l});
  valid_ = false;
}
void LinearizationBatch::AddResidualBlock(LineProjectionFactor* f, double* pose) {
  f->batch_ = this;
  f->slot_ = (int)lf_.size();
  lf_.push_back({f, pose, pose_index(pose)});
  valid_ = false;
}

void LinearizationBatch::PrepareForEvaluation(bool /*evaluate_jacobians*/, bool /*new_evaluation_point*/) {
  viml_ctx* ctx = need_ctx();
  scalar_sqrt_info();
  const int P = (int)poses_.size(), F = std::max<int>(1, (int)feats_.size());
  const size_t NP = pf_.size(), NL = lf_.size();
  if (P > 255 || F > 65535) die("LinearizationBatch: window too large for the packed index");
  std::vector<double> poses((size_t)P * 7), lam(F, 1.0), ex(7, 0.0), obs(NP * 4), zi(NP), geom(9 * NL);
  std::vector<uint32_t> idx(NP);
  std::vector<int32_t> frame(NL);
  for (int p = 0; p < P; ++p) std::memcpy(&poses[7 * p], poses_[p], 56);
  for (size_t l = 0; l < feats_.size(); ++l) lam[l] = feats_[l][0];
  if (ex_) {
    std::memcpy(ex.data(), ex_, 56);
  } else if (NL) {  // line factors only: their frozen extrinsic
    const LineProjectionFactor* f = lf_[0].f;
    ex[0] = f->b_c_T.x(), ex[1] = f->b_c_T.y(), ex[2] = f->b_c_T.z();
    rot_to_quat(f->b_c_R, &ex[3]);
  }
  for (size_t k = 0; k < NP; ++k) {
    const PF& e = pf_[k];
    idx[k] = (uint32_t)e.i | ((uint32_t)e.j << 8) | ((uint32_t)e.l << 16);
    obs[4 * k] = e.f->pts_i.x(), obs[4 * k + 1] = e.f->pts_i.y(), obs[4 * k + 2] = e.f->pts_j.x(), obs[4 * k + 3] = e.f->pts_j.y();
    zi[k] = e.f->pts_i.z();
  }
  for (size_t k = 0; k < NL; ++k) {
    const LineProjectionFactor* f = lf_[k].f;
    frame[k] = lf_[k].frame;
    const double g[9] = {f->pts_start.x(), f->pts_start.y(), f->pts_start.z(), f->pts_end.x(), f->pts_end.y(),
                         f->pts_end.z(), f->line_param.x(), f->line_param.y(), f->line_param.z()};
    for (int c = 0; c < 9; ++c) geom[(size_t)c * NL + k] = g[c];
  }
  if (NP > cap_pf_ || NL > cap_lf_) {
    free_pinned();
    cap_pf_ = NP + NP / 4 + 16, cap_lf_ = NL + NL / 4 + 16;
    void* p;
    auto grab = [&](double** dst, size_t n) { check_rc(viml_host_alloc(&p, n * sizeof(double)), "viml_host_alloc"); *dst = (double*)p; };
    grab(&r_pf_, cap_pf_ * 2), grab(&ji_, cap_pf_ * 14), grab(&jj_, cap_pf_ * 14), grab(&je_, cap_pf_ * 14), grab(&jl_, cap_pf_ * 2);
    grab(&r_lf_, cap_lf_ * 2), grab(&jp_, cap_lf_ * 14);
  }
  const int32_t poff[2] = {0, (int32_t)NP}, loff[2] = {0, (int32_t)NL};
  viml_window_batch in{};
  in.n_windows = 1, in.poses_per_window = P, in.feats_per_window = F;
  in.poses = poses.data(), in.ex_pose = ex.data(), in.inv_depth = lam.data();
  in.n_point_factors = (int64_t)NP, in.pf_window_offset = poff, in.pf_idx = idx.data(), in.pf_obs = obs.data(), in.pf_pts_i_z = zi.data();
  in.n_line_factors = (int64_t)NL, in.lf_window_offset = loff, in.lf_frame = frame.data(), in.lf_geom = geom.data();
  viml_linearize_out out{};
  out.pf_residual = r_pf_, out.pf_jac_pose_i = ji_, out.pf_jac_pose_j = jj_, out.pf_jac_ex = je_, out.pf_jac_feat = jl_;
  out.lf_residual = r_lf_, out.lf_jac_pose = jp_;
  last_rc_ = viml_linearize_batch(ctx, &in, &out, VIML_OUT_RESIDUAL_JACOBIAN);
  check_rc(last_rc_, "viml_linearize_batch");
  snap_poses_.swap(poses), snap_lam_.swap(lam), snap_ex_.swap(ex);
  valid_ = true;
}

bool LinearizationBatch::fetch(const ProjectionFactor* f, double const* const* parameters, double* residuals, double** jacobians) const {
  if (!valid_) return false;
  const PF& e = pf_[f->slot_];
  // same evaluation point?  compared by value (56-byte blocks), not by address: see snap_poses_ in viml_host.h
  if (std::memcmp(parameters[0], &snap_poses_[7 * (size_t)e.i], 56) || std::memcmp(parameters[1], &snap_poses_[7 * (size_t)e.j], 56) ||
      std::memcmp(parameters[2], snap_ex_.data(), 56) || std::memcmp(parameters[3], &snap_lam_[(size_t)e.l], 8))
    return false;
  const size_t k = (size_t)f->slot_;
  residuals[0] = r_pf_[2 * k], residuals[1] = r_pf_[2 * k + 1];
  if (jacobians) {
    if (jacobians[0]) std::memcpy(jacobians[0], ji_ + 14 * k, 112);
    if (jacobians[1]) std::memcpy(jacobians[1], jj_ + 14 * k, 112);
    if (jacobians[2]) std::memcpy(jacobians[2], je_ + 14 * k, 112);
    if (jacobians[3]) std::memcpy(jacobians[3], jl_ + 2 * k, 16);
  }
  ++served_;
  return true;
}
bool LinearizationBatch::fetch(const LineProjectionFactor* f, double const* const* parameters, double* residuals, double** jacobians) const {
  if (!valid_) return false;
  const LF& e = lf_[f->slot_];
  if (std::memcmp(parameters[0], &snap_poses_[7 * (size_t)e.frame], 56)) return false;
  const size_t k = (size_t)f->slot_;
  residuals[0] = r_lf_[2 * k], residuals[1] = r_lf_[2 * k + 1];
  if (jacobians) std::memcpy(jacobians[0], jp_ + 14 * k, 112);
  ++served_;
  return true;
}

// ---------------------------------------------------------------------------------------------- marginalisation
namespace {
// Robust-loss correction of ResidualBlockInfo::Evaluate (marginalization_factor.cpp:37-68) for cost functions
// that stay on the host (IMU, prior).  Point factors get the same correction on the device.
void host_loss_correct(ceres::LossFunction* loss, std::vector<double>& r, std::vector<std::vector<double>>& J, const std::vector<int>& sizes) {
  double sq_norm = 0.0, rho[3];
  for (double v : r) sq_norm += v * v;
  loss->Evaluate(sq_norm, rho);
  const double sqrt_rho1_ = std::sqrt(rho[1]);
  double residual_scaling_, alpha_sq_norm_;
  if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
    residual_scaling_ = sqrt_rho1_;
    alpha_sq_norm_ = 0.0;
  } else {
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling_ = sqrt_rho1_ / (1 - alpha);
    alpha_sq_norm_ = alpha / sq_norm;
  }
  const int nres = (int)r.size();
  for (size_t k = 0; k < J.size(); ++k) {
    const int nc = sizes[k];
    std::vector<double> rtJ(nc, 0.0);
    for (int c = 0; c < nc; ++c)
      for (int q = 0; q < nres; ++q) rtJ[c] += r[q] * J[k][(size_t)q * nc + c];
    for (int q = 0; q < nres; ++q)
      for (int c = 0; c < nc; ++c) J[k][(size_t)q * nc + c] = sqrt_rho1_ * (J[k][(size_t)q * nc + c] - alpha_sq_norm_ * r[q] * rtJ[c]);
  }
  for (double& v : r) v *= residual_scaling_;
}
}  // namespace

void ResidualBlockInfo::Evaluate() {
  residuals.assign(cost_function->num_residuals(), 0.0);
  std::vector<int> block_sizes(cost_function->parameter_block_sizes().begin(), cost_function->parameter_block_sizes().end());
  delete[] raw_jacobians;
  raw_jacobians = new double*[block_sizes.size()];
  jacobians.resize(block_sizes.size());
  for (size_t i = 0; i < block_sizes.size(); i++) {
    jacobians[i].assign((size_t)cost_function->num_residuals() * block_sizes[i], 0.0);
    raw_jacobians[i] = jacobians[i].data();
  }
  cost_function->Evaluate(parameter_blocks.data(), residuals.data(), raw_jacobians);
  if (loss_function) host_loss_correct(loss_function, residuals, jacobians, block_sizes);
}

MarginalizationInfo::~MarginalizationInfo() {
  for (auto it = parameter_block_data.begin(); it != parameter_block_data.end(); ++it) delete[] it->second;
  for (int i = 0; i < (int)factors.size(); i++) {  // marginalization_factor.cpp:78-86: owns the cost functions
    delete[] factors[i]->raw_jacobians;
    delete factors[i]->cost_function;
    delete factors[i];
  }
}
int MarginalizationInfo::localSize(int size) const { return size == 7 ? 6 : size; }
int MarginalizationInfo::globalSize(int size) const { return size == 6 ? 7 : size; }

void MarginalizationInfo::addResidualBlockInfo(ResidualBlockInfo* residual_block_info) {
  factors.emplace_back(residual_block_info);
  std::vector<double*>& parameter_blocks = residual_block_info->parameter_blocks;
  const std::vector<int32_t>& sizes = residual_block_info->cost_function->parameter_block_sizes();
  for (int i = 0; i < static_cast<int>(parameter_blocks.size()); i++) {
    const long addr = reinterpret_cast<long>(parameter_blocks[i]);
    if (parameter_block_size.find(addr) == parameter_block_size.end()) order_.push_back(addr);
    parameter_block_size[addr] = sizes[i];
  }
  for (int i = 0; i < static_cast<int>(residual_block_info->drop_set.size()); i++) {
    double* addr = parameter_blocks[residual_block_info->drop_set[i]];
    parameter_block_idx[reinterpret_cast<long>(addr)] = 0;
  }
}

void MarginalizationInfo::preMarginalize() {
  viml_ctx* ctx = need_ctx();
  // split: ProjectionFactors go to the device in one batch, everything else through its own Evaluate
  std::vector<ResidualBlockInfo*> pts;
  for (auto it : factors) {
    if (dynamic_cast<ProjectionFactor*>(it->cost_function)) pts.push_back(it);
    else it->Evaluate();
    const std::vector<int32_t>& block_sizes = it->cost_function->parameter_block_sizes();
    for (int i = 0; i < static_cast<int>(block_sizes.size()); i++) {  // :117-127 snapshot of the linearisation point
      const long addr = reinterpret_cast<long>(it->parameter_blocks[i]);
      if (parameter_block_data.find(addr) == parameter_block_data.end()) {
        double* data = new double[block_sizes[i]];
        std::memcpy(data, it->parameter_blocks[i], sizeof(double) * block_sizes[i]);
        parameter_block_data[addr] = data;
      }
    }
  }
  gpu_pose_blocks_.clear();
  gpu_ex_block_ = nullptr;
  gpu_landmarks_ = 0;
  S_.clear(), g_.clear();
  if (pts.empty()) return;
  scalar_sqrt_info();
  std::unordered_map<double*, int> pose_id, feat_id;
  std::vector<double*> feats;
  const size_t NP = pts.size();
  std::vector<uint32_t> idx(NP);
  std::vector<double> obs(NP * 4), zi(NP);
  const bool with_loss = pts[0]->loss_function != nullptr;
  for (size_t k = 0; k < NP; ++k) {
    ResidualBlockInfo* it = pts[k];
    const ProjectionFactor* f = static_cast<ProjectionFactor*>(it->cost_function);
    if ((it->loss_function != nullptr) != with_loss) die("MarginalizationInfo: mixed loss functions on ProjectionFactors");
    if (with_loss) {
      auto* cl = dynamic_cast<ceres::CauchyLoss*>(it->loss_function);
      if (!cl) die("MarginalizationInfo: only ceres::CauchyLoss is supported on the device (estimator.cpp:1682)");
    }
    bool drops_feature = false;
    for (int d : it->drop_set) drops_feature |= d == 3;
    if (!drops_feature) die("MarginalizationInfo: ProjectionFactor must marginalise its feature (estimator.cpp:1979 drop set {0,3})");
    int ids[2];
    for (int b = 0; b < 2; ++b) {
      double* p = it->parameter_blocks[b];
      auto pit = pose_id.find(p);
      if (pit == pose_id.end()) {
        ids[b] = (int)gpu_pose_blocks_.size();
        gpu_pose_blocks_.push_back(p);
        pose_id[p] = ids[b];
      } else {
        ids[b] = pit->second;
      }
    }
    if (gpu_ex_block_ && gpu_ex_block_ != it->parameter_blocks[2]) die("MarginalizationInfo: more than one extrinsic block");
    gpu_ex_block_ = it->parameter_blocks[2];
    double* fp = it->parameter_blocks[3];
    auto fit = feat_id.find(fp);
    int l;
    if (fit == feat_id.end()) {
      l = (int)feats.size();
      feats.push_back(fp);
      feat_id[fp] = l;
    } else {
      l = fit->second;
    }
    idx[k] = (uint32_t)ids[0] | ((uint32_t)ids[1] << 8) | ((uint32_t)l << 16);
    obs[4 * k] = f->pts_i.x(), obs[4 * k + 1] = f->pts_i.y(), obs[4 * k + 2] = f->pts_j.x(), obs[4 * k + 3] = f->pts_j.y();
    zi[k] = f->pts_i.z();
  }
  const int P = (int)gpu_pose_blocks_.size(), F = (int)feats.size(), D = 6 * (P + 1);
  gpu_landmarks_ = F;
  std::vector<double> poses((size_t)P * 7), lam(F), ex(7);
  for (int p = 0; p < P; ++p) std::memcpy(&poses[7 * p], gpu_pose_blocks_[p], 56);
  for (int l = 0; l < F; ++l) lam[l] = feats[l][0];
  std::memcpy(ex.data(), gpu_ex_block_, 56);
  const int32_t poff[2] = {0, (int32_t)NP};
  viml_window_batch in{};
  in.n_windows = 1, in.poses_per_window = P, in.feats_per_window = F;
  in.poses = poses.data(), in.ex_pose = ex.data(), in.inv_depth = lam.data();
  in.n_point_factors = (int64_t)NP, in.pf_window_offset = poff, in.pf_idx = idx.data(), in.pf_obs = obs.data(), in.pf_pts_i_z = zi.data();
  std::vector<double> r(NP * 2), ji(NP * 14), jj(NP * 14), je(NP * 14), jl(NP * 2);
  S_.assign((size_t)D * D, 0.0), g_.assign(D, 0.0);
  viml_linearize_out out{};
  out.pf_residual = r.data(), out.pf_jac_pose_i = ji.data(), out.pf_jac_pose_j = jj.data(), out.pf_jac_ex = je.data(), out.pf_jac_feat = jl.data();
  out.S = S_.data(), out.g = g_.data();
  last_error = viml_linearize_batch(ctx, &in, &out, VIML_OUT_RESIDUAL_JACOBIAN | VIML_OUT_SCHUR | (with_loss ? VIML_LOSS_CAUCHY : 0u));
  check_rc(last_error, "viml_linearize_batch");
  for (size_t k = 0; k < NP; ++k) {  // keep the public per-factor members valid (ThreadsConstructA's inputs, :17)
    ResidualBlockInfo* it = pts[k];
    it->residuals.assign(&r[2 * k], &r[2 * k] + 2);
    it->jacobians.resize(4);
    it->jacobians[0].assign(&ji[14 * k], &ji[14 * k] + 14);
    it->jacobians[1].assign(&jj[14 * k], &jj[14 * k] + 14);
    it->jacobians[2].assign(&je[14 * k], &je[14 * k] + 14);
    it->jacobians[3].assign(&jl[2 * k], &jl[2 * k] + 2);
    delete[] it->raw_jacobians;
    it->raw_jacobians = new double*[4];
    for (int b = 0; b < 4; ++b) it->raw_jacobians[b] = it->jacobians[b].data();
  }
}

void MarginalizationInfo::marginalize() {
  viml_ctx* ctx = need_ctx();
  // Block order (SURVEY.md §8a UB policy): marginalised blocks in first-appearance order with the landmarks
  // last among them, then the kept blocks in first-appearance order.  (The reference iterates an
  // unordered_map keyed by heap addresses, marginalization_factor.cpp:176-194.)
  std::unordered_map<long, bool> is_landmark;
  for (auto it : factors)
    if (dynamic_cast<ProjectionFactor*>(it->cost_function)) is_landmark[reinterpret_cast<long>(it->parameter_blocks[3])] = true;
  int pos = 0;
  for (long a : order_)
    if (parameter_block_idx.count(a) && !is_landmark.count(a)) parameter_block_idx[a] = pos, pos += localSize(parameter_block_size[a]);
  const int m_dense = pos;
  for (long a : order_)
    if (parameter_block_idx.count(a) && is_landmark.count(a)) parameter_block_idx[a] = pos, pos += 1;
  m = pos;
  for (long a : order_)
    if (!parameter_block_idx.count(a)) parameter_block_idx[a] = pos, pos += localSize(parameter_block_size[a]);
  n = pos - m;
  const int L = m - m_dense, pd = m_dense + n;  // dense system after the device eliminated the L landmarks
  auto didx = [&](long a) { const int i = parameter_block_idx[a]; return i < m_dense ? i : i - L; };
  std::vector<double> A((size_t)pd * pd, 0.0), b(pd, 0.0);
  // (1) + (2): the landmark-eliminated S, g of the point factors (device, preMarginalize) and the host-resident cost functions
  // (IMU, prior) as evaluated dense blocks, accumulated on the device by ThreadsConstructA's rule (:141-172) in ONE call of
  // viml_reduced_from_schur.  Device column space: [gpu poses | extrinsic | every other block in dense order].
  {
    const int P = S_.empty() ? 0 : (int)gpu_pose_blocks_.size();
    const int D = S_.empty() ? 0 : 6 * (P + 1);
    std::vector<int> dev_col(pd, -1);   // dense index -> device column
    if (D) {
      for (int p = 0; p < P; ++p)
        for (int r = 0; r < 6; ++r) dev_col[didx(reinterpret_cast<long>(gpu_pose_blocks_[p])) + r] = 6 * p + r;
      for (int r = 0; r < 6; ++r) dev_col[didx(reinterpret_cast<long>(gpu_ex_block_)) + r] = 6 * P + r;
    }
    int X = 0;
    for (int k = 0; k < pd; ++k)
      if (dev_col[k] < 0) dev_col[k] = D + X++;
    std::vector<int32_t> woff(2, 0), col_index;
    std::vector<int64_t> roff(1, 0), coff(1, 0), joff(1, 0);
    std::vector<double> res, jac;
    for (auto it : factors) {
      if (dynamic_cast<ProjectionFactor*>(it->cost_function)) continue;
      const std::vector<int32_t>& sizes = it->cost_function->parameter_block_sizes();
      const int nres = it->cost_function->num_residuals();
      std::vector<int> cols, src_block, src_col;
      for (size_t i = 0; i < it->parameter_blocks.size(); i++) {
        const long ai = reinterpret_cast<long>(it->parameter_blocks[i]);
        if (is_landmark.count(ai)) die("MarginalizationInfo: a landmark block is shared with a non-projection factor");
        for (int c = 0; c < localSize(sizes[i]); ++c) cols.push_back(dev_col[didx(ai) + c]), src_block.push_back((int)i), src_col.push_back(c);
      }
      const int nc = (int)cols.size();
      for (int q = 0; q < nres; ++q) {
        res.push_back(it->residuals[q]);
        for (int c = 0; c < nc; ++c) jac.push_back(it->jacobians[src_block[c]][(size_t)q * sizes[src_block[c]] + src_col[c]]);
      }
      col_index.insert(col_index.end(), cols.begin(), cols.end());
      roff.push_back(roff.back() + nres), coff.push_back(coff.back() + nc), joff.push_back(joff.back() + (int64_t)nres * nc);
      ++woff[1];
    }
    const int Dx = D + X;
    std::vector<double> S0((size_t)std::max(D, 1) * std::max(D, 1), 0.0), g0(std::max(D, 1), 0.0), Sx((size_t)Dx * Dx), gx(Dx);
    viml_dense_factors dn{};
    dn.extra_dim = D ? X : X - 1;   // D == 0: one dummy column keeps the C-ABI's D >= 1
    dn.n_factors = woff[1];
    dn.window_offset = woff.data(), dn.row_offset = roff.data(), dn.col_offset = coff.data(), dn.jac_offset = joff.data();
    dn.col_index = col_index.data(), dn.residual = res.data(), dn.jacobian = jac.data();
    viml_reduced_out ro{};
    ro.Sx = Sx.data(), ro.gx = gx.data();
    last_error = viml_reduced_from_schur(ctx, 1, D ? D : 1, D ? S_.data() : S0.data(), D ? g_.data() : g0.data(), &dn, &ro, 0);
    check_rc(last_error, "viml_reduced_from_schur");
    for (int r = 0; r < pd; ++r) {
      for (int c = 0; c < pd; ++c) A[(size_t)r * pd + c] = Sx[(size_t)dev_col[r] * Dx + dev_col[c]];
      b[r] = gx[dev_col[r]];
    }
  }
  // (3) dense elimination of the remaining marginalised blocks + square-root factorisation on the device (:264-293)
  A_schur.resize(n, n);
  b_schur.resize(n);
  linearized_jacobians.resize(n, n);
  linearized_residuals.resize(n);
  viml_marg_batch in{};
  in.n_problems = 1, in.pos = pd, in.m = m_dense, in.eps = eps, in.A = A.data(), in.b = b.data();
  viml_marg_out out{};
#ifdef VIML_WITH_EIGEN
  std::vector<double> As((size_t)n * n), Jl((size_t)n * n);
  out.A_schur = As.data(), out.linearized_jacobians = Jl.data();
#else
  out.A_schur = A_schur.data(), out.linearized_jacobians = linearized_jacobians.data();
#endif
  out.b_schur = b_schur.data(), out.linearized_residuals = linearized_residuals.data();
  last_error = viml_marginalize_batch(ctx, &in, &out, 0);
  check_rc(last_error, "viml_marginalize_batch");
#ifdef VIML_WITH_EIGEN
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) A_schur(r, c) = As[(size_t)r * n + c], linearized_jacobians(r, c) = Jl[(size_t)r * n + c];
#endif
}

std::vector<double*> MarginalizationInfo::getParameterBlocks(std::unordered_map<long, double*>& addr_shift) {
  std::vector<double*> keep_block_addr;
  keep_block_size.clear();
  keep_block_idx.clear();
  keep_block_data.clear();
  for (long a : order_) {  // :308-317, in the shim's deterministic order
    if (parameter_block_idx[a] >= m) {
      keep_block_size.push_back(parameter_block_size[a]);
      keep_block_idx.push_back(parameter_block_idx[a]);
      keep_block_data.push_back(parameter_block_data[a]);
      keep_block_addr.push_back(addr_shift[a]);
    }
  }
  sum_block_size = std::accumulate(std::begin(keep_block_size), std::end(keep_block_size), 0);
  return keep_block_addr;
}

MarginalizationFactor::MarginalizationFactor(MarginalizationInfo* _marginalization_info) : marginalization_info(_marginalization_info) {
  for (auto it : marginalization_info->keep_block_size) mutable_parameter_block_sizes()->push_back(it);
  set_num_residuals(marginalization_info->n);
}

// marginalization_factor.cpp:335-384 — stays on the host: n <= 76, a dense GEMV.
bool MarginalizationFactor::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
  const int n = marginalization_info->n, m = marginalization_info->m;
  std::vector<double> dx(n, 0.0);
  for (int i = 0; i < static_cast<int>(marginalization_info->keep_block_size.size()); i++) {
    const int size = marginalization_info->keep_block_size[i], idx = marginalization_info->keep_block_idx[i] - m;
    const double* x = parameters[i];
    const double* x0 = marginalization_info->keep_block_data[i];
    if (size != 7) {
      for (int k = 0; k < size; ++k) dx[idx + k] = x[k] - x0[k];
    } else {
      for (int k = 0; k < 3; ++k) dx[idx + k] = x[k] - x0[k];
      // q0.inverse() * q  (Eigen: conjugate / squaredNorm; Utility::positify is the identity, utility.h:41-48)
      const double n2 = x0[3] * x0[3] + x0[4] * x0[4] + x0[5] * x0[5] + x0[6] * x0[6];
      const double aw = x0[6] / n2, ax = -x0[3] / n2, ay = -x0[4] / n2, az = -x0[5] / n2;
      const double bw = x[6], bx = x[3], by = x[4], bz = x[5];
      const double w = aw * bw - ax * bx - ay * by - az * bz;
      const double vx = aw * bx + ax * bw + ay * bz - az * by;
      const double vy = aw * by + ay * bw + az * bx - ax * bz;
      const double vz = aw * bz + az * bw + ax * by - ay * bx;
      const double sgn = (w >= 0) ? 1.0 : -1.0;  // :361-364
      dx[idx + 3] = 2.0 * sgn * vx, dx[idx + 4] = 2.0 * sgn * vy, dx[idx + 5] = 2.0 * sgn * vz;
    }
  }
  for (int r = 0; r < n; ++r) {
    double s = 0.0;
    for (int c = 0; c < n; ++c) s += marginalization_info->linearized_jacobians(r, c) * dx[c];
    residuals[r] = marginalization_info->linearized_residuals(r) + s;
  }
  if (jacobians)
    for (int i = 0; i < static_cast<int>(marginalization_info->keep_block_size.size()); i++)
      if (jacobians[i]) {
        const int size = marginalization_info->keep_block_size[i], local_size = marginalization_info->localSize(size);
        const int idx = marginalization_info->keep_block_idx[i] - m;
        for (int r = 0; r < n; ++r)
          for (int c = 0; c < size; ++c)
            jacobians[i][(size_t)r * size + c] = c < local_size ? marginalization_info->linearized_jacobians(r, idx + c) : 0.0;
      }
  return true;
}

// ---------------------------------------------------------------------------------------------- association
void pack_point_channels(int n, const PointChannels& c, int num_of_cam, int32_t* feature_id, int32_t* camera_id, double* out) {
  for (int j = 0; j < n; ++j) {
    const int v = (int)(c.id[j] + 0.5);                       // estimator_node.cpp:386
    feature_id[j] = v / num_of_cam, camera_id[j] = v % num_of_cam;
    double* o = out + 7 * (size_t)j;
    o[0] = c.x[j], o[1] = c.y[j], o[2] = c.z[j], o[3] = c.u[j], o[4] = c.v[j], o[5] = c.vx[j], o[6] = c.vy[j];
    if (o[2] != 1.0) die("pack_point_channels: z != 1 (ROS_ASSERT(z == 1), estimator_node.cpp:397)");
  }
}

void pack_line_channels(int n, const LineChannels& c, int32_t* feature_id, double* lines2d) {
  for (int j = 0; j < n; ++j) {
    feature_id[j] = (int)(c.id[j] / 1);                       // estimator_node.cpp:405
    double* o = lines2d + 4 * (size_t)j;
    o[0] = c.sx[j], o[1] = c.sy[j], o[2] = c.ex[j], o[3] = c.ey[j];
  }
}

LineMapAssociator::LineMapAssociator(const std::vector<viml::Vector6d>& lines3d_map) : map_(lines3d_map) {
  std::memset(have_, 0, sizeof(have_));
  std::memset(cull_pose_, 0, sizeof(cull_pose_));
  std::memset(cull_ex_, 0, sizeof(cull_ex_));
  static_assert(sizeof(viml::Vector6d) == 6 * sizeof(double), "map rows must be packed");
  check_rc(viml_set_map(need_ctx(), map_.empty() ? nullptr : map_[0].data(), (int64_t)map_.size()), "viml_set_map");
}

int LineMapAssociator::UpdateLinesInFoV(int i, const double* para_Pose_i, const double* para_Ex_Pose) {
  viml_ctx* ctx = need_ctx();
  std::memcpy(cull_pose_[i], para_Pose_i, 56);
  std::memcpy(cull_ex_[i], para_Ex_Pose, 56);
  have_[i] = true;
  // the list stays on the device in window slot i (viml_fov_update); the public WorldLinesInFOV[i] mirror is read back once
  int32_t count = 0;
  check_rc(viml_fov_update(ctx, i, cull_pose_[i], cull_ex_[i], &count), "viml_fov_update");
  std::vector<int32_t> list((size_t)count + 1);
  const int32_t slot = i;
  viml_assoc_query q{};
  q.n_poses = 1, q.lines_per_pose = 0, q.match_poses = cull_pose_[i], q.ex_pose = cull_ex_[i], q.fov_slot = &slot;
  int32_t c2 = 0;
  viml_assoc_out o{};
  o.fov_count = &c2, o.fov_index = list.data(), o.fov_capacity = (int32_t)list.size();
  check_rc(viml_line_associate(ctx, &q, &o, VIML_FOV_CACHED), "viml_line_associate");
  WorldLinesInFOV[i].assign(list.begin(), list.begin() + count);
  return count;
}

int LineMapAssociator::updateLinePairInWindow(const double (*para_Pose)[7], const double* para_Ex_Pose,
                                              const std::vector<Observation>& obs, std::vector<Match>* out) {
  viml_ctx* ctx = need_ctx();
  const int NPOSE = kWindowSize + 1;
  std::vector<std::vector<int>> per(NPOSE);
  for (size_t k = 0; k < obs.size(); ++k) {
    if (obs[k].frame < 0 || obs[k].frame >= NPOSE || !have_[obs[k].frame]) die("updateLinePairInWindow: frame without a FoV list");
    per[obs[k].frame].push_back((int)k);
  }
  int L = 1;
  for (auto& v : per) L = std::max<int>(L, (int)v.size());
  std::vector<double> match((size_t)NPOSE * 7), ex((size_t)NPOSE * 7), l2d((size_t)NPOSE * L * 4, 0.0);
  std::vector<int32_t> nl(NPOSE);
  for (int f = 0; f < NPOSE; ++f) {
    std::memcpy(&match[7 * f], para_Pose[f], 56);
    std::memcpy(&ex[7 * f], para_Ex_Pose, 56);
    nl[f] = (int32_t)per[f].size();
    for (size_t s = 0; s < per[f].size(); ++s) std::memcpy(&l2d[((size_t)f * L + s) * 4], obs[per[f][s]].line, 32);
  }
  std::vector<int32_t> mi((size_t)NPOSE * L, -1);
  std::vector<float> err((size_t)NPOSE * L * 3, -1.f);
  std::vector<double> proj((size_t)NPOSE * L * 4, 0.0);
  viml_assoc_query q{};
  // against the FoV lists cached on the device when the frames entered (no re-cull): slot f <-> pose f
  q.n_poses = NPOSE, q.lines_per_pose = L, q.match_poses = match.data(), q.ex_pose = ex.data();
  q.lines2d = l2d.data(), q.n_lines2d = nl.data();
  viml_assoc_out o{};
  o.match_index = mi.data(), o.err = err.data(), o.projected = proj.data();
  check_rc(viml_line_associate(ctx, &q, &o, VIML_FOV_CACHED), "viml_line_associate");
  out->assign(obs.size(), Match());
  int matched = 0;
  for (int f = 0; f < NPOSE; ++f)
    for (size_t s = 0; s < per[f].size(); ++s) {
      const size_t src = (size_t)f * L + s;
      Match& mt = (*out)[per[f][s]];
      mt.errA = err[3 * src], mt.errD = err[3 * src + 1], mt.overlap = err[3 * src + 2];
      mt.map_index = mi[src];
      std::memcpy(mt.projectedLine, &proj[4 * src], 32);
      mt.use_flag = true;                      // estimator.cpp:468
      mt.credible_line = !(mt.errA == -1);     // estimator.cpp:470-477
      matched += mt.map_index >= 0;
    }
  return matched;
}

LineMapAssociator::Match LineMapAssociator::LineCorrespondenceInFrame(int frame_index, const double detect_line[4],
                                                                        const double (*para_Pose)[7], const double* para_Ex_Pose) {
  viml_ctx* ctx = need_ctx();
  if (!have_[frame_index]) die("LineCorrespondenceInFrame: frame without a FoV list");
  viml_assoc_query q{};
  const int32_t slot = frame_index;
  q.n_poses = 1, q.lines_per_pose = 1, q.fov_slot = &slot;
  q.match_poses = para_Pose[frame_index], q.ex_pose = para_Ex_Pose, q.lines2d = detect_line;
  Match mt;
  int32_t mi = -1;
  float err[3] = {-1, -1, -1};
  viml_assoc_out o{};
  o.match_index = &mi, o.err = err, o.projected = mt.projectedLine;
  check_rc(viml_line_associate(ctx, &q, &o, VIML_FOV_CACHED), "viml_line_associate");
  mt.errA = err[0], mt.errD = err[1], mt.overlap = err[2], mt.map_index = mi;
  mt.use_flag = true, mt.credible_line = !(mt.errA == -1);
  return mt;
}

void LineMapAssociator::slideWindowOld() {  // estimator.cpp:2131-2160: slot i <- slot i+1
  check_rc(viml_fov_slide(need_ctx(), 1), "viml_fov_slide");
  for (int i = 0; i < kWindowSize; ++i) {
    WorldLinesInFOV[i].swap(WorldLinesInFOV[i + 1]);
    std::memcpy(cull_pose_[i], cull_pose_[i + 1], 56);
    std::memcpy(cull_ex_[i], cull_ex_[i + 1], 56);
    have_[i] = have_[i + 1];
  }
  WorldLinesInFOV[kWindowSize] = WorldLinesInFOV[kWindowSize - 1];  // :2160 keeps a copy in the newest slot
  std::memcpy(cull_pose_[kWindowSize], cull_pose_[kWindowSize - 1], 56);
  std::memcpy(cull_ex_[kWindowSize], cull_ex_[kWindowSize - 1], 56);
  have_[kWindowSize] = have_[kWindowSize - 1];
}
void LineMapAssociator::slideWindowNew() {  // estimator.cpp:2218: second-newest <- newest
  check_rc(viml_fov_slide(need_ctx(), 0), "viml_fov_slide");
  WorldLinesInFOV[kWindowSize - 1] = WorldLinesInFOV[kWindowSize];
  std::memcpy(cull_pose_[kWindowSize - 1], cull_pose_[kWindowSize], 56);
  std::memcpy(cull_ex_[kWindowSize - 1], cull_ex_[kWindowSize], 56);
  have_[kWindowSize - 1] = have_[kWindowSize];
}

void LineMapAssociator::removeLineOutliers(const std::vector<int32_t>& track_offset, const std::vector<int32_t>& line_index,
                                           std::vector<uint8_t>* credible_line, std::vector<uint8_t>* credible_matching) {
  // feature_manager.cpp:494-541 for every line track in one device call (viml_track_gate)
  const int T = (int)track_offset.size() - 1;
  credible_line->assign(line_index.size(), 1);
  credible_matching->assign(T > 0 ? T : 0, 1);
  if (T <= 0) return;
  check_rc(viml_track_gate(need_ctx(), T, track_offset.data(), line_index.data(), credible_line->data(), credible_matching->data(), 0),
           "viml_track_gate");
}

bool LineMapAssociator::removeLineOutlier(const std::vector<viml::Vector3d>& line_vec, std::vector<bool>* credible_line) {
  // feature_manager.cpp:494-541 for ONE track on the host (<= 11 observations); removeLineOutliers is the batched device path.
  const int obvers_time = (int)line_vec.size();
  credible_line->assign(obvers_time, true);
  if (obvers_time < 1) return true;
  int count = 0;
  for (int k = 0; k < obvers_time; ++k) {
    const double dx = line_vec[0][0] - line_vec[k][0], dy = line_vec[0][1] - line_vec[k][1], dz = line_vec[0][2] - line_vec[k][2];
    const float diff_ = (float)std::sqrt((dx * dx + dy * dy) + dz * dz);
    if (diff_ > 0.1) {
      count++;
      (*credible_line)[k] = false;
    }
  }
  return !((count / obvers_time) >= 0.5);  // integer division (:524)
}
