// viml_host.h — host-side mirror of the reference's hot-path classes, implemented over the C-ABI (viml.h).
//
// Same class names, member names, argument meaning and return conventions as the reference so that
// Estimator::OptimizationWithLine (estimator.cpp:1677-2119) and processImagewithLine (:342-346) compile
// against it unchanged (with -DVIML_WITH_EIGEN -DVIML_WITH_CERES the small stand-in types below become the
// real Eigen/Ceres types):
//   ProjectionFactor          factor/projection_factor.h:10-21
//   LineProjectionFactor      factor/line_projection_factor.h:13-34
//   ResidualBlockInfo         factor/marginalization_factor.h:15-35
//   MarginalizationInfo       factor/marginalization_factor.h:47-72
//   MarginalizationFactor     factor/marginalization_factor.h:74-81
//   LineMapAssociator         the UpdateLinesInFoV / updateLinePairInWindow / LineCorrespondenceInFrame /
//                             removeLineOutlier group of Estimator + FeatureManager (estimator.cpp:385-481,
//                             :671-885, feature_manager.cpp:494-541)
// There is no CPU implementation of the hot path in here: every Evaluate / marginalize / association call
// ends in viml_linearize_batch / viml_marginalize_batch / viml_line_associate on the GPU.  What stays on the
// host is what the reference keeps there: cost functions of other types (IMU, prior) are
// evaluated through their own virtual Evaluate, and MarginalizationFactor::Evaluate is a small dense GEMV.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "ceres_compat.h"
#include "viml.h"

#ifdef VIML_WITH_EIGEN
#include <Eigen/Dense>
#endif

namespace viml {

#ifdef VIML_WITH_EIGEN
using Vector3d = Eigen::Vector3d;
using Matrix2d = Eigen::Matrix2d;
using Matrix3d = Eigen::Matrix3d;
using MatrixXd = Eigen::MatrixXd;
using VectorXd = Eigen::VectorXd;
#else
// Minimal stand-ins with the accessors the reference code uses on these members.
struct Vector3d {
  double v[3] = {0, 0, 0};
  Vector3d() {}
  Vector3d(double a, double b, double c) : v{a, b, c} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  const double* data() const { return v; }
};
struct Matrix2d {
  double m[4] = {0, 0, 0, 0};
  double& operator()(int r, int c) { return m[2 * r + c]; }
  double operator()(int r, int c) const { return m[2 * r + c]; }
  static Matrix2d Identity() {
    Matrix2d I;
    I.m[0] = I.m[3] = 1;
    return I;
  }
};
inline Matrix2d operator*(double s, const Matrix2d& A) {
  Matrix2d B;
  for (int k = 0; k < 4; ++k) B.m[k] = s * A.m[k];
  return B;
}
struct Matrix3d {
  double m[9] = {0};
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
};
struct MatrixXd {  // row-major dynamic matrix
  int r = 0, c = 0;
  std::vector<double> d;
  void resize(int rr, int cc) { r = rr, c = cc, d.assign((size_t)rr * cc, 0.0); }
  int rows() const { return r; }
  int cols() const { return c; }
  double& operator()(int i, int j) { return d[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
  double* data() { return d.data(); }
  const double* data() const { return d.data(); }
};
struct VectorXd {
  std::vector<double> d;
  void resize(int n) { d.assign(n, 0.0); }
  int size() const { return (int)d.size(); }
  double& operator()(int i) { return d[i]; }
  double operator()(int i) const { return d[i]; }
  double* data() { return d.data(); }
  const double* data() const { return d.data(); }
};
#endif
using Vector6d = std::array<double, 6>;

// Process-wide handle to the CUDA library (one context, one device, one calling thread: SURVEY.md §8b threading).
class Runtime {
 public:
  static Runtime& instance();
  // Must be called once before any factor is evaluated (Estimator::setParameters is the natural place).
  // Returns a viml error code; there is no CPU fallback when this fails.
  int configure(const viml_config& cfg, int device = 0);
  viml_ctx* ctx() const { return ctx_; }
  const viml_config& config() const { return cfg_; }
  void shutdown();
  ~Runtime();

 private:
  viml_ctx* ctx_ = nullptr;
  viml_config cfg_{};
};

}  // namespace viml

class LinearizationBatch;

class ProjectionFactor : public ceres::SizedCostFunction<2, 7, 7, 7, 1> {
 public:
  ProjectionFactor(const viml::Vector3d& _pts_i, const viml::Vector3d& _pts_j);
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const;
  // projection_factor.cpp:126-228: forward-difference check; returns max |analytic - numeric| / max |analytic|.
  double check(double** parameters);

  viml::Vector3d pts_i, pts_j;
  static viml::Matrix2d sqrt_info;
  static double sum_t;

 private:
  friend class LinearizationBatch;
  LinearizationBatch* batch_ = nullptr;
  int slot_ = -1;
};

class LineProjectionFactor : public ceres::SizedCostFunction<2, 7> {
 public:
  LineProjectionFactor(const viml::Vector3d& _pts_start, const viml::Vector3d& _pts_end, const viml::Vector3d& _line_param,
                       const viml::Matrix3d _K, const viml::Matrix3d _b_c_R, const viml::Vector3d _b_c_T);
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const;

  viml::Vector3d pts_start, pts_end, line_param;
  viml::Matrix3d K, b_c_R;
  viml::Vector3d b_c_T;
  static double sum_t;

 private:
  friend class LinearizationBatch;
  LinearizationBatch* batch_ = nullptr;
  int slot_ = -1;
};

// One GPU launch per evaluation point instead of one per residual block: register it as
// Solver::Options::evaluation_callback (Ceres 1.14) / Problem::Options::evaluation_callback (2.x).
// PrepareForEvaluation packs the CURRENT values of the registered parameter blocks, runs
// viml_linearize_batch (mode A) into pinned buffers; each factor's Evaluate then copies its slice.
class LinearizationBatch : public ceres::EvaluationCallback {
 public:
  ~LinearizationBatch();
  // same argument order as problem.AddResidualBlock(f, loss, para_Pose[i], para_Pose[j], para_Ex_Pose[0], para_Feature[k])
  void AddResidualBlock(ProjectionFactor* f, double* pose_i, double* pose_j, double* ex_pose, double* feature);
  void AddResidualBlock(LineProjectionFactor* f, double* pose);
  void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) override;
  // Call when the solve has returned: later Evaluate calls (e.g. ResidualBlockInfo::Evaluate at other states) must not be
  // served from the last evaluation point by accident.  (fetch also compares the parameter VALUES, so this is belt and braces.)
  void Invalidate() { valid_ = false; }
  int num_point_factors() const { return (int)pf_.size(); }
  int num_line_factors() const { return (int)lf_.size(); }
  int last_error() const { return last_rc_; }
  long served() const { return served_; }   // Evaluate calls answered from the batch so far (tests, diagnostics)

 private:
  friend class ProjectionFactor;
  friend class LineProjectionFactor;
  struct PF { ProjectionFactor* f; double *pi, *pj, *ex, *feat; int i, j, l; };
  struct LF { LineProjectionFactor* f; double* pose; int frame; };
  int pose_index(double* p);
  bool fetch(const ProjectionFactor* f, double const* const* parameters, double* residuals, double** jacobians) const;
  bool fetch(const LineProjectionFactor* f, double const* const* parameters, double* residuals, double** jacobians) const;
  void free_pinned();
  std::vector<PF> pf_;
  std::vector<LF> lf_;
  std::vector<double*> poses_, feats_;
  std::unordered_map<double*, int> pose_id_, feat_id_;
  double* ex_ = nullptr;
  // values of the registered blocks at the last PrepareForEvaluation: during ceres::Solve the pointers handed to
  // Evaluate are the minimizer's own state vector (ResidualBlock::Evaluate passes parameter_block->state()), not the
  // user's arrays, so a factor is served by slot when the VALUES it is asked at equal the evaluation point
  std::vector<double> snap_poses_, snap_lam_, snap_ex_;
  bool valid_ = false;
  mutable long served_ = 0;
  int last_rc_ = 0;
  // pinned result buffers
  double *r_pf_ = nullptr, *ji_ = nullptr, *jj_ = nullptr, *je_ = nullptr, *jl_ = nullptr, *r_lf_ = nullptr, *jp_ = nullptr;
  size_t cap_pf_ = 0, cap_lf_ = 0;
};

struct ResidualBlockInfo {
  ResidualBlockInfo(ceres::CostFunction* _cost_function, ceres::LossFunction* _loss_function,
                    std::vector<double*> _parameter_blocks, std::vector<int> _drop_set)
      : cost_function(_cost_function), loss_function(_loss_function), parameter_blocks(_parameter_blocks), drop_set(_drop_set) {}
  void Evaluate();  // marginalization_factor.cpp:3-69 (one factor; MarginalizationInfo batches the projection factors)

  ceres::CostFunction* cost_function;
  ceres::LossFunction* loss_function;
  std::vector<double*> parameter_blocks;
  std::vector<int> drop_set;
  double** raw_jacobians = nullptr;
  std::vector<std::vector<double>> jacobians;  // row-major num_residuals x block_size, like the reference
  std::vector<double> residuals;
  int localSize(int size) { return size == 7 ? 6 : size; }
};

class MarginalizationInfo {
 public:
  ~MarginalizationInfo();
  int localSize(int size) const;
  int globalSize(int size) const;
  void addResidualBlockInfo(ResidualBlockInfo* residual_block_info);
  void preMarginalize();
  void marginalize();
  std::vector<double*> getParameterBlocks(std::unordered_map<long, double*>& addr_shift);

  std::vector<ResidualBlockInfo*> factors;
  int m = 0, n = 0;
  std::unordered_map<long, int> parameter_block_size;  // global size
  int sum_block_size = 0;
  std::unordered_map<long, int> parameter_block_idx;   // local size
  std::unordered_map<long, double*> parameter_block_data;
  std::vector<int> keep_block_size;  // global size
  std::vector<int> keep_block_idx;   // local size
  std::vector<double*> keep_block_data;
  viml::MatrixXd linearized_jacobians;
  viml::VectorXd linearized_residuals;
  const double eps = 1e-8;
  // Reduced system before the final factorisation (A x = b on the kept blocks), exposed for tests.
  viml::MatrixXd A_schur;
  viml::VectorXd b_schur;
  int last_error = 0;

 private:
  std::vector<long> order_;  // documented deviation: first-appearance order instead of unordered_map order
  // landmark-eliminated point-factor system from the GPU (filled by preMarginalize)
  std::vector<double> S_, g_;
  std::vector<double*> gpu_pose_blocks_;
  double* gpu_ex_block_ = nullptr;
  int gpu_landmarks_ = 0;
};

class MarginalizationFactor : public ceres::CostFunction {
 public:
  MarginalizationFactor(MarginalizationInfo* _marginalization_info);
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const;
  MarginalizationInfo* marginalization_info;
};

// Ingest of the front-end messages (estimator_node.cpp:382-413): the sensor_msgs::PointCloud of the point tracker and of the line
// tracker arrive as float32 channel arrays; they are unpacked straight into the SoA layout the C-ABI takes (double precision, the
// 2D line endpoints as viml_assoc_query::lines2d rows), without the per-feature std::map / Eigen temporaries of the reference.
struct PointChannels {   // img_msg: points[j] = (x, y, z), channels[0] = id * NUM_OF_CAM + camera, [1..2] = pixel u v, [3..4] = velocity
  const float *x, *y, *z, *id, *u, *v, *vx, *vy;
};
struct LineChannels {    // line_msg: channels[0] = id, channels[1..4] = start x y, end x y (raw pixels)
  const float *id, *sx, *sy, *ex, *ey;
};
// feature_id[j], camera_id[j], xyz_uv_velocity[j][7]  (estimator_node.cpp:384-399; z must be 1, :397)
void pack_point_channels(int n, const PointChannels& c, int num_of_cam, int32_t* feature_id, int32_t* camera_id, double* xyz_uv_velocity);
// feature_id[j], lines2d[j][4]  (estimator_node.cpp:403-413)
void pack_line_channels(int n, const LineChannels& c, int32_t* feature_id, double* lines2d);

// The line-association group of Estimator / FeatureManager.
class LineMapAssociator {
 public:
  static const int kWindowSize = 10;  // WINDOW_SIZE (parameters.h:20)
  struct Match {                      // what updateLinePairInWindow stores per observation (estimator.cpp:463-477)
    float errA = -1, errD = -1, overlap = -1;
    int map_index = -1;               // index into lines3d_map of lineWorld; -1 when unmatched
    double projectedLine[4] = {0, 0, 0, 0};
    bool use_flag = false, credible_line = false;
  };
  struct Observation { int frame; double line[4]; };  // one lineFeaturePerFrame: frame = start_frame + k, raw-pixel endpoints

  explicit LineMapAssociator(const std::vector<viml::Vector6d>& lines3d_map);  // Estimator::setParameters (estimator.cpp:54-58)
  // estimator.cpp:385-447; the list is computed once, when the frame enters, and stays ON THE DEVICE in window slot i
  // (viml_fov_update); LineCorrespondenceInFrame / updateLinePairInWindow match against it under the pose they are given later.
  int UpdateLinesInFoV(int i, const double* para_Pose_i, const double* para_Ex_Pose);
  // estimator.cpp:449-481: re-match every observation of every track with the current poses (one GPU call).
  int updateLinePairInWindow(const double (*para_Pose)[7], const double* para_Ex_Pose,
                             const std::vector<Observation>& obs, std::vector<Match>* out);
  // estimator.cpp:671-885 for a single query.
  Match LineCorrespondenceInFrame(int frame_index, const double detect_line[4], const double (*para_Pose)[7],
                                  const double* para_Ex_Pose);
  // window shifting of the cached lists (estimator.cpp:2148, :2160, :2218)
  void slideWindowOld();
  void slideWindowNew();
  // feature_manager.cpp:494-541 on one track: line_vec[k] = PtrEnd-PtrStart of the matched map line of observation k.
  static bool removeLineOutlier(const std::vector<viml::Vector3d>& line_vec, std::vector<bool>* credible_line);
  // the same gate for every track in one device call: track t owns observations track_offset[t] .. track_offset[t+1]-1,
  // line_index[k] = map index of observation k's lineWorld (-1 = the fake line of estimator.cpp:709)
  static void removeLineOutliers(const std::vector<int32_t>& track_offset, const std::vector<int32_t>& line_index,
                                 std::vector<uint8_t>* credible_line, std::vector<uint8_t>* credible_matching);

  std::vector<int> WorldLinesInFOV[kWindowSize + 1];
  const std::vector<viml::Vector6d>& map() const { return map_; }

 private:
  std::vector<viml::Vector6d> map_;
  double cull_pose_[kWindowSize + 1][7];
  double cull_ex_[kWindowSize + 1][7];
  bool have_[kWindowSize + 1];
};
