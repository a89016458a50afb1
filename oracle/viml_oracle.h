/*
 * viml_oracle.h — CPU ORACLE for the TC-VIML linearisation hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call
 * this.  The product (tc-viml_b200/) never links, loads or calls anything under oracle/.
 *
 * PARITY STATUS: "parity unpinned" in the strict sense of the build contract.  The reference ships no
 * tests, golden vectors or known-answer fixtures and cannot be compiled here (it needs
 * Eigen3, Ceres, ROS and OpenCV C++, none of which exist in this image; probed: no Eigen/ or ceres.h on
 * disk).  What pins this restatement instead:
 *   (1) it follows the reference sources statement by statement (file:line cited at each function);
 *   (2) oracle_mp.py re-derives residuals/Jacobians/H/b/Schur independently in 50-digit mpmath;
 *   (3) the reference's own self-check convention, ProjectionFactor::check (projection_factor.cpp:126-228:
 *       forward differences, Q <- Q*deltaQ(d)), is applied to the restated ProjectionFactor;
 *   (4) the association is run on the reference's real fixtures (line_3d.txt, sensor.yaml, GT data.csv)
 *       and the results are committed as golden vectors under tests/golden/.
 * Third-party arithmetic that is not in /root/reference (Eigen 3.3.x quaternion / small dense ops,
 * Ceres 1.14-2.1 CauchyLoss + corrector, SelfAdjointEigenSolver) is restated from the published
 * algorithms.  Summation order contract: fixed-size dot products are
 * ((a0*b0 + a1*b1) + a2*b2), A*B*C is (A*B)*C; compiled with -ffp-contract=off so no FMA is formed.
 */
#ifndef VIML_ORACLE_H_
#define VIML_ORACLE_H_

#include "../include/viml.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ProjectionFactor::Evaluate (projection_factor.cpp:21-124).  parameters = {pose_i[7], pose_j[7],
 * ex[7], inv_dep[1]}; jacobians NULL or 4 pointers, each nullable; layouts as the reference. */
int orc_projection_evaluate(const double* pts_i, const double* pts_j, double sqrt_info,
                            double const* const* parameters, double* residuals, double** jacobians);

/* LineProjectionFactor::Evaluate (line_projection_factor.cpp:19-120).  K, b_c_R row-major 3x3. */
int orc_line_evaluate(const double* pts_start, const double* pts_end, const double* line_param,
                      const double* K, const double* b_c_R, const double* b_c_T,
                      double const* const* parameters, double* residuals, double** jacobians);

/* ceres::CauchyLoss(a)::Evaluate. */
void orc_cauchy_loss(double a, double sq_norm, double rho[3]);

/* Loss correction of ResidualBlockInfo::Evaluate (marginalization_factor.cpp:37-68) applied in place to
 * one residual (nres) and its nblk Jacobian blocks (row-major nres x sizes[k]). */
void orc_loss_correct(double cauchy_a, int nres, double* residuals, int nblk, const int* sizes,
                      double** jacobians);

/* Whole-batch restatement with the same in/out structs as viml_linearize_batch (HOST pointers).
 * Mode B follows ThreadsConstructA (marginalization_factor.cpp:141-172) on a dense A per window in
 * the canonical order [pose 0..P-1 | ex | landmark 0..F-1], 4 round-robin partial sums added in join
 * order 3,2,1,0 (marginalization_factor.cpp:232-261); blocks are then cut out of the dense A.
 * VIML_OUT_SCHUR follows the header's definition.  nthreads parallelises over windows only. */
int orc_linearize_batch(const viml_config* cfg, const viml_window_batch* in,
                        const viml_linearize_out* out, uint32_t flags, int nthreads);

/* Dense A/b of one window in canonical order (pos = 6*(P+1)+F), for tests. */
int orc_window_dense(const viml_config* cfg, const viml_window_batch* in, int w, uint32_t flags,
                     double* A, double* b);

/* MarginalizationInfo::marginalize numeric core (marginalization_factor.cpp:264-293), literal:
 * dense eigen-decomposition (cyclic Jacobi standing in for SelfAdjointEigenSolver) of the whole Amm. */
int orc_marginalize_dense(const double* A, const double* b, int pos, int m, double eps,
                          double* A_schur, double* b_schur, double* lin_jac, double* lin_res);

/* Symmetric eigen-decomposition used above: a [n][n] row-major (lower triangle read), eigenvalues
 * ascending in w, eigenvectors in the COLUMNS of v (row-major [n][n]). */
int orc_sym_eig(const double* a, int n, double* w, double* v);

/* MarginalizationFactor::Evaluate (marginalization_factor.cpp:335-384).  keep_* as produced by
 * getParameterBlocks; jacobians nullable like Ceres. */
int orc_marginalization_factor_evaluate(int n, int m, int nblk, const int* keep_block_size,
                                        const int* keep_block_idx, double const* const* keep_block_data,
                                        const double* lin_jac, const double* lin_res,
                                        double const* const* parameters, double* residuals,
                                        double** jacobians);

/* Estimator::UpdateLinesInFoV (estimator.cpp:385-447): returns the list length, writes map indices
 * in map order. Ric is derived from ex_pose (normalised quaternion) as initialLineFoVWindow does. */
int orc_update_lines_in_fov(const viml_config* cfg, const double* pose, const double* ex_pose,
                            const double* map_xyzxyz, int64_t n_lines, int32_t* out_index);

/* Estimator::LineCorrespondenceInFrame (estimator.cpp:671-885) against an explicit candidate list.
 * Returns the chosen MAP index or -1; err = (errA, errD, overlap) floats; projected[4]. */
int orc_line_correspondence(const viml_config* cfg, const double* pose, const double* ex_pose,
                            const double* map_xyzxyz, const int32_t* fov_index, int fov_count,
                            const double* line2d, float* err, double* projected);

/* Batch with the same structs as viml_line_associate (HOST pointers). */
int orc_line_associate(const viml_config* cfg, const double* map_xyzxyz, int64_t n_lines,
                       const viml_assoc_query* q, const viml_assoc_out* out, int nthreads);

/* FeatureManager::removeLineOutlier track gate (feature_manager.cpp:494-541) on one track of n_obs
 * observations; line_vec[k][3] = LineVec of the matched map line (oracle definition PtrEnd-PtrStart,
 * SURVEY.md §8a UB policy).  Writes credible_line[k]; returns credible_matching (0/1). */
/* Line2D::Line2D(Vector4d) (feature_manager.cpp:4-15): out = {A, B, C, A2B2, Length, Direction.x, Direction.y};
 * Line2D::Point2Flined (feature_manager.cpp:46-71): out = the returned point.  (Exposed for the check against oracle/_ref.) */
void orc_line2d(const double* seg, double* out);
void orc_point2flined(const double* seg, const double* p, double* out);
int orc_track_gate(int n_obs, const double* line_vec, uint8_t* credible_line);

/* c* = min{ c in [0,1] : acos(c) <= angle_th } for the host libm (SURVEY.md §7 "hard parts").
 * Exposed so tests can check the product's own threshold against the oracle's libm. */
double orc_cos_threshold(double angle_th);

int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
