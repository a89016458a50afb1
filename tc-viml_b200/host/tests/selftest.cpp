// selftest.cpp — tests of the host shim written the way the reference's own tests would read (it has none):
// build factors and a MarginalizationInfo exactly like Estimator::OptimizationWithLine does, and compare with the
// CPU oracle (oracle/viml_oracle.h).  `selftest --cpu` runs the parts that need no GPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "oracle/viml_oracle.h"
#include "viml_host.h"

static int g_fail = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);          \
      ++g_fail;                                                            \
    }                                                                      \
  } while (0)

static uint64_t g_seed = 0x5EED;
static double urand() {  // splitmix64 -> [0,1)
  uint64_t z = (g_seed += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) / 9007199254740992.0;
}
static double nrand() { return std::sqrt(-2.0 * std::log(urand() + 1e-300)) * std::cos(6.283185307179586 * urand()); }

static double rel_err(const double* a, const double* b, size_t n) {
  double scale = 1e-300, err = 0.0;
  for (size_t k = 0; k < n; ++k) scale = std::fmax(scale, std::fabs(b[k])), err = std::fmax(err, std::fabs(a[k] - b[k]));
  return err / scale;
}

static void random_pose(double* p, double tscale, double rscale) {
  for (int k = 0; k < 3; ++k) p[k] = tscale * nrand();
  double q[4] = {rscale * nrand(), rscale * nrand(), rscale * nrand(), 1.0};
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; ++k) p[3 + k] = q[k] / n * (1.0 + 1e-12 * nrand());
}

static viml_config euroc_config() {
  viml_config c{};
  c.fx = 461.6, c.fy = 460.3, c.cx = 363.0, c.cy = 248.1, c.width = 752, c.height = 480;
  const double R[9] = {0.958882, 0.283788, -0.00258614, -0.283713, 0.958774, 0.016038, 0.00703105, -0.0146448, 0.999868};
  const double T[3] = {-1.4494, -1.83337, -0.899281};
  std::memcpy(c.Rbw, R, sizeof(R));
  std::memcpy(c.Tbw, T, sizeof(T));
  c.overlap_th = 0.45, c.dist_th = 50, c.angle_th = 0.1745, c.sqrt_info = 460.0 / 1.5, c.cauchy_a = 1.0;
  return c;
}

// A linear cost function standing in for the host-resident factors (IMUFactor, the previous prior).
class DenseFactor : public ceres::CostFunction {
 public:
  DenseFactor(int nres, const std::vector<int>& sizes) {
    set_num_residuals(nres);
    for (int s : sizes) mutable_parameter_block_sizes()->push_back(s);
    r0.resize(nres);
    for (double& v : r0) v = nrand();
    for (int s : sizes) {
      std::vector<double> j((size_t)nres * s);
      for (int r = 0; r < nres; ++r)
        for (int c = 0; c < s; ++c) j[(size_t)r * s + c] = (s == 7 && c == 6) ? 0.0 : 3.0 * nrand();
      J.push_back(j);
    }
  }
  bool Evaluate(double const* const*, double* residuals, double** jacobians) const override {
    std::memcpy(residuals, r0.data(), sizeof(double) * r0.size());
    if (jacobians)
      for (size_t k = 0; k < J.size(); ++k)
        if (jacobians[k]) std::memcpy(jacobians[k], J[k].data(), sizeof(double) * J[k].size());
    return true;
  }
  std::vector<double> r0;
  std::vector<std::vector<double>> J;
};

static void test_ingest() {
  // estimator_node.cpp:382-413 channel unpacking
  const int n = 5;
  float x[n] = {0.1f, -0.2f, 0.3f, 0.0f, 0.5f}, y[n] = {0.2f, 0.1f, -0.4f, 0.3f, 0.0f}, z[n] = {1, 1, 1, 1, 1};
  float id[n] = {7, 8, 1000003, 12, 0}, u[n] = {100.5f, 200, 300, 400, 500}, v[n] = {50, 60, 70.25f, 80, 90};
  float vx[n] = {1, 2, 3, 4, 5}, vy[n] = {-1, -2, -3, -4, -5};
  int32_t fid[n], cam[n];
  double out[n][7];
  pack_point_channels(n, PointChannels{x, y, z, id, u, v, vx, vy}, 1, fid, cam, &out[0][0]);
  CHECK(fid[2] == (int)(1000003.0f + 0.5) && cam[2] == 0 && out[0][3] == (double)100.5f && out[2][4] == (double)70.25f && out[4][6] == -5.0);
  float lid[2] = {3, 41}, sx[2] = {10.5f, 20}, sy[2] = {30, 40}, ex[2] = {50, 60.75f}, ey[2] = {70, 80};
  int32_t lfid[2];
  double l2d[2][4];
  pack_line_channels(2, LineChannels{lid, sx, sy, ex, ey}, lfid, &l2d[0][0]);
  CHECK(lfid[1] == 41 && l2d[0][0] == (double)10.5f && l2d[1][2] == (double)60.75f && l2d[1][3] == 80.0);
}

static void test_host_only() {
  test_ingest();
  // ceres_compat CauchyLoss == oracle
  for (double a : {1.0, 2.5}) {
    ceres::CauchyLoss loss(a);
    for (double s : {0.0, 0.3, 7.0, 1e4}) {
      double r1[3], r2[3];
      loss.Evaluate(s, r1);
      orc_cauchy_loss(a, s, r2);
      CHECK(r1[0] == r2[0] && r1[1] == r2[1] && r1[2] == r2[2]);
    }
  }
  // MarginalizationFactor::Evaluate (host GEMV) == oracle, incl. the w < 0 branch
  MarginalizationInfo* info = new MarginalizationInfo();
  const int sizes[4] = {7, 9, 7, 1};
  info->m = 15;
  int pos = info->m, n = 0;
  std::vector<std::vector<double>> x0(4), x(4);
  for (int k = 0; k < 4; ++k) {
    info->keep_block_size.push_back(sizes[k]);
    info->keep_block_idx.push_back(pos);
    pos += sizes[k] == 7 ? 6 : sizes[k];
    x0[k].resize(sizes[k]);
    if (sizes[k] == 7) random_pose(x0[k].data(), 1.0, 0.5);
    else for (double& v : x0[k]) v = nrand();
    x[k] = x0[k];
    for (double& v : x[k]) v += 1e-2 * nrand();
  }
  for (int k = 3; k < 7; ++k) x[2][k] = -x[2][k];  // exercise the sign flip
  n = pos - info->m;
  info->n = n;
  for (int k = 0; k < 4; ++k) info->keep_block_data.push_back(x0[k].data());
  info->linearized_jacobians.resize(n, n);
  info->linearized_residuals.resize(n);
  std::vector<double> lj((size_t)n * n), lr(n);
  for (int r = 0; r < n; ++r) {
    lr[r] = nrand();
    info->linearized_residuals(r) = lr[r];
    for (int c = 0; c < n; ++c) lj[(size_t)r * n + c] = nrand(), info->linearized_jacobians(r, c) = lj[(size_t)r * n + c];
  }
  MarginalizationFactor mf(info);
  CHECK(mf.num_residuals() == n && mf.parameter_block_sizes().size() == 4);
  std::vector<double> r1(n), r2(n);
  std::vector<std::vector<double>> J1(4), J2(4);
  double *p1[4], *p2[4];
  const double* xp[4];
  const double* x0p[4];
  for (int k = 0; k < 4; ++k) J1[k].assign((size_t)n * sizes[k], 7.0), J2[k].assign((size_t)n * sizes[k], 9.0), p1[k] = J1[k].data(), p2[k] = J2[k].data(), xp[k] = x[k].data(), x0p[k] = x0[k].data();
  mf.Evaluate(xp, r1.data(), p1);
  orc_marginalization_factor_evaluate(n, info->m, 4, info->keep_block_size.data(), info->keep_block_idx.data(), x0p, lj.data(), lr.data(), xp, r2.data(), p2);
  CHECK(rel_err(r1.data(), r2.data(), n) < 1e-14);
  for (int k = 0; k < 4; ++k) CHECK(rel_err(J1[k].data(), J2[k].data(), J1[k].size()) == 0.0);
  info->keep_block_data.clear();  // not owned here
  delete info;
  // removeLineOutlier == oracle gate
  std::vector<viml::Vector3d> lv = {{1, 0, 0}, {1.05, 0, 0}, {2, 0, 0}, {1, 0.2, 0}};
  std::vector<bool> cred;
  const bool ok = LineMapAssociator::removeLineOutlier(lv, &cred);
  double flat[12];
  for (int k = 0; k < 4; ++k)
    for (int c = 0; c < 3; ++c) flat[3 * k + c] = lv[k][c];
  uint8_t oc[4];
  const int ook = orc_track_gate(4, flat, oc);
  CHECK(ok == (ook != 0));
  for (int k = 0; k < 4; ++k) CHECK(cred[k] == (oc[k] != 0));
}

struct Window {
  double para_Pose[11][7], para_SpeedBias[11][9], para_Ex_Pose[1][7], para_Feature[64][1];
};
static void make_window(Window& w) {
  double base[7];
  random_pose(base, 1.0, 0.3);
  for (int i = 0; i < 11; ++i) {
    std::memcpy(w.para_Pose[i], base, 56);
    for (int k = 0; k < 3; ++k) w.para_Pose[i][k] += 0.05 * i + 0.02 * nrand();
    for (int k = 3; k < 7; ++k) w.para_Pose[i][k] += 0.01 * nrand();
    double qn = 0.0;
    for (int k = 3; k < 7; ++k) qn += w.para_Pose[i][k] * w.para_Pose[i][k];
    for (int k = 3; k < 7; ++k) w.para_Pose[i][k] *= (1.0 + 1e-12 * nrand()) / std::sqrt(qn);  // unit up to 1e-12
    for (int k = 0; k < 9; ++k) w.para_SpeedBias[i][k] = 0.1 * nrand();
  }
  const double ric[7] = {-0.0216, -0.0647, 0.0098, 0.0077, -0.0105, 0.7018, 0.7123};
  std::memcpy(w.para_Ex_Pose[0], ric, 56);
  for (int l = 0; l < 64; ++l) w.para_Feature[l][0] = 0.15 + 0.3 * urand();
}

static void test_factors(const viml_config& cfg) {
  Window w;
  make_window(w);
  // ProjectionFactor::Evaluate, every NULL pattern
  for (int t = 0; t < 8; ++t) {
    viml::Vector3d pi(0.4 * nrand(), 0.3 * nrand(), 1.0), pj(0.4 * nrand(), 0.3 * nrand(), 1.0);
    ProjectionFactor f(pi, pj);
    const int i = t % 4, j = 4 + t % 7;
    const double* params[4] = {w.para_Pose[i], w.para_Pose[j], w.para_Ex_Pose[0], w.para_Feature[t]};
    double r[2], ro[2], J[4][14], Jo[4][14];
    double* jp[4] = {J[0], (t & 1) ? nullptr : J[1], J[2], (t & 2) ? nullptr : J[3]};
    double* jo[4] = {Jo[0], Jo[1], Jo[2], Jo[3]};
    CHECK(f.Evaluate(params, r, jp));
    orc_projection_evaluate(pi.data(), pj.data(), cfg.sqrt_info, params, ro, jo);
    CHECK(rel_err(r, ro, 2) < 1e-9);
    const int sz[4] = {14, 14, 14, 2};
    for (int k = 0; k < 4; ++k)
      if (jp[k]) CHECK(rel_err(J[k], Jo[k], sz[k]) < 1e-9);
    CHECK(f.Evaluate(params, r, nullptr));
    CHECK(rel_err(r, ro, 2) < 1e-9);
    if (t == 0) {
      double* pp[4] = {w.para_Pose[i], w.para_Pose[j], w.para_Ex_Pose[0], w.para_Feature[t]};
      const double worst = f.check(pp);
      std::printf("ProjectionFactor::check: max |analytic - numeric| / max |analytic| = %.3e\n", worst);
      CHECK(worst < 1e-3);  // forward differences with eps 1e-6
    }
  }
  // LineProjectionFactor::Evaluate
  viml::Matrix3d K, Ric;
  K(0, 0) = cfg.fx, K(1, 1) = cfg.fy, K(0, 2) = cfg.cx, K(1, 2) = cfg.cy, K(2, 2) = 1;
  {
    const double* e = w.para_Ex_Pose[0];
    const double n = std::sqrt(e[3] * e[3] + e[4] * e[4] + e[5] * e[5] + e[6] * e[6]);
    const double x = e[3] / n, y = e[4] / n, z = e[5] / n, ww = e[6] / n;
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - ww * z), 2 * (x * z + ww * y), 2 * (x * y + ww * z), 1 - 2 * (x * x + z * z),
                         2 * (y * z - ww * x), 2 * (x * z - ww * y), 2 * (y * z + ww * x), 1 - 2 * (x * x + y * y)};
    std::memcpy(Ric.m, R, sizeof(R));
  }
  viml::Vector3d Tic(w.para_Ex_Pose[0][0], w.para_Ex_Pose[0][1], w.para_Ex_Pose[0][2]);
  for (int t = 0; t < 4; ++t) {
    // a segment a few metres in front of pose t's camera (camera looks along body x-ish after Ric; just search)
    viml::Vector3d ps(w.para_Pose[t][0] + 3 + nrand(), w.para_Pose[t][1] + nrand(), w.para_Pose[t][2] + nrand());
    viml::Vector3d pe(ps.x() + 0.5 * nrand(), ps.y() + 1.0, ps.z() + 0.5 * nrand());
    viml::Vector3d abc(120.0 + 10 * nrand(), -80.0 + 10 * nrand(), 1e4 * nrand());
    LineProjectionFactor f(ps, pe, abc, K, Ric, Tic);
    const double* params[1] = {w.para_Pose[t]};
    double r[2], ro[2], J[14], Jo[14];
    double *jp[1] = {J}, *jo[1] = {Jo};
    CHECK(f.Evaluate(params, r, jp));
    orc_line_evaluate(ps.data(), pe.data(), abc.data(), K.m, Ric.m, Tic.data(), params, ro, jo);
    CHECK(rel_err(r, ro, 2) < 1e-9 && rel_err(J, Jo, 14) < 1e-9);
  }
  // LinearizationBatch (the Ceres EvaluationCallback bridge): one launch, then per-factor fetches
  {
    LinearizationBatch batch;
    std::vector<ProjectionFactor*> fs;
    std::vector<std::array<double*, 4>> ps;
    for (int l = 0; l < 40; ++l) {
      const int i = l % 5;
      for (int j = i + 1; j < i + 4; ++j) {
        auto* f = new ProjectionFactor(viml::Vector3d(0.4 * nrand(), 0.3 * nrand(), 1.0), viml::Vector3d(0.4 * nrand(), 0.3 * nrand(), 1.0));
        batch.AddResidualBlock(f, w.para_Pose[i], w.para_Pose[j], w.para_Ex_Pose[0], w.para_Feature[l]);
        fs.push_back(f);
        ps.push_back({w.para_Pose[i], w.para_Pose[j], w.para_Ex_Pose[0], w.para_Feature[l]});
      }
    }
    auto* lf = new LineProjectionFactor(viml::Vector3d(w.para_Pose[2][0] + 3, w.para_Pose[2][1], w.para_Pose[2][2]),
                                        viml::Vector3d(w.para_Pose[2][0] + 3, w.para_Pose[2][1] + 1, w.para_Pose[2][2] + 0.2),
                                        viml::Vector3d(100, -50, 3000), K, Ric, Tic);
    batch.AddResidualBlock(lf, w.para_Pose[2]);
    for (int round = 0; round < 2; ++round) {
      if (round == 1) w.para_Pose[3][0] += 0.01, w.para_Feature[7][0] *= 1.01;  // a new evaluation point
      batch.PrepareForEvaluation(true, true);
      for (size_t k = 0; k < fs.size(); ++k) {
        const double* params[4] = {ps[k][0], ps[k][1], ps[k][2], ps[k][3]};
        double r[2], ro[2], J[4][14], Jo[4][14];
        double *jp[4] = {J[0], J[1], J[2], J[3]}, *jo[4] = {Jo[0], Jo[1], Jo[2], Jo[3]};
        fs[k]->Evaluate(params, r, jp);
        orc_projection_evaluate(fs[k]->pts_i.data(), fs[k]->pts_j.data(), cfg.sqrt_info, params, ro, jo);
        CHECK(rel_err(r, ro, 2) < 1e-9 && rel_err(J[0], Jo[0], 14) < 1e-9 && rel_err(J[1], Jo[1], 14) < 1e-9 &&
              rel_err(J[2], Jo[2], 14) < 1e-9 && rel_err(J[3], Jo[3], 2) < 1e-9);
      }
      const double* lp[1] = {w.para_Pose[2]};
      double r[2], ro[2], J[14], Jo[14];
      double *jp[1] = {J}, *jo[1] = {Jo};
      lf->Evaluate(lp, r, jp);
      orc_line_evaluate(lf->pts_start.data(), lf->pts_end.data(), lf->line_param.data(), K.m, Ric.m, Tic.data(), lp, ro, jo);
      CHECK(rel_err(r, ro, 2) < 1e-9 && rel_err(J, Jo, 14) < 1e-9);
    }
    // What ceres::Solve does: ResidualBlock::Evaluate hands out pointers into the minimizer's own state vector, never
    // the user's arrays.  Copies of the evaluation point must be served from the batch (no per-factor launch) ...
    {
      batch.PrepareForEvaluation(true, true);
      const long before = batch.served();
      for (size_t k = 0; k < fs.size(); ++k) {
        double cp[4][7];
        for (int b = 0; b < 4; ++b) std::memcpy(cp[b], ps[k][b], (b == 3 ? 1 : 7) * sizeof(double));
        const double* params[4] = {cp[0], cp[1], cp[2], cp[3]};
        double r[2], ro[2], J[4][14], Jo[4][14];
        double *jp[4] = {J[0], J[1], J[2], J[3]}, *jo[4] = {Jo[0], Jo[1], Jo[2], Jo[3]};
        fs[k]->Evaluate(params, r, jp);
        orc_projection_evaluate(fs[k]->pts_i.data(), fs[k]->pts_j.data(), cfg.sqrt_info, params, ro, jo);
        CHECK(rel_err(r, ro, 2) < 1e-9 && rel_err(J[0], Jo[0], 14) < 1e-9 && rel_err(J[3], Jo[3], 2) < 1e-9);
      }
      CHECK(batch.served() - before == (long)fs.size());
      // ... and a state that is NOT the evaluation point must not be (it is evaluated on its own, correctly)
      double cp[4][7];
      for (int b = 0; b < 4; ++b) std::memcpy(cp[b], ps[0][b], (b == 3 ? 1 : 7) * sizeof(double));
      cp[1][1] += 0.02;
      const double* params[4] = {cp[0], cp[1], cp[2], cp[3]};
      double r[2], ro[2], J[4][14], Jo[4][14];
      double *jp[4] = {J[0], J[1], J[2], J[3]}, *jo[4] = {Jo[0], Jo[1], Jo[2], Jo[3]};
      const long b2 = batch.served();
      fs[0]->Evaluate(params, r, jp);
      orc_projection_evaluate(fs[0]->pts_i.data(), fs[0]->pts_j.data(), cfg.sqrt_info, params, ro, jo);
      CHECK(batch.served() == b2 && rel_err(r, ro, 2) < 1e-9 && rel_err(J[1], Jo[1], 14) < 1e-9);
      batch.Invalidate();
      const double* up[4] = {ps[0][0], ps[0][1], ps[0][2], ps[0][3]};
      fs[0]->Evaluate(up, r, jp);
      CHECK(batch.served() == b2);
    }
    for (auto* f : fs) delete f;
    delete lf;
  }
}

// MARGIN_OLD of Estimator::OptimizationWithLine (estimator.cpp:1911-2044) against the literal dense algorithm.
static void test_marginalization(const viml_config& cfg) {
  Window w;
  make_window(w);
  MarginalizationInfo* info = new MarginalizationInfo();
  ceres::LossFunction* loss_function = new ceres::CauchyLoss(1.0);
  struct Rec { ceres::CostFunction* f; ceres::LossFunction* loss; std::vector<double*> blocks; };
  std::vector<Rec> recs;
  auto add = [&](ceres::CostFunction* f, ceres::LossFunction* loss, std::vector<double*> blocks, std::vector<int> drop) {
    info->addResidualBlockInfo(new ResidualBlockInfo(f, loss, blocks, drop));
    recs.push_back({f, loss, blocks});
  };
  // "last prior": a dense factor on pose0, sb0, pose1..3, ex (drop pose0, sb0)  (:1916-1931)
  add(new DenseFactor(40, {7, 9, 7, 7, 7, 7}), nullptr,
      {w.para_Pose[0], w.para_SpeedBias[0], w.para_Pose[1], w.para_Pose[2], w.para_Pose[3], w.para_Ex_Pose[0]}, {0, 1});
  // "IMU factor" between frames 0 and 1 (:1934-1942)
  add(new DenseFactor(15, {7, 9, 7, 9}), nullptr, {w.para_Pose[0], w.para_SpeedBias[0], w.para_Pose[1], w.para_SpeedBias[1]}, {0, 1});
  // every ProjectionFactor of features that start in frame 0 (:1945-1990)
  for (int l = 0; l < 30; ++l)
    for (int j = 1; j <= 1 + l % 5; ++j)
      add(new ProjectionFactor(viml::Vector3d(0.4 * nrand(), 0.3 * nrand(), 1.0), viml::Vector3d(0.4 * nrand(), 0.3 * nrand(), 1.0)),
          loss_function, {w.para_Pose[0], w.para_Pose[j], w.para_Ex_Pose[0], w.para_Feature[l]}, {0, 3});
  info->preMarginalize();
  info->marginalize();
  CHECK(info->m == 6 + 9 + 30);
  // literal dense reference in the shim's block order
  const int pos = info->m + info->n;
  std::vector<double> A((size_t)pos * pos, 0.0), b(pos, 0.0);
  for (const Rec& rc : recs) {
    const auto& sizes = rc.f->parameter_block_sizes();
    const int nres = rc.f->num_residuals(), nb = (int)sizes.size();
    std::vector<double> r(nres);
    std::vector<std::vector<double>> J(nb);
    std::vector<double*> jp(nb);
    std::vector<int> sz(sizes.begin(), sizes.end());
    for (int k = 0; k < nb; ++k) J[k].assign((size_t)nres * sizes[k], 0.0), jp[k] = J[k].data();
    if (auto* pf = dynamic_cast<ProjectionFactor*>(rc.f)) {
      const double* params[4] = {rc.blocks[0], rc.blocks[1], rc.blocks[2], rc.blocks[3]};
      orc_projection_evaluate(pf->pts_i.data(), pf->pts_j.data(), cfg.sqrt_info, params, r.data(), jp.data());
    } else {
      rc.f->Evaluate(rc.blocks.data(), r.data(), jp.data());
    }
    if (rc.loss) orc_loss_correct(1.0, nres, r.data(), nb, sz.data(), jp.data());
    for (int i = 0; i < nb; ++i) {
      const int idx_i = info->parameter_block_idx[reinterpret_cast<long>(rc.blocks[i])], si = sizes[i] == 7 ? 6 : sizes[i];
      for (int j = i; j < nb; ++j) {
        const int idx_j = info->parameter_block_idx[reinterpret_cast<long>(rc.blocks[j])], sj = sizes[j] == 7 ? 6 : sizes[j];
        for (int rr = 0; rr < si; ++rr)
          for (int cc = 0; cc < sj; ++cc) {
            double v = 0.0;
            for (int q = 0; q < nres; ++q) v += J[i][(size_t)q * sizes[i] + rr] * J[j][(size_t)q * sizes[j] + cc];
            A[(size_t)(idx_i + rr) * pos + idx_j + cc] += v;
            if (i != j) A[(size_t)(idx_j + cc) * pos + idx_i + rr] = A[(size_t)(idx_i + rr) * pos + idx_j + cc];
          }
      }
      for (int rr = 0; rr < si; ++rr) {
        double v = 0.0;
        for (int q = 0; q < nres; ++q) v += J[i][(size_t)q * sizes[i] + rr] * r[q];
        b[idx_i + rr] += v;
      }
    }
  }
  const int n = info->n;
  std::vector<double> As((size_t)n * n), bs(n), lj((size_t)n * n), lr(n);
  orc_marginalize_dense(A.data(), b.data(), pos, info->m, 1e-8, As.data(), bs.data(), lj.data(), lr.data());
  CHECK(rel_err(info->A_schur.data(), As.data(), As.size()) < 1e-9);
  CHECK(rel_err(info->b_schur.data(), bs.data(), bs.size()) < 1e-9);
  // eigenvector signs are not unique: compare J^T J and J^T r
  std::vector<double> JtJ((size_t)n * n, 0.0), JtJo((size_t)n * n, 0.0), Jtr(n, 0.0), Jtro(n, 0.0);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      for (int k = 0; k < n; ++k) {
        JtJ[(size_t)r * n + c] += info->linearized_jacobians(k, r) * info->linearized_jacobians(k, c);
        JtJo[(size_t)r * n + c] += lj[(size_t)k * n + r] * lj[(size_t)k * n + c];
      }
    }
  for (int r = 0; r < n; ++r)
    for (int k = 0; k < n; ++k) Jtr[r] += info->linearized_jacobians(k, r) * info->linearized_residuals(k), Jtro[r] += lj[(size_t)k * n + r] * lr[k];
  CHECK(rel_err(JtJ.data(), JtJo.data(), JtJ.size()) < 1e-9);
  CHECK(rel_err(Jtr.data(), Jtro.data(), Jtr.size()) < 1e-8);
  // the new prior as a cost function (estimator.cpp:2027-2044, :1717-1719)
  std::unordered_map<long, double*> addr_shift;
  for (int i = 1; i <= 10; i++) {
    addr_shift[reinterpret_cast<long>(w.para_Pose[i])] = w.para_Pose[i - 1];
    addr_shift[reinterpret_cast<long>(w.para_SpeedBias[i])] = w.para_SpeedBias[i - 1];
  }
  addr_shift[reinterpret_cast<long>(w.para_Ex_Pose[0])] = w.para_Ex_Pose[0];
  std::vector<double*> blocks = info->getParameterBlocks(addr_shift);
  CHECK(blocks.size() == info->keep_block_size.size() && !blocks.empty());
  CHECK(info->sum_block_size == 7 * 6 + 9);  // pose1..5, ex, speed-bias 1
  MarginalizationFactor mf(info);
  std::vector<double> r(n);
  std::vector<const double*> params;
  for (double* p : info->keep_block_data) params.push_back(p);
  mf.Evaluate(params.data(), r.data(), nullptr);  // at the linearisation point the prior residual is linearized_residuals
  CHECK(rel_err(r.data(), info->linearized_residuals.data(), n) < 1e-15);
  delete info;  // owns and frees every cost function (marginalization_factor.cpp:71-87)
  delete loss_function;
}

static void test_association(const viml_config& cfg) {
  // a small map on a plane 6..12 m in front of the first camera, poses drifting slowly
  Window w;
  make_window(w);
  for (int i = 0; i < 11; ++i) {  // identity-ish orientation so that the geometry below is in view
    w.para_Pose[i][3] = 0.01 * nrand(), w.para_Pose[i][4] = 0.01 * nrand(), w.para_Pose[i][5] = 0.01 * nrand(), w.para_Pose[i][6] = 1.0;
    w.para_Pose[i][0] = 0.05 * i, w.para_Pose[i][1] = 0.02 * nrand(), w.para_Pose[i][2] = 0.02 * nrand();
  }
  const double exq[7] = {0.01, -0.02, 0.03, 0.005, -0.004, 0.003, 1.0};  // camera ~ body frame
  std::memcpy(w.para_Ex_Pose[0], exq, 56);
  // p_cam = Ric^T (Rbi^T (Rbw p_map + Tbw - Tbi) - Tic); put map points so that Rbw p + Tbw is in front (z > 0)
  std::vector<viml::Vector6d> map;
  const double* R = cfg.Rbw;
  const double* T = cfg.Tbw;
  auto to_map = [&](double x, double y, double z, double* o) {  // p_map = Rbw^T (p_vio - Tbw)
    const double d[3] = {x - T[0], y - T[1], z - T[2]};
    for (int c = 0; c < 3; ++c) o[c] = R[c] * d[0] + R[3 + c] * d[1] + R[6 + c] * d[2];
  };
  for (int k = 0; k < 700; ++k) {
    viml::Vector6d l;
    const double x = 8.0 * (urand() - 0.5) * 3, y = 6.0 * (urand() - 0.5) * 3, z = 6 + 6 * urand();
    to_map(x, y, z, &l[0]);
    to_map(x + 1.5 * nrand(), y + 1.5 * nrand(), z + 0.5 * nrand(), &l[3]);
    map.push_back(l);
  }
  LineMapAssociator assoc(map);
  std::vector<int32_t> ref(map.size());
  for (int i = 0; i < 11; ++i) {
    const int cnt = assoc.UpdateLinesInFoV(i, w.para_Pose[i], w.para_Ex_Pose[0]);
    const int rc = orc_update_lines_in_fov(&cfg, w.para_Pose[i], w.para_Ex_Pose[0], map[0].data(), (int64_t)map.size(), ref.data());
    CHECK(cnt == rc && cnt > 20);
    CHECK(std::equal(ref.begin(), ref.begin() + rc, assoc.WorldLinesInFOV[i].begin()));
  }
  // the optimiser moved the poses: matching uses the new ones against the cached lists
  Window w2 = w;
  for (int i = 0; i < 11; ++i) w2.para_Pose[i][0] += 0.01 * nrand(), w2.para_Pose[i][4] += 0.002 * nrand();
  w2.para_Ex_Pose[0][1] += 0.001;
  std::vector<LineMapAssociator::Observation> obs;
  for (int k = 0; k < 200; ++k) {
    LineMapAssociator::Observation o;
    o.frame = k % 11;
    const double u = 752 * urand(), v = 480 * urand(), a = 6.28 * urand(), len = 60 + 200 * urand();
    const float e[4] = {(float)u, (float)v, (float)(u + len * std::cos(a)), (float)(v + len * std::sin(a))};
    for (int c = 0; c < 4; ++c) o.line[c] = e[c];
    obs.push_back(o);
  }
  std::vector<LineMapAssociator::Match> out;
  const int matched = assoc.updateLinePairInWindow(w2.para_Pose, w2.para_Ex_Pose[0], obs, &out);
  int ref_matched = 0;
  for (size_t k = 0; k < obs.size(); ++k) {
    const int f = obs[k].frame;
    float err[3];
    double proj[4] = {0, 0, 0, 0};
    const int idx = orc_line_correspondence(&cfg, w2.para_Pose[f], w2.para_Ex_Pose[0], map[0].data(), assoc.WorldLinesInFOV[f].data(),
                                            (int)assoc.WorldLinesInFOV[f].size(), obs[k].line, err, proj);
    ref_matched += idx >= 0;
    CHECK(out[k].map_index == idx && out[k].errD == err[1] && out[k].overlap == err[2]);
    CHECK(std::fabs(out[k].errA - err[0]) <= 1.2e-7f * std::fabs(err[0]));
    CHECK(out[k].credible_line == (idx >= 0) && out[k].use_flag);
    if (idx >= 0) CHECK(std::memcmp(out[k].projectedLine, proj, 32) == 0);
    if (k < 5) {
      LineMapAssociator::Match one = assoc.LineCorrespondenceInFrame(f, obs[k].line, w2.para_Pose, w2.para_Ex_Pose[0]);
      CHECK(one.map_index == idx && one.errD == err[1]);
    }
  }
  CHECK(matched == ref_matched && matched > 5);
  const std::vector<int> slot1 = assoc.WorldLinesInFOV[1];
  assoc.slideWindowOld();
  CHECK(assoc.WorldLinesInFOV[0] == slot1);
  std::printf("association: %d/%zu observations matched\n", matched, obs.size());
}

// BASELINE.json configs[0] through the drop-in path: ONE EuRoC-shaped window (11 poses, 150 features, ~600 ProjectionFactors,
// 110 LineProjectionFactors), one evaluation point = LinearizationBatch::PrepareForEvaluation (pack, H2D, kernels, D2H) + one
// Evaluate per residual block called the way ceres::Solve calls it (copies of the state, all Jacobians), next to the
// reference's own per-factor CPU Evaluate (oracle port, single thread, like the reference's single process() thread,
// estimator.cpp:1900).  Prints one JSON object.
static int bench_cfg1(const viml_config& cfg, int reps) {
  static double pose[11][7], ex[7], feat[150][1];
  double base[7];
  random_pose(base, 1.0, 0.3);
  for (int i = 0; i < 11; ++i) {
    std::memcpy(pose[i], base, 56);
    for (int k = 0; k < 3; ++k) pose[i][k] += 0.05 * i + 0.02 * nrand();
  }
  const double ric[7] = {-0.0216, -0.0647, 0.0098, 0.0077, -0.0105, 0.7018, 0.7123};
  std::memcpy(ex, ric, 56);
  for (int l = 0; l < 150; ++l) feat[l][0] = 0.15 + 0.3 * urand();
  viml::Matrix3d K, Ric;
  K(0, 0) = cfg.fx, K(1, 1) = cfg.fy, K(0, 2) = cfg.cx, K(1, 2) = cfg.cy, K(2, 2) = 1;
  {
    const double n = std::sqrt(ex[3] * ex[3] + ex[4] * ex[4] + ex[5] * ex[5] + ex[6] * ex[6]);
    const double x = ex[3] / n, y = ex[4] / n, z = ex[5] / n, ww = ex[6] / n;
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - ww * z), 2 * (x * z + ww * y), 2 * (x * y + ww * z), 1 - 2 * (x * x + z * z),
                         2 * (y * z - ww * x), 2 * (x * z - ww * y), 2 * (y * z + ww * x), 1 - 2 * (x * x + y * y)};
    std::memcpy(Ric.m, R, sizeof(R));
  }
  viml::Vector3d Tic(ex[0], ex[1], ex[2]);
  LinearizationBatch batch;
  struct PFE { ProjectionFactor* f; int i, j, l; };
  std::vector<PFE> pfs;
  std::vector<LineProjectionFactor*> lfs;
  std::vector<int> lframe;
  for (int l = 0; l < 150; ++l) {
    const int i = (int)(urand() * 8) % 8, len = 2 + (int)(urand() * (10 - i)) % (10 - i);   // start_frame < 8, track to frame <= 10
    for (int j = i + 1; j < i + len && j < 11; ++j) {
      auto* f = new ProjectionFactor(viml::Vector3d(0.4 * nrand(), 0.3 * nrand(), 1.0), viml::Vector3d(0.4 * nrand(), 0.3 * nrand(), 1.0));
      batch.AddResidualBlock(f, pose[i], pose[j], ex, feat[l]);
      pfs.push_back({f, i, j, l});
    }
  }
  for (int p = 0; p < 11; ++p)
    for (int q = 0; q < 10; ++q) {
      auto* f = new LineProjectionFactor(viml::Vector3d(pose[p][0] + 3 + nrand(), pose[p][1] + nrand(), pose[p][2] + nrand()),
                                         viml::Vector3d(pose[p][0] + 3 + nrand(), pose[p][1] + 1 + nrand(), pose[p][2] + nrand()),
                                         viml::Vector3d(120.0 + 10 * nrand(), -80.0 + 10 * nrand(), 1e4 * nrand()), K, Ric, Tic);
      batch.AddResidualBlock(f, pose[p]);
      lfs.push_back(f), lframe.push_back(p);
    }
  auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };
  double sink = 0.0;
  auto gpu_point = [&](double* t_prepare) {
    pose[3][0] += 1e-4;   // a new evaluation point
    const double t0 = now();
    batch.PrepareForEvaluation(true, true);
    const double t1 = now();
    for (const PFE& e : pfs) {
      double cp[4][7], r[2], J[4][14];
      std::memcpy(cp[0], pose[e.i], 56), std::memcpy(cp[1], pose[e.j], 56), std::memcpy(cp[2], ex, 56), cp[3][0] = feat[e.l][0];
      const double* params[4] = {cp[0], cp[1], cp[2], cp[3]};
      double* jp[4] = {J[0], J[1], J[2], J[3]};
      e.f->Evaluate(params, r, jp);
      sink += r[0] + J[0][0];
    }
    for (size_t k = 0; k < lfs.size(); ++k) {
      double cp[7], r[2], J[14];
      std::memcpy(cp, pose[lframe[k]], 56);
      const double* params[1] = {cp};
      double* jp[1] = {J};
      lfs[k]->Evaluate(params, r, jp);
      sink += r[0] + J[0];
    }
    *t_prepare += t1 - t0;
    return now() - t0;
  };
  auto cpu_point = [&] {
    const double t0 = now();
    for (const PFE& e : pfs) {
      double r[2], J[4][14];
      const double* params[4] = {pose[e.i], pose[e.j], ex, feat[e.l]};
      double* jp[4] = {J[0], J[1], J[2], J[3]};
      orc_projection_evaluate(e.f->pts_i.data(), e.f->pts_j.data(), cfg.sqrt_info, params, r, jp);
      sink += r[0] + J[0][0];
    }
    for (size_t k = 0; k < lfs.size(); ++k) {
      double r[2], J[14];
      const double* params[1] = {pose[lframe[k]]};
      double* jp[1] = {J};
      orc_line_evaluate(lfs[k]->pts_start.data(), lfs[k]->pts_end.data(), lfs[k]->line_param.data(), K.m, Ric.m, Tic.data(), params, r, jp);
      sink += r[0] + J[0];
    }
    return now() - t0;
  };
  double tp = 0.0, tg = 0.0, tc = 0.0;
  for (int w = 0; w < 5; ++w) gpu_point(&tp), cpu_point();
  tp = 0.0;
  const long served0 = batch.served();
  for (int r = 0; r < reps; ++r) tg += gpu_point(&tp);
  const long served = batch.served() - served0;
  for (int r = 0; r < reps; ++r) tc += cpu_point();
  const size_t nfac = pfs.size() + lfs.size();
  std::printf("{\"workload\": \"cfg1: one window, %zu ProjectionFactors + %zu LineProjectionFactors, one evaluation point = "
              "PrepareForEvaluation + one Evaluate per residual block with copied parameter blocks\", \"evaluation_points\": %d, "
              "\"gpu_bridge_ms_per_point\": %.4f, \"gpu_prepare_ms_per_point\": %.4f, \"served_from_batch_per_point\": %.1f, "
              "\"cpu_reference_ms_per_point\": %.4f, \"cpu_threads\": 1, \"gpu_over_cpu\": %.3f, \"sink\": %.3e}\n",
              pfs.size(), lfs.size(), reps, 1e3 * tg / reps, 1e3 * tp / reps, (double)served / reps, 1e3 * tc / reps, tc / tg, sink);
  (void)nfac;
  for (auto& e : pfs) delete e.f;
  for (auto* f : lfs) delete f;
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::strcmp(argv[1], "--bench-cfg1") == 0) {
    const viml_config cfg = euroc_config();
    if (viml::Runtime::instance().configure(cfg, 0) != VIML_OK) {
      std::printf("{\"error\": \"no GPU\"}\n");
      return 2;
    }
    const int rc = bench_cfg1(cfg, argc > 2 ? std::atoi(argv[2]) : 200);
    viml::Runtime::instance().shutdown();
    return rc;
  }
  const bool cpu_only = argc > 1 && std::strcmp(argv[1], "--cpu") == 0;
  test_host_only();
  if (!cpu_only) {
    const viml_config cfg = euroc_config();
    const int rc = viml::Runtime::instance().configure(cfg, 0);
    if (rc != VIML_OK) {
      std::printf("FAIL viml::Runtime::configure -> %d (no GPU?)\n", rc);
      return 2;
    }
    test_factors(cfg);
    test_marginalization(cfg);
    test_association(cfg);
    viml::Runtime::instance().shutdown();
  }
  if (g_fail) std::printf("selftest: %d FAILED\n", g_fail);
  else std::printf("selftest: all passed%s\n", cpu_only ? " (host-only subset)" : "");
  return g_fail ? 1 : 0;
}
