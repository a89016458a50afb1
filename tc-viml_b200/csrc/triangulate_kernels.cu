// triangulate_kernels.cu — FeatureManager::triangulate (feature_manager.cpp:440-492) for a batch of features.
//
// Per feature: the camera of the start frame is the reference frame (P0 = [I | 0]); every observation in frame j adds the two
// rows  f.x P.row(2) - f.z P.row(0),  f.y P.row(2) - f.z P.row(1)  with P = [R^T | -R^T t], R = R0^T R1, t = R0^T (t1 - t0),
// f = point.normalized() (:468-472); the right singular vector of the smallest singular value gives depth = V[2] / V[3]
// (:477-478), and a depth below 0.1 is replaced by INIT_DEPTH (:482-485).  Eigen's JacobiSVD is restated as a one-sided
// (Hestenes) Jacobi SVD on the 2n x 4 matrix — it works on A itself, not on A^T A, so the small singular vector keeps the
// accuracy of an SVD.  One thread per feature (n <= 32 observations; the window has 11 frames).
#include "common.cuh"

namespace {

constexpr int kMaxObs = 32;

__device__ __forceinline__ void quat_rot_normalized(const double* q7, double* R) {
  const double n = sqrt(q7[3] * q7[3] + q7[4] * q7[4] + q7[5] * q7[5] + q7[6] * q7[6]);
  const double x = q7[3] / n, y = q7[4] / n, z = q7[5] / n, w = q7[6] / n;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz), R[1] = txy - twz, R[2] = txz + twy;
  R[3] = txy + twz, R[4] = 1.0 - (txx + tzz), R[5] = tyz - twx;
  R[6] = txz - twy, R[7] = tyz + twx, R[8] = 1.0 - (txx + tyy);
}

__global__ void __launch_bounds__(64) triangulate_kernel(int P, int64_t NF, const double* __restrict__ poses, const double* __restrict__ ex,
                                                         const int32_t* __restrict__ fwin, const int32_t* __restrict__ start,
                                                         const int64_t* __restrict__ off, const double* __restrict__ pts, double init_depth,
                                                         double* __restrict__ depth) {
  const int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (l >= NF) return;
  const int w = fwin[l], i0 = start[l];
  const int n = (int)min((int64_t)kMaxObs, off[l + 1] - off[l]);
  if (n < 1 || i0 < 0 || i0 + n > P) {
    depth[l] = init_depth;
    return;
  }
  double ric[9], R0b[9], R0[9], t0[3];
  const double* e = ex + (size_t)w * 7;
  quat_rot_normalized(e, ric);
  const double* p0 = poses + ((size_t)w * P + i0) * 7;
  quat_rot_normalized(p0, R0b);
  for (int r = 0; r < 3; ++r) {
    t0[r] = p0[r] + R0b[3 * r] * e[0] + R0b[3 * r + 1] * e[1] + R0b[3 * r + 2] * e[2];                        // Ps + Rs tic
    for (int c = 0; c < 3; ++c) R0[3 * r + c] = R0b[3 * r] * ric[c] + R0b[3 * r + 1] * ric[3 + c] + R0b[3 * r + 2] * ric[6 + c];   // Rs ric
  }
  double A[2 * kMaxObs][4];
  for (int k = 0; k < n; ++k) {
    const double* pj = poses + ((size_t)w * P + i0 + k) * 7;
    double Rb[9], R1[9], t1[3];
    quat_rot_normalized(pj, Rb);
    for (int r = 0; r < 3; ++r) {
      t1[r] = pj[r] + Rb[3 * r] * e[0] + Rb[3 * r + 1] * e[1] + Rb[3 * r + 2] * e[2];
      for (int c = 0; c < 3; ++c) R1[3 * r + c] = Rb[3 * r] * ric[c] + Rb[3 * r + 1] * ric[3 + c] + Rb[3 * r + 2] * ric[6 + c];
    }
    // t = R0^T (t1 - t0), R = R0^T R1, P = [R^T | -R^T t]
    double d[3] = {t1[0] - t0[0], t1[1] - t0[1], t1[2] - t0[2]}, t[3], R[9];
    for (int r = 0; r < 3; ++r) {
      t[r] = R0[r] * d[0] + R0[3 + r] * d[1] + R0[6 + r] * d[2];
      for (int c = 0; c < 3; ++c) R[3 * r + c] = R0[r] * R1[c] + R0[3 + r] * R1[3 + c] + R0[6 + r] * R1[6 + c];
    }
    double Pm[3][4];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) Pm[r][c] = R[3 * c + r];
      Pm[r][3] = -(R[r] * t[0] + R[3 + r] * t[1] + R[6 + r] * t[2]);
    }
    const double* pt = pts + (size_t)(off[l] + k) * 3;
    const double nn = sqrt(pt[0] * pt[0] + pt[1] * pt[1] + pt[2] * pt[2]);
    const double f0 = pt[0] / nn, f1 = pt[1] / nn, f2 = pt[2] / nn;
    for (int c = 0; c < 4; ++c) {
      A[2 * k][c] = f0 * Pm[2][c] - f2 * Pm[0][c];
      A[2 * k + 1][c] = f1 * Pm[2][c] - f2 * Pm[1][c];
    }
  }
  // one-sided Jacobi: rotate column pairs until orthogonal; V accumulates the rotations
  const int m = 2 * n;
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    double offn = 0.0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        double a = 0.0, b = 0.0, g = 0.0;
        for (int r = 0; r < m; ++r) a += A[r][p] * A[r][p], b += A[r][q] * A[r][q], g += A[r][p] * A[r][q];
        if (g == 0.0) continue;
        offn = fmax(offn, fabs(g) / sqrt(fmax(a * b, 1e-300)));
        const double zeta = (b - a) / (2.0 * g);
        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
        for (int r = 0; r < m; ++r) {
          const double x = A[r][p], y = A[r][q];
          A[r][p] = c * x - s * y, A[r][q] = s * x + c * y;
        }
        for (int r = 0; r < 4; ++r) {
          const double x = V[r][p], y = V[r][q];
          V[r][p] = c * x - s * y, V[r][q] = s * x + c * y;
        }
      }
    if (offn < 1e-15) break;
  }
  int best = 0;
  double bn = 1e300;
  for (int c = 0; c < 4; ++c) {
    double s = 0.0;
    for (int r = 0; r < m; ++r) s += A[r][c] * A[r][c];
    if (s < bn) bn = s, best = c;
  }
  double dep = V[2][best] / V[3][best];   // svd_V[2] / svd_V[3]
  if (dep < 0.1) dep = init_depth;         // also catches NaN? no: NaN < 0.1 is false, like the reference
  depth[l] = dep;
}

}  // namespace

int viml_launch_triangulate(viml_ctx* ctx, int P, int64_t NF, const double* poses, const double* ex, const int32_t* fwin,
                            const int32_t* start, const int64_t* off, const double* pts, double init_depth, double* depth) {
  LaunchScope ls(ctx, K_GN);
  triangulate_kernel<<<(unsigned)((NF + 63) / 64), 64, 0, ctx->stream>>>(P, NF, poses, ex, fwin, start, off, pts, init_depth, depth);
  VIML_TRY_CUDA(ctx, cudaGetLastError());
  return VIML_OK;
}
