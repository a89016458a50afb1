// marg_kernels.cu — dense numeric core of MarginalizationInfo::marginalize, batched: one CTA per problem.
//
// Reference: /root/reference/vins_estimator/src/factor/marginalization_factor.cpp:264-293
//   Amm <- (Amm + Amm^T)/2; Amm^+ = V diag(l > eps ? 1/l : 0) V^T (SelfAdjointEigenSolver);
//   A = Arr - Arm Amm^+ Amr; b = brr - Arm Amm^+ bmm; eig(A) -> linearized_jacobians = sqrt(S) V^T,
//   linearized_residuals = sqrt(S^+) V^T b.
// The eigen-decomposition is a parallel two-sided cyclic Jacobi (round-robin pair ordering, n/2 disjoint
// rotations per round); eigen pairs are sorted ascending like Eigen.  Work matrices live in a per-CTA slice
// of device scratch (L1/L2 resident: <= 6 * 256^2 doubles).
#include "common.cuh"

namespace {

constexpr int kMargThreads = 256;

// Jacobi on the symmetric n x n matrix a (row-major, fully populated), eigenvectors in the columns of v.
__device__ void block_jacobi(double* __restrict__ a, double* __restrict__ v, int n, double* red, int* pq,
                             double* cs) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < n * n; e += nt) v[e] = (e / n == e % n) ? 1.0 : 0.0;
  const int ne = (n + 1) & ~1;  // players in the round-robin tournament (one dummy if n is odd)
  const int half = ne / 2;
  __syncthreads();
  for (int sweep = 0; sweep < 40; ++sweep) {
    // convergence: off(A)^2 <= 1e-30 * ||A||_F^2
    double off = 0.0, all = 0.0;
    for (int e = tid; e < n * n; e += nt) {
      const double x = a[e] * a[e];
      all += x;
      if (e / n != e % n) off += x;
    }
    for (int d = 16; d > 0; d >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, d);
      all += __shfl_xor_sync(0xffffffffu, all, d);
    }
    if ((tid & 31) == 0) red[tid >> 5] = off, red[32 + (tid >> 5)] = all;
    __syncthreads();
    if (tid == 0) {
      double o = 0.0, s = 0.0;
      for (int k = 0; k < (nt >> 5); ++k) o += red[k], s += red[32 + k];
      red[64] = (o <= 1e-30 * s || o == 0.0) ? 1.0 : 0.0;
    }
    __syncthreads();
    const bool done = red[64] != 0.0;
    __syncthreads();
    if (done) break;
    for (int round = 0; round < ne - 1; ++round) {
      // circle method: player ne-1 fixed, the others rotate
      for (int k = tid; k < half; k += nt) {
        int p = (k == 0) ? ne - 1 : (round + k) % (ne - 1);
        int q = (round + (ne - 1) - k) % (ne - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < n) {
          const double apq = a[p * n + q];
          if (apq != 0.0) {
            const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            s = t * c;
          }
        } else {
          p = q = -1;
        }
        pq[2 * k] = p, pq[2 * k + 1] = q;
        cs[2 * k] = c, cs[2 * k + 1] = s;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int e = tid; e < n * half; e += nt) {
        const int i = e / half, k = e % half;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p < 0) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        const double aip = a[i * n + p], aiq = a[i * n + q];
        a[i * n + p] = c * aip - s * aiq;
        a[i * n + q] = s * aip + c * aiq;
        const double vip = v[i * n + p], viq = v[i * n + q];
        v[i * n + p] = c * vip - s * viq;
        v[i * n + q] = s * vip + c * viq;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int e = tid; e < n * half; e += nt) {
        const int k = e / n, j = e % n;
        const int p = pq[2 * k], q = pq[2 * k + 1];
        if (p < 0) continue;
        const double c = cs[2 * k], s = cs[2 * k + 1];
        const double apj = a[p * n + j], aqj = a[q * n + j];
        a[p * n + j] = c * apj - s * aqj;
        a[q * n + j] = s * apj + c * aqj;
      }
      __syncthreads();
    }
  }
}

// rank[k] = position of eigenvalue k in ascending order (ties by index)
__device__ void ascending_rank(const double* __restrict__ a, int n, int* rank) {
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double lk = a[k * n + k];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const double lj = a[j * n + j];
      r += (lj < lk) || (lj == lk && j < k);
    }
    rank[k] = r;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kMargThreads) marg_kernel(int K, int pos, int m, double eps,
                                                            const double* __restrict__ A, const double* __restrict__ b,
                                                            double* __restrict__ A_schur, double* __restrict__ b_schur,
                                                            double* __restrict__ lin_jac, double* __restrict__ lin_res,
                                                            double* __restrict__ scratch, size_t per_cta) {
  __shared__ double red[65];
  __shared__ int pq[256];
  __shared__ double cs[256];
  __shared__ int rank[256];
  __shared__ double lam[256];
  const int n = pos - m, tid = threadIdx.x, nt = blockDim.x;
  double* ws = scratch + (size_t)blockIdx.x * per_cta;
  double* Wm = ws;                          // m*m   work / eigenvalues on the diagonal
  double* Vm = Wm + (size_t)m * m;          // m*m
  double* Inv = Vm + (size_t)m * m;         // m*m
  double* T = Inv + (size_t)m * m;          // n*m
  double* Wn = T + (size_t)n * m;           // n*n
  double* Vn = Wn + (size_t)n * n;          // n*n
  double* br = Vn + (size_t)n * n;          // n
  for (int k = blockIdx.x; k < K; k += gridDim.x) {
    const double* Ak = A + (size_t)k * pos * pos;
    const double* bk = b + (size_t)k * pos;
    __syncthreads();
    if (m > 0) {
      for (int e = tid; e < m * m; e += nt) {
        const int r = e / m, c = e % m;
        Wm[e] = 0.5 * (Ak[(size_t)r * pos + c] + Ak[(size_t)c * pos + r]);                 // :267
      }
      __syncthreads();
      block_jacobi(Wm, Vm, m, red, pq, cs);                                                // :268
      for (int e = tid; e < m; e += nt) {
        const double l = Wm[e * m + e];
        lam[e] = l > eps ? 1.0 / l : 0.0;                                                  // :272
      }
      __syncthreads();
      for (int e = tid; e < m * m; e += nt) {
        const int r = e / m, c = e % m;
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += Vm[r * m + j] * lam[j] * Vm[c * m + j];
        Inv[e] = s;
      }
      __syncthreads();
      for (int e = tid; e < n * m; e += nt) {                                              // Arm * Amm_inv
        const int r = e / m, c = e % m;
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += Ak[(size_t)(m + r) * pos + j] * Inv[j * m + c];
        T[e] = s;
      }
      __syncthreads();
    }
    for (int e = tid; e < n * n + n; e += nt) {                                            // :281-282
      if (e < n * n) {
        const int r = e / n, c = e % n;
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += T[r * m + j] * Ak[(size_t)j * pos + m + c];
        const double val = Ak[(size_t)(m + r) * pos + m + c] - s;
        Wn[e] = val;
        if (A_schur) A_schur[(size_t)k * n * n + e] = val;
      } else {
        const int r = e - n * n;
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += T[r * m + j] * bk[j];
        const double val = bk[m + r] - s;
        br[r] = val;
        if (b_schur) b_schur[(size_t)k * n + r] = val;
      }
    }
    __syncthreads();
    if (!lin_jac && !lin_res) continue;
    // SelfAdjointEigenSolver reads the lower triangle (:284)
    for (int e = tid; e < n * n; e += nt) {
      const int r = e / n, c = e % n;
      if (r < c) Wn[e] = Wn[c * n + r];
    }
    __syncthreads();
    // the mirrored copy above reads entries other threads may be writing: redo from A_schur semantics
    // (upper <- lower) is idempotent because only r < c entries are written and only r > c are read.
    block_jacobi(Wn, Vn, n, red, pq, cs);
    ascending_rank(Wn, n, rank);
    for (int e = tid; e < n; e += nt) lam[e] = Wn[e * n + e];
    __syncthreads();
    for (int e = tid; e < n * n + n; e += nt) {
      if (e < n * n) {
        if (!lin_jac) continue;
        const int kk = e / n, c = e % n;  // source eigenpair kk -> output row rank[kk]
        const double S = lam[kk] > eps ? lam[kk] : 0.0;                                    // :285
        lin_jac[(size_t)k * n * n + (size_t)rank[kk] * n + c] = sqrt(S) * Vn[c * n + kk];  // :292
      } else {
        if (!lin_res) continue;
        const int kk = e - n * n;
        const double Sinv = lam[kk] > eps ? 1.0 / lam[kk] : 0.0;                           // :286
        const double sq = sqrt(Sinv);
        double s = 0.0;
        for (int c = 0; c < n; ++c) s += (sq * Vn[c * n + kk]) * br[c];                     // :293
        lin_res[(size_t)k * n + rank[kk]] = s;
      }
    }
  }
}

}  // namespace

int viml_launch_marginalize(viml_ctx* ctx, int K, int pos, int m, double eps, const double* A, const double* b,
                            double* A_schur, double* b_schur, double* lin_jac, double* lin_res) {
  const int n = pos - m;
  const size_t per_cta = (size_t)3 * m * m + (size_t)n * m + (size_t)2 * n * n + n + 32;
  const int grid = K < 2 * ctx->sm_count ? K : 2 * ctx->sm_count;
  VIML_TRY_CUDA(ctx, ctx->scratch.reserve(per_cta * grid * sizeof(double) + 256));
  double* scratch = ctx->scratch.take<double>(per_cta * grid);
  {
    LaunchScope ls(ctx, K_MARG);
    marg_kernel<<<grid, kMargThreads, 0, ctx->stream>>>(K, pos, m, eps, A, b, A_schur, b_schur, lin_jac, lin_res,
                                                        scratch, per_cta);
  }
  VIML_TRY_CUDA(ctx, cudaGetLastError());
  return VIML_OK;
}
