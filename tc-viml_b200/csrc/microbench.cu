// microbench.cu — FP64 peaks of the device (the roofline denominators MEASURED_PEAKS.json does not carry).
// Register-resident dependent chains, 8 independent accumulators per thread, enough warps to fill every SM.
#include "common.cuh"

namespace {

constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < kIters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) dmul_dadd_kernel(double* out, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < kIters; ++i) {
    x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b);
    x2 = __dadd_rn(__dmul_rn(x2, a), b); x3 = __dadd_rn(__dmul_rn(x3, a), b);
    x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
    x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace

extern "C" int viml_microbench_fp64(viml_ctx* ctx, double* dfma_tflops, double* dmul_dadd_tops) {
  if (!ctx) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8, threads = 256;
  VIML_TRY_CUDA(ctx, ctx->scratch.reserve((size_t)blocks * threads * sizeof(double)));
  double* out = ctx->scratch.take<double>((size_t)blocks * threads);
  cudaEvent_t e0, e1;
  VIML_TRY_CUDA(ctx, cudaEventCreate(&e0));
  VIML_TRY_CUDA(ctx, cudaEventCreate(&e1));
  double best[2] = {0, 0};
  for (int which = 0; which < 2; ++which)
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0, ctx->stream);
      {
        LaunchScope ls(ctx, K_MICRO);
        if (which == 0) dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, 0.999999, 1e-9);
        else dmul_dadd_kernel<<<blocks, threads, 0, ctx->stream>>>(out, 0.999999, 1e-9);
      }
      cudaEventRecord(e1, ctx->stream);
      VIML_TRY_CUDA(ctx, cudaEventSynchronize(e1));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)blocks * threads * kIters * 8 * 2;  // 2 flop / FMA, or 1 DMUL + 1 DADD
      const double rate = ops / (ms * 1e-3) / 1e12;
      if (rep > 0 && rate > best[which]) best[which] = rate;
    }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (dfma_tflops) *dfma_tflops = best[0];
  if (dmul_dadd_tops) *dmul_dadd_tops = best[1];
  return VIML_OK;
}

extern "C" int viml_microbench_dmma(viml_ctx* ctx, double* dmma_tflops) {
  if (!ctx || !dmma_tflops) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  *dmma_tflops = viml_dmma_peak_tflops(ctx);
  return VIML_OK;
}
