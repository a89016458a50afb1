// Minimal canonical TMA bulk load: one elected thread arms the mbarrier (arrive.expect_tx) and issues one cp.async.bulk
// global -> shared completing on it; every thread waits on the barrier phase and reads the data.  Used to find out what
// `compute-sanitizer --tool racecheck` reports for the textbook pattern (profiles/sanitizer_r2.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tma_probe tma_racecheck_probe.cu
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const double* __restrict__ src, double* __restrict__ dst, int rounds) {
  __shared__ __align__(128) double buf[256];
  __shared__ uint64_t full, empty;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full)) : "memory");
#ifdef WARP_ARRIVE
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty)), "r"((int)blockDim.x / 32) : "memory");
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty)), "r"((int)blockDim.x) : "memory");
#endif
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double acc = 0.0;
  for (int it = 0; it < rounds; ++it) {
    if (threadIdx.x == 0) {
      if (it > 0) {   // every reader has arrived on `empty` for the previous round
        asm volatile("{\n\t.reg .pred P1;\n\tW0:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D0;\n\tbra W0;\n\tD0:\n\t}" ::"r"(s32(&empty)), "r"((it - 1) & 1) : "memory");
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full)), "r"(2048u) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf)),
                   "l"(src + (size_t)it * 256), "r"(2048u), "r"(s32(&full))
                   : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D1;\n\tbra W1;\n\tD1:\n\t}" ::"r"(s32(&full)), "r"(it & 1) : "memory");
    acc += buf[threadIdx.x] + buf[(threadIdx.x + 17) & 255];
#ifdef WARP_ARRIVE   // one arrival per warp, by lane 0, after the warp's reads (the pattern of schur_tma_kernel's consumers)
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty)) : "memory");
#else
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty)) : "memory");
#endif
  }
  dst[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  const int rounds = 8;
  double *src, *dst;
  cudaMalloc(&src, rounds * 256 * sizeof(double));
  cudaMalloc(&dst, 4 * 256 * sizeof(double));
  cudaMemset(src, 0, rounds * 256 * sizeof(double));
  probe<<<4, 256>>>(src, dst, rounds);
  printf("probe: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
