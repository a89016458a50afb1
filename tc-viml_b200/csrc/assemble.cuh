// assemble.cuh — fused "evaluate + loss-correct + J^T J / J^T r" kernel: one CTA per sliding window.
// Included by linearize_kernels.cu after eval_point / eval_line.
//
// What the reference does per window (marginalization_factor.cpp:3-69, :141-172): evaluate every factor into
// heap-allocated Jacobian blocks, then 4 pthreads each scatter J_i^T J_j into a private dense pos x pos A.
// Here the Jacobians never leave the SM:
//   P0  deterministic counting sorts of the window's factors (warp match_any ranks + per-slab counts):
//         by pose pair (lo,hi) -> record position (all factors that hit the same H blocks are contiguous),
//         by feature          -> per-landmark factor lists,   lines by frame.
//   P1  one thread per factor: residual + Jacobian in registers (eval_point / eval_line), written as
//       row-pair records (J[0][c], J[1][c]) into XOR-swizzled shared memory at the sorted position.
//   P2a pose part of H: every target block is owned by one TEAM of lanes (32/16/8/4 lanes by block type);
//       lanes stride the block's factor ranges, accumulate the full block in registers, then a shuffle
//       reduce-scatter leaves each lane with a slice that it writes straight to HBM (block + mirror).
//   P2b landmark part: one thread per feature walks its factor list and writes the whole strip row
//       H_lp[l][:] (structural zeros included), H_ll[l], b_l[l].
// A window is processed in PARTS of <= NF point / NL line factors (any split of the factor list is valid: J^T J is
// a sum over factors).  Part 0 writes every output entry of the window ("=", structural zeros included), later
// parts accumulate ("+=" as one RED per entry, same CTA, ordered by __syncthreads), so any window size is handled
// by the same kernel; an EuRoC-shaped window (~560 + 110 factors) is a single part.  Every entry is produced by
// exactly one thread per part: no contended atomics, no memset, bit-reproducible run to run.
// (Measured, round 1: 336-factor parts with two 256-thread CTAs per SM were 1.5x SLOWER than one 512-thread CTA per
// SM: +32 % instructions for the per-part sorts/reductions and instruction-cache misses outweighed the overlap.)
#pragma once

namespace fused {

constexpr int AT = 512;          // threads per CTA (16 warps), 1 CTA per SM
constexpr int NF = 704;          // point factors per PART (a window is processed in ceil(nf/NF) accumulating parts)
constexpr int NL = 160;          // line factors per part
constexpr int PMAX = 12;         // max poses per window
constexpr int FMAX = 160;        // max features per window
constexpr int KA = PMAX * PMAX;  // pair keys
constexpr int SLABS = (NF + 31) / 32;
constexpr int LSLABS = (NL + 31) / 32;
constexpr int RECW = 17;         // double2 per point record: 0..5 lo, 6..8 hi_rot, 9..14 ex, 15 r, 16 d (odd stride: conflict-free)
constexpr int LRECW = 7;         // double2 per line record: 0..5 pose, 6 r
constexpr int WSLOTS = 64;       // windows per CTA whose CSR offsets are staged in shared memory

struct Smem {
  double2 rec[NF * RECW];
  union {
    double2 lrec[NL * LRECW];
    struct {
      uint16_t cntA[SLABS * KA];
      uint16_t cntB[SLABS * FMAX];
      uint16_t cntC[LSLABS * PMAX];
    } cnt;
  } u;
  double cache[PMAX * kPoseCache + kExCache];
  uint32_t ridx[NF];
  uint16_t fperm[NF];
  uint16_t hperm[NF];  // record positions ordered by hi pose (flattened "hi-role" ranges)
  uint16_t baseA[KA + 1], totA[KA], baseB[FMAX + 1], totB[FMAX], baseC[PMAX + 1], totC[PMAX], baseD[PMAX + 1], offD[KA], pairs[KA];
  int npairs, next_feature;
  int woff[4 * WSLOTS];  // this CTA's windows: {p0, p1, l0, l1} per slot (filled once per launch)
};

__constant__ unsigned char c_sym_r[21] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5};
__constant__ unsigned char c_sym_c[21] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5};

__device__ __forceinline__ void prefetch_l2(const void* base, size_t bytes, int tid, int nthreads) {
  const char* p = reinterpret_cast<const char*>(base);
  for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)nthreads * 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
}

__device__ __forceinline__ double2 ldrec(const double2* __restrict__ rec, int pos, int k) {
  return rec[pos * RECW + k];  // k is a compile-time constant at every call site: one base register per record
}
__device__ __forceinline__ void strec(double2* __restrict__ rec, int pos, int k, double x, double y) {
  rec[pos * RECW + k] = make_double2(x, y);
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// reduce-scatter step: N live values -> N/2, lanes with bit w keep the upper half
template <int N>
__device__ __forceinline__ void rs_step(double* a, bool up, int w, unsigned mask) {
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const double send = up ? a[k] : a[k + N / 2];
    const double keep = up ? a[k + N / 2] : a[k];
    a[k] = keep + __shfl_xor_sync(mask, send, w);
  }
}
template <int N>
__device__ __forceinline__ void bf_step(double* a, int w, unsigned mask) {
#pragma unroll
  for (int k = 0; k < N; ++k) a[k] += __shfl_xor_sync(mask, a[k], w);
}

// acc(6x6) += X^T Y with X, Y given as row pairs
__device__ __forceinline__ void acc_full(double* acc, const double2* X, const double2* Y) {
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[r * 6 + c] = fma(X[r].x, Y[c].x, fma(X[r].y, Y[c].y, acc[r * 6 + c]));
}
// acc[0..20] += upper(X^T X), acc[21..26] += X^T r
__device__ __forceinline__ void acc_sym(double* acc, const double2* X, double2 r) {
  int e = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int c = a; c < 6; ++c, ++e) acc[e] = fma(X[a].x, X[c].x, fma(X[a].y, X[c].y, acc[e]));
#pragma unroll
  for (int a = 0; a < 6; ++a) acc[21 + a] = fma(X[a].x, r.x, fma(X[a].y, r.y, acc[21 + a]));
}

__device__ __forceinline__ void load_lo(const double2* rec, int pos, double2* X) {
#pragma unroll
  for (int k = 0; k < 6; ++k) X[k] = ldrec(rec, pos, k);
}
__device__ __forceinline__ void load_hi(const double2* rec, int pos, double2* X) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double2 v = ldrec(rec, pos, k);
    X[k] = make_double2(-v.x, -v.y);          // hi translation block == -lo translation block
    X[3 + k] = ldrec(rec, pos, 6 + k);
  }
}
__device__ __forceinline__ void load_ex(const double2* rec, int pos, double2* X) {
#pragma unroll
  for (int k = 0; k < 6; ++k) X[k] = ldrec(rec, pos, 9 + k);
}

// "=" for the first part of a window, "+=" for the following ones.  The add is a fire-and-forget RED (no load on
// the critical path); parts of a window run on the same CTA and are separated by __syncthreads, each address gets
// at most one add per part, so the result is still deterministic.
__device__ __forceinline__ void put(double* __restrict__ p, double v, bool accum) {
  if (accum) atomicAdd(p, v);
  else *p = v;
}
__device__ __forceinline__ void put2(double2* __restrict__ p, double x, double y, bool accum) {
  if (accum) {
    atomicAdd(&p->x, x);
    atomicAdd(&p->y, y);
  } else {
    *p = make_double2(x, y);
  }
}

// write the sym(21)+b(6) slice a lane holds after the reduction: entries e0..e0+6 of block (p,p)
__device__ __forceinline__ void write_sym_slice(const double* a, int e0, double* __restrict__ H, double* __restrict__ bp,
                                                int D, int p, bool accum) {
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int e = e0 + k;
    if (e < 21) {
      const int r = c_sym_r[e], c = c_sym_c[e];
      put(&H[(size_t)(6 * p + r) * D + 6 * p + c], a[k], accum);
      if (r != c) put(&H[(size_t)(6 * p + c) * D + 6 * p + r], a[k], accum);
    } else if (e < 27) {
      put(&bp[6 * p + e - 21], a[k], accum);
    }
  }
}

// P0 sorts on a workspace WS (the fused kernel's Smem or the plan kernel's PlanSmem: same member names).  The count
// tables S.u.cnt must be zero on entry (and a barrier passed); on return every thread holds the record positions of
// its two point factors (posA) and of its line factor (posC), S.fperm / S.hperm / base* / tot* are filled and a
// barrier is still needed before they are read.
template <bool BIGF, bool WITH_RIDX, class WS>
__device__ __forceinline__ void p0_sort(WS& S, int tid, int P, int F, int nf, int nl, uint32_t idx0, uint32_t idx1, int frame,
                                        uint32_t (&fidx)[2], int (&posA)[2], int& lframe, int& posC) {
  const int lane = tid & 31, warp = tid >> 5;
  const int nkeyA = P * P;
  int keyA[2], rankA[2], rankB[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int f = tid + r * AT;
    const bool valid = f < nf;
    fidx[r] = r == 0 ? idx0 : idx1;
    const int i = fidx[r] & 0xff, j = (fidx[r] >> 8) & 0xff, l = fidx[r] >> 16;
    const int lo = min(i, j), hi = max(i, j);
    keyA[r] = valid ? lo * P + hi : 0xffff;
    const int keyB = (valid && !BIGF) ? l : 0xffff;
    if (r * AT < nf) {  // warp-uniform: some lane of this warp may be valid
      const unsigned ma = __match_any_sync(0xffffffffu, keyA[r]);
      const unsigned mb = __match_any_sync(0xffffffffu, keyB);
      const unsigned lt = (1u << lane) - 1u;
      rankA[r] = __popc(ma & lt);
      rankB[r] = __popc(mb & lt);
      const int slab = warp + r * (AT / 32);
      if (valid && rankA[r] == 0) S.u.cnt.cntA[slab * KA + keyA[r]] = (uint16_t)__popc(ma);
      if (!BIGF && valid && rankB[r] == 0) S.u.cnt.cntB[slab * FMAX + keyB] = (uint16_t)__popc(mb);
    }
  }
  // line factor t is handled by thread AT-1-t: the second round of point factors uses the LOW threads, so no
  // thread gets two point factors and a line factor on its critical path
  const int lt_ = AT - 1 - tid;
  int rankC = 0;
  lframe = 0;
  if (warp >= AT / 32 - LSLABS) {  // line factors by frame (whole warps take part in match_any)
    const bool valid = lt_ < nl;
    lframe = valid ? frame : 0xffff;
    const unsigned mc = __match_any_sync(0xffffffffu, lframe);
    rankC = __popc(mc & ((1u << lane) - 1u));
    if (valid && rankC == 0) S.u.cnt.cntC[(AT / 32 - 1 - warp) * PMAX + lframe] = (uint16_t)__popc(mc);
  }
  __syncthreads();
  // exclusive prefix over slabs for every key (thread per key); loads are issued together, the running sum
  // stays in registers (the serial load->store chain of a rolled loop cost ~1 K cycles per window)
  const int Fs = BIGF ? 0 : F;  // feature keys handled in shared memory
  for (int e = tid; e < nkeyA + Fs + P; e += AT) {
    uint16_t* col;
    int stride, ns;
    uint16_t* tot;
    if (e < nkeyA) col = S.u.cnt.cntA + e, stride = KA, ns = SLABS, tot = S.totA + e;
    else if (e < nkeyA + Fs) col = S.u.cnt.cntB + (e - nkeyA), stride = FMAX, ns = SLABS, tot = S.totB + (e - nkeyA);
    else col = S.u.cnt.cntC + (e - nkeyA - Fs), stride = PMAX, ns = LSLABS, tot = S.totC + (e - nkeyA - Fs);
    int c[SLABS];
#pragma unroll
    for (int q = 0; q < SLABS; ++q) c[q] = q < ns ? col[q * stride] : 0;
    int run = 0;
#pragma unroll
    for (int q = 0; q < SLABS; ++q) {
      if (q < ns) col[q * stride] = (uint16_t)run;
      run += c[q];
    }
    *tot = (uint16_t)run;
  }
  __syncthreads();
  // exclusive scans over keys: warp 0 -> baseA (+ non-empty pair list), warp 1 -> baseB, warp 2 -> baseC
  if (warp < 3) {
    const uint16_t* tot = warp == 0 ? S.totA : (warp == 1 ? S.totB : S.totC);
    uint16_t* base = warp == 0 ? S.baseA : (warp == 1 ? S.baseB : S.baseC);
    const int n = warp == 0 ? nkeyA : (warp == 1 ? Fs : P);
    int carry = 0, npairs = 0;
    for (int b0 = 0; b0 < n; b0 += 32) {
      const int k = b0 + lane;
      const int v = k < n ? tot[k] : 0;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (k < n) base[k] = (uint16_t)(carry + inc - v);
      carry += __shfl_sync(0xffffffffu, inc, 31);
      if (warp == 0) {
        const unsigned nz = __ballot_sync(0xffffffffu, v > 0);
        if (v > 0) S.pairs[npairs + __popc(nz & ((1u << lane) - 1u))] = (uint16_t)k;
        npairs += __popc(nz);
      }
    }
    if (lane == 0) {
      base[n] = (uint16_t)carry;
      if (warp == 0) S.npairs = npairs;
    }
  } else if (warp == 3) {
    // records ordered by (hi, lo) WITHOUT another sort: inside one (lo,hi) group the pair-sorted positions are
    // already contiguous, so  posD = baseD[hi] + sum_{lo' < lo} tot(lo',hi) + (posA - baseA[lo,hi]).
    // lane = hi: column sums of totA, exclusive scan over hi, then the per-(lo,hi) offsets.
    const int hi = lane;
    int col = 0;
    if (hi < P)
      for (int lo = 0; lo < hi; ++lo) col += S.totA[lo * P + hi];
    int inc = col;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (hi < P) {
      int run = inc - col;
      S.baseD[hi] = (uint16_t)run;
      for (int lo = 0; lo < hi; ++lo) {
        S.offD[lo * P + hi] = (uint16_t)run;
        run += S.totA[lo * P + hi];
      }
    }
    if (lane == 31) S.baseD[P] = (uint16_t)inc;  // lanes >= P contribute 0: inclusive sum at lane 31 is the total
    static_assert(PMAX <= 31, "hi scan uses one warp");
    static_assert(PMAX + 1 < AT / 32, "P2a: one warp per pose, one for the extrinsic block, at least one helper");
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int f = tid + r * AT;
    posA[r] = 0;
    if (f < nf) {
      const int slab = warp + r * (AT / 32);
      const int l = fidx[r] >> 16;
      posA[r] = S.baseA[keyA[r]] + S.u.cnt.cntA[slab * KA + keyA[r]] + rankA[r];
      if (WITH_RIDX) S.ridx[posA[r]] = fidx[r];
      if (!BIGF) S.fperm[S.baseB[l] + S.u.cnt.cntB[slab * FMAX + l] + rankB[r]] = (uint16_t)posA[r];
      S.hperm[S.offD[keyA[r]] + posA[r] - S.baseA[keyA[r]]] = (uint16_t)posA[r];
    }
  }
  posC = 0;
  if (lt_ < nl) posC = S.baseC[lframe] + S.u.cnt.cntC[(AT / 32 - 1 - warp) * PMAX + lframe] + rankC;
}

// Layout of one window's sort plan in HBM (uint16 units), written by plan_kernel for single-part windows.
constexpr int kPlanPosA = 0, kPlanFperm = NF, kPlanHperm = 2 * NF, kPlanPosC = 3 * NF, kPlanTab = 3 * NF + NL;
constexpr int kPlanTabN = (KA + 1) + KA + (FMAX + 1) + (PMAX + 1) + (PMAX + 1);   // baseA, totA, baseB, baseC, baseD
constexpr int kPlanStride = (kPlanTab + kPlanTabN + 7) / 8 * 8;

struct PlanSmem {
  struct {
    struct {
      uint16_t cntA[SLABS * KA];
      uint16_t cntB[SLABS * FMAX];
      uint16_t cntC[LSLABS * PMAX];
    } cnt;
  } u;
  uint16_t fperm[NF], hperm[NF];
  uint16_t baseA[KA + 1], totA[KA], baseB[FMAX + 1], totB[FMAX], baseC[PMAX + 1], totC[PMAX], baseD[PMAX + 1], offD[KA], pairs[KA];
  uint32_t ridx[1];
  int npairs;
};

// table entry e of the plan <-> the workspace arrays
template <class WS>
__device__ __forceinline__ uint16_t* plan_tab(WS& S, int e) {
  if (e < KA + 1) return S.baseA + e;
  e -= KA + 1;
  if (e < KA) return S.totA + e;
  e -= KA;
  if (e < FMAX + 1) return S.baseB + e;
  e -= FMAX + 1;
  if (e < PMAX + 1) return S.baseC + e;
  e -= PMAX + 1;
  return S.baseD + e;
}

// The sorts of a window depend on its factor indices only, and inside the persistent fused kernel (one CTA per SM)
// their ~10 barriers are pure latency.  This kernel runs them ahead for every single-part window with several CTAs
// per SM; the fused kernel then just copies ~5.5 KB of positions and tables.
__global__ void __launch_bounds__(AT, 4) plan_kernel(LinearizeArgs A, uint16_t* __restrict__ plan) {
  __shared__ PlanSmem S;
  const int tid = threadIdx.x, w = blockIdx.x;
  const int a0 = A.pf_window_offset[w], nf = A.pf_window_offset[w + 1] - a0;
  const int b0 = A.NL > 0 ? A.lf_window_offset[w] : 0, nl = A.NL > 0 ? A.lf_window_offset[w + 1] - b0 : 0;
  if (nf > NF || nl > NL) return;   // multi-part window: sorted part by part inside the fused kernel
  uint32_t* z = reinterpret_cast<uint32_t*>(&S.u.cnt);
  for (int e = tid; e < (int)(sizeof(S.u.cnt) / 4); e += AT) z[e] = 0u;
  const uint32_t idx0 = tid < nf ? A.pf_idx[a0 + tid] : 0u, idx1 = tid + AT < nf ? A.pf_idx[a0 + tid + AT] : 0u;
  const int frame = AT - 1 - tid < nl ? A.lf_frame[b0 + AT - 1 - tid] : 0xffff;
  __syncthreads();
  uint32_t fidx[2];
  int posA[2], lframe, posC;
  p0_sort<false, false>(S, tid, A.P, A.F, nf, nl, idx0, idx1, frame, fidx, posA, lframe, posC);
  __syncthreads();
  uint16_t* __restrict__ out = plan + (size_t)w * kPlanStride;
  if (tid < nf) out[kPlanPosA + tid] = (uint16_t)posA[0];
  if (tid + AT < nf) out[kPlanPosA + tid + AT] = (uint16_t)posA[1];
  if (AT - 1 - tid < nl) out[kPlanPosC + AT - 1 - tid] = (uint16_t)posC;
  for (int e = tid; e < nf; e += AT) out[kPlanFperm + e] = S.fperm[e], out[kPlanHperm + e] = S.hperm[e];
  for (int e = tid; e < kPlanTabN; e += AT) out[kPlanTab + e] = *plan_tab(S, e);
}

// BIGF: more features per window than the shared-memory feature tables hold (F > FMAX, e.g. the 2000-landmark
// marginalisation stress): the per-feature sort and P2b are skipped; every factor adds its landmark terms with
// RED.ADD.F64 onto rows the launcher zero-filled (their summation order is then not reproducible bit for bit).
template <bool MODE_A, bool BIGF>
__global__ void __launch_bounds__(AT, 1) assemble_kernel(LinearizeArgs A, int w_begin, int w_end) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = A.P, F = A.F, D = A.D, E = A.P;
  const int cstride = P * kPoseCache + kExCache;

  // stage this CTA's CSR offsets once so that per-window address generation never waits on HBM
  // (the launcher keeps (w_end - w_begin) <= WSLOTS * gridDim.x)
  for (int e = tid; e < WSLOTS; e += AT) {
    const int w = w_begin + blockIdx.x + e * gridDim.x;
    if (w < w_end) {
      S.woff[4 * e] = A.pf_window_offset[w];
      S.woff[4 * e + 1] = A.pf_window_offset[w + 1];
      S.woff[4 * e + 2] = A.NL > 0 ? A.lf_window_offset[w] : 0;
      S.woff[4 * e + 3] = A.NL > 0 ? A.lf_window_offset[w + 1] : 0;
    }
  }
  __syncthreads();
  // register prefetch of the next window's sort keys and pose cache (consumed in its P0)
  uint32_t nx_idx0 = 0u, nx_idx1 = 0u;
  int nx_frame = 0xffff;
  double nx_c0 = 0.0, nx_c1 = 0.0, nx_c2 = 0.0;
  static_assert(PMAX * kPoseCache + kExCache <= 3 * AT, "pose cache prefetch registers");
  // factor / line ranges of part `part` of the window in slot `sl`
  auto seg = [&](int sl, int part, int& q0, int& qn, int& m0, int& mn) {
    const int a0 = S.woff[4 * sl], an = S.woff[4 * sl + 1] - a0, b0 = S.woff[4 * sl + 2], bn = S.woff[4 * sl + 3] - b0;
    const int np = max(1, max((an + NF - 1) / NF, (bn + NL - 1) / NL));
    if (np == 1) {   // the common case: no 64-bit divisions on the per-window path
      q0 = a0, qn = an, m0 = b0, mn = bn;
      return 1;
    }
    q0 = a0 + (int)((int64_t)an * part / np), qn = a0 + (int)((int64_t)an * (part + 1) / np) - q0;
    m0 = b0 + (int)((int64_t)bn * part / np), mn = b0 + (int)((int64_t)bn * (part + 1) / np) - m0;
    return np;
  };
  auto fetch_next = [&](int wn, int sl, int part) {
    int q0, qn, m0, mn;
    seg(sl, part, q0, qn, m0, mn);
    nx_idx0 = tid < qn ? A.pf_idx[q0 + tid] : 0u;
    nx_idx1 = tid + AT < qn ? A.pf_idx[q0 + tid + AT] : 0u;
    nx_frame = AT - 1 - tid < mn ? A.lf_frame[m0 + AT - 1 - tid] : 0xffff;
    const double* __restrict__ gc = A.cache + (size_t)wn * cstride;
    nx_c0 = tid < cstride ? gc[tid] : 0.0;
    nx_c1 = tid + AT < cstride ? gc[tid + AT] : 0.0;
    nx_c2 = tid + 2 * AT < cstride ? gc[tid + 2 * AT] : 0.0;
  };
  if (w_begin + (int)blockIdx.x < w_end) fetch_next(w_begin + blockIdx.x, 0, 0);

  long long tph[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
#define VIML_TICK(k) do { if (A.dbg && tid == 0) { const long long now_ = clock64(); tph[k] += now_ - tlast; tlast = now_; } } while (0)
  int slot = 0;
  for (int w = w_begin + blockIdx.x; w < w_end; w += gridDim.x, ++slot) {
    const bool has_next = w + (int)gridDim.x < w_end;
    if (has_next) {  // pull the next window's per-factor inputs into L2 while this one computes
      const int wn = w + gridDim.x;
      const int q0 = S.woff[4 * slot + 4], q1 = S.woff[4 * slot + 5], m0 = S.woff[4 * slot + 6], m1 = S.woff[4 * slot + 7];
      prefetch_l2(A.pf_obs + (size_t)q0 * 4, (size_t)(q1 - q0) * 32, tid, AT);
      prefetch_l2(A.inv_depth + (size_t)wn * F, (size_t)F * 8, tid, AT);
      if (A.plan) prefetch_l2(A.plan + (size_t)wn * kPlanStride, (size_t)kPlanStride * 2, tid, AT);
      if (A.pf_pts_i_z) prefetch_l2(A.pf_pts_i_z + q0, (size_t)(q1 - q0) * 8, tid, AT);
      if (A.NL > 0)
        for (int c = 0; c < 9; ++c) prefetch_l2(A.lf_geom + (size_t)c * A.NL_stride + m0, (size_t)(m1 - m0) * 8, tid, AT);
    }
    double* __restrict__ Hpp = A.out.H_pp + (size_t)w * D * D;
    double* __restrict__ Hlp = A.out.H_lp + (size_t)w * F * D;
    double* __restrict__ Hll = A.out.H_ll + (size_t)w * F;
    double* __restrict__ bp = A.out.b_p + (size_t)w * D;
    double* __restrict__ bl = A.out.b_l + (size_t)w * F;
    int p0, nf, l0, nl;
    const int nparts = seg(slot, 0, p0, nf, l0, nl);
    for (int part = 0; part < nparts; ++part) {
    seg(slot, part, p0, nf, l0, nl);
    const bool accum = part > 0;  // later parts add onto what part 0 wrote
    // ------------------------------------------------------------------ P0: cache + sorts (or the precomputed plan)
    const bool planned = !BIGF && A.plan != nullptr && nparts == 1;
    {
      if (tid < cstride) S.cache[tid] = nx_c0;
      if (tid + AT < cstride) S.cache[tid + AT] = nx_c1;
      if (tid + 2 * AT < cstride) S.cache[tid + 2 * AT] = nx_c2;
      if (tid == 0) S.next_feature = 0;
    }
    uint32_t fidx[2];
    int posA[2], lframe, posC;
    const int lt_ = AT - 1 - tid;   // line factor t is handled by thread AT-1-t
    if (planned) {
      const uint16_t* __restrict__ pl = A.plan + (size_t)w * kPlanStride;
      fidx[0] = nx_idx0, fidx[1] = nx_idx1;
      lframe = lt_ < nl ? nx_frame : 0xffff;
      // all loads first (one L2 round trip; the row was prefetched into L2 during the previous window)
      static_assert(NF / 2 <= AT && kPlanTabN <= AT, "plan copy: one element per thread");
      const uint32_t* __restrict__ fp = reinterpret_cast<const uint32_t*>(pl + kPlanFperm);
      const uint32_t* __restrict__ hp = reinterpret_cast<const uint32_t*>(pl + kPlanHperm);
      const int pa0 = tid < nf ? pl[kPlanPosA + tid] : 0, pa1 = tid + AT < nf ? pl[kPlanPosA + tid + AT] : 0;
      const int pc = lt_ < nl ? pl[kPlanPosC + lt_] : 0;
      const bool cp = tid < (nf + 1) / 2;
      const uint32_t wf = cp ? fp[tid] : 0u, wh = cp ? hp[tid] : 0u;
      const uint16_t tb = tid < kPlanTabN ? pl[kPlanTab + tid] : (uint16_t)0;
      posA[0] = pa0, posA[1] = pa1, posC = pc;
      if (cp) reinterpret_cast<uint32_t*>(S.fperm)[tid] = wf, reinterpret_cast<uint32_t*>(S.hperm)[tid] = wh;
      if (tid < kPlanTabN) *plan_tab(S, tid) = tb;
      if (tid < nf) S.ridx[pa0] = fidx[0];
      if (tid + AT < nf) S.ridx[pa1] = fidx[1];
    } else {
      uint32_t* z = reinterpret_cast<uint32_t*>(&S.u.cnt);
      for (int e = tid; e < (int)(sizeof(S.u.cnt) / 4); e += AT) z[e] = 0u;
      __syncthreads();
      p0_sort<BIGF, true>(S, tid, P, F, nf, nl, nx_idx0, nx_idx1, nx_frame, fidx, posA, lframe, posC);
    }
    VIML_TICK(0);
    __syncthreads();  // count tables are dead from here on (lrec aliases them)
    VIML_TICK(1);
    // ------------------------------------------------------------------ P1: evaluate, write records
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int f = tid + r * AT;
      if (f >= nf) continue;
      const int64_t k = (int64_t)p0 + f;
      const int i = fidx[r] & 0xff, j = (fidx[r] >> 8) & 0xff, l = fidx[r] >> 16;
      const double4 ob = reinterpret_cast<const double4*>(A.pf_obs)[k];
      const double piz = A.pf_pts_i_z ? A.pf_pts_i_z[k] : 1.0;
      const double lam = A.inv_depth[(size_t)w * F + l];
      PointJac J;
      eval_point(A, S.cache, i, j, lam, ob.x, ob.y, piz, ob.z, ob.w, J);
      if (MODE_A) {
        if (A.out.pf_residual) reinterpret_cast<double2*>(A.out.pf_residual)[k] = make_double2(J.r[0], J.r[1]);
        if (A.out.pf_jac_pose_i) store_jac7(A.out.pf_jac_pose_i + 14 * k, J.a);
        if (A.out.pf_jac_pose_j) store_jac7(A.out.pf_jac_pose_j + 14 * k, J.b);
        if (A.out.pf_jac_ex) store_jac7(A.out.pf_jac_ex + 14 * k, J.c);
        if (A.out.pf_jac_feat) reinterpret_cast<double2*>(A.out.pf_jac_feat)[k] = make_double2(J.d[0], J.d[1]);
      }
      const int pos = posA[r];
      const bool sw = i > j;  // lo block is pose j
#pragma unroll
      for (int c = 0; c < 6; ++c) strec(S.rec, pos, c, sw ? J.b[0][c] : J.a[0][c], sw ? J.b[1][c] : J.a[1][c]);
#pragma unroll
      for (int c = 0; c < 3; ++c) strec(S.rec, pos, 6 + c, sw ? J.a[0][3 + c] : J.b[0][3 + c], sw ? J.a[1][3 + c] : J.b[1][3 + c]);
#pragma unroll
      for (int c = 0; c < 6; ++c) strec(S.rec, pos, 9 + c, J.c[0][c], J.c[1][c]);
      strec(S.rec, pos, 15, J.r[0], J.r[1]);
      strec(S.rec, pos, 16, J.d[0], J.d[1]);
      if (BIGF) {  // landmark row terms straight from registers (rows were zero-filled by the launcher)
        double* row = Hlp + (size_t)l * D;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          atomicAdd(row + 6 * i + c, J.d[0] * J.a[0][c] + J.d[1] * J.a[1][c]);
          atomicAdd(row + 6 * E + c, J.d[0] * J.c[0][c] + J.d[1] * J.c[1][c]);
          row[6 * j + c] = J.d[0] * J.b[0][c] + J.d[1] * J.b[1][c];   // unique per (landmark, frame)
        }
        atomicAdd(Hll + l, J.d[0] * J.d[0] + J.d[1] * J.d[1]);
        atomicAdd(bl + l, J.d[0] * J.r[0] + J.d[1] * J.r[1]);
      }
    }
    if (lt_ < nl) {
      const int64_t k = (int64_t)l0 + lt_;
      double g9[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) g9[c] = A.lf_geom[(size_t)c * A.NL_stride + k];
      LineJac J;
      eval_line(A, S.cache, lframe, g9, J);
      if (MODE_A) {
        if (A.out.lf_residual) reinterpret_cast<double2*>(A.out.lf_residual)[k] = make_double2(J.r[0], J.r[1]);
        if (A.out.lf_jac_pose) store_jac7(A.out.lf_jac_pose + 14 * k, J.a);
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) S.u.lrec[posC * LRECW + c] = make_double2(J.a[0][c], J.a[1][c]);
      S.u.lrec[posC * LRECW + 6] = make_double2(J.r[0], J.r[1]);
    }
    VIML_TICK(2);
    __syncthreads();
    VIML_TICK(3);
    // ------------------------------------------------------------------ P2a: pose blocks on the FP64 tensor pipe
    // Every 6x6 block of H_pp is a sum over factors of X^T Y with X, Y 2x6 Jacobian blocks: two factors make one
    // K = 4 step of mma.m8n8k4.f64 (C[m][n] += sum_k A[m][k] B[k][n], A[m][k] = X[k][m], B[k][n] = Y[k][n]); lane
    // (g = lane >> 2, k = lane & 3) holds column g of row k & 1 of factor k >> 1 for both operand layouts, so one
    // 8-byte shared-memory load per lane is a whole fragment and the same register serves X as A and as B.  Column
    // 6 of a fragment carries the residual, which makes column 6 of X^T [X | r] the gradient block for free.
    //   warp p < P : pose p.  lo list (pair-sorted, contiguous): (p,p) += X^T X, (p,ex) += X^T Z, (p,hi) += X^T Y per
    //                (p,hi) segment; hi list (through hperm): (p,p) += Y^T Y, (p,ex) += Y^T Z; line factors of frame
    //                p: (p,p) += L^T L.  One warp owns all of pose p's blocks: single writer, no reduction.
    //   warp P     : (ex,ex) + b_ex over every point factor.
    //   warps > P  : structural zeros of untouched pair blocks, then straight on to P2b.
    const long long t2a0 = clock64();
    {
      const int g = lane >> 2, krow = lane & 1, kf = (lane >> 1) & 1;
      const int c2 = 2 * (lane & 3);                       // accumulator columns c2, c2 + 1 of row g
      // double offsets inside a record for the fragment kinds (column 7 is always zero)
      const int offX = (g < 6 ? g : 15) * 2 + krow;                                 // lo block | r
      const int offZ = (g < 6 ? 9 + g : 15) * 2 + krow;                             // ex block | r
      const int offY = (g < 3 ? g : (g < 6 ? g + 3 : 15)) * 2 + krow;               // hi block (-lo translation, hi rotation) | r
      const double sgnY = g < 3 ? -1.0 : 1.0;
      const bool col_ok = g < 7;
      const double* __restrict__ recd = reinterpret_cast<const double*>(S.rec);
      auto store_block = [&](int rb, int cb, double v0, double v1, bool mirror) {
        if (g < 6 && c2 < 6) {
          put(&Hpp[(size_t)(6 * rb + g) * D + 6 * cb + c2], v0, accum);
          put(&Hpp[(size_t)(6 * rb + g) * D + 6 * cb + c2 + 1], v1, accum);
          if (mirror) {
            put(&Hpp[(size_t)(6 * cb + c2) * D + 6 * rb + g], v0, accum);
            put(&Hpp[(size_t)(6 * cb + c2 + 1) * D + 6 * rb + g], v1, accum);
          }
        }
      };
      // Two K-steps (four factors) per iteration on two accumulator sets: the loads of both steps are issued before
      // the first mma, and consecutive mma on one block do not wait for each other.
      if (warp < P) {
        const int p = warp;
        double pp0 = 0.0, pp1 = 0.0, pe0 = 0.0, pe1 = 0.0, qq0 = 0.0, qq1 = 0.0, qe0 = 0.0, qe1 = 0.0;
        for (int hi = p + 1; hi < P; ++hi) {
          const int key = p * P + hi, n = S.totA[key];
          if (n == 0) continue;
          const int b0 = S.baseA[key], end = b0 + n;
          double ph0 = 0.0, ph1 = 0.0, qh0 = 0.0, qh1 = 0.0;
          for (int q = b0; q < end; q += 4) {
            const int posa = q + kf, posb = q + 2 + kf;
            const bool oka = col_ok && posa < end, okb = col_ok && posb < end;
            const double* ra = recd + (size_t)posa * (2 * RECW);
            const double* rb = recd + (size_t)posb * (2 * RECW);
            const double xa = oka ? ra[offX] : 0.0, za = oka ? ra[offZ] : 0.0, ya = oka ? sgnY * ra[offY] : 0.0;
            const double xb = okb ? rb[offX] : 0.0, zb = okb ? rb[offZ] : 0.0, yb = okb ? sgnY * rb[offY] : 0.0;
            dmma(pp0, pp1, xa, xa);
            dmma(pe0, pe1, xa, za);
            dmma(ph0, ph1, xa, ya);
            dmma(qq0, qq1, xb, xb);
            dmma(qe0, qe1, xb, zb);
            dmma(qh0, qh1, xb, yb);
          }
          store_block(p, hi, ph0 + qh0, ph1 + qh1, true);
        }
        for (int q = S.baseD[p], end = S.baseD[p + 1]; q < end; q += 4) {
          const bool oka = col_ok && q + kf < end, okb = col_ok && q + 2 + kf < end;
          const int posa = oka ? S.hperm[q + kf] : 0, posb = okb ? S.hperm[q + 2 + kf] : 0;
          const double* ra = recd + (size_t)posa * (2 * RECW);
          const double* rb = recd + (size_t)posb * (2 * RECW);
          const double ya = oka ? sgnY * ra[offY] : 0.0, za = oka ? ra[offZ] : 0.0;
          const double yb = okb ? sgnY * rb[offY] : 0.0, zb = okb ? rb[offZ] : 0.0;
          dmma(pp0, pp1, ya, ya);
          dmma(pe0, pe1, ya, za);
          dmma(qq0, qq1, yb, yb);
          dmma(qe0, qe1, yb, zb);
        }
        {
          const double* lrd = reinterpret_cast<const double*>(S.u.lrec);
          for (int q = S.baseC[p], end = S.baseC[p + 1]; q < end; q += 4) {
            const bool oka = col_ok && q + kf < end, okb = col_ok && q + 2 + kf < end;
            const double la = oka ? lrd[(size_t)(q + kf) * (2 * LRECW) + 2 * g + krow] : 0.0;   // columns 0..5 = J, 6 = r
            const double lb = okb ? lrd[(size_t)(q + 2 + kf) * (2 * LRECW) + 2 * g + krow] : 0.0;
            dmma(pp0, pp1, la, la);
            dmma(qq0, qq1, lb, lb);
          }
        }
        pp0 += qq0, pp1 += qq1, pe0 += qe0, pe1 += qe1;
        store_block(p, p, pp0, pp1, false);                 // the tile holds both triangles (bitwise symmetric)
        if (g < 6 && c2 == 6) put(&bp[6 * p + g], pp0, accum);
        store_block(p, E, pe0, pe1, true);
      } else if (warp == P) {
        double e0[4] = {0, 0, 0, 0}, e1[4] = {0, 0, 0, 0};
        for (int q = 0; q < nf; q += 8) {
          double z[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int pos = q + 2 * u + kf;
            z[u] = (col_ok && pos < nf) ? recd[(size_t)pos * (2 * RECW) + offZ] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma(e0[u], e1[u], z[u], z[u]);
        }
        const double ee0 = (e0[0] + e0[1]) + (e0[2] + e0[3]), ee1 = (e1[0] + e1[1]) + (e1[2] + e1[3]);
        store_block(E, E, ee0, ee1, false);
        if (g < 6 && c2 == 6) put(&bp[6 * E + g], ee0, accum);
      } else if (!accum) {
        // pose-pair blocks that no factor touches are structural zeros
        const int nh = AT / 32 - (P + 1), h = warp - (P + 1);
        int e = 0;
        for (int lo = 0; lo < P; ++lo)
          for (int hi = lo + 1; hi < P; ++hi)
            if (S.totA[lo * P + hi] == 0 && (e++ % nh) == h)
              for (int q = lane; q < 36; q += 32) {
                const int r = q / 6, c = q % 6;
                Hpp[(size_t)(6 * lo + r) * D + 6 * hi + c] = 0.0;
                Hpp[(size_t)(6 * hi + c) * D + 6 * lo + r] = 0.0;
              }
      }
    }
    if (A.dbg && lane == 0) atomicAdd((unsigned long long*)&A.dbg[gridDim.x * 6 + blockIdx.x * 16 + warp], (unsigned long long)(clock64() - t2a0));
    VIML_TICK(4);
    if (part + 1 < nparts) fetch_next(w, slot, part + 1);
    else if (has_next) fetch_next(w + gridDim.x, slot + 1, 0);
    // ------------------------------------------------------------------ P2b: landmark strips, thread per feature
    for (; !BIGF;) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&S.next_feature, 32);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= F) break;
      const int l = base + lane;
      if (l >= F) continue;
      double2* __restrict__ row = reinterpret_cast<double2*>(Hlp + (size_t)l * D);
      const int k0 = S.baseB[l], k1 = S.baseB[l + 1];
      double ai[6] = {0, 0, 0, 0, 0, 0}, ae[6] = {0, 0, 0, 0, 0, 0}, dd = 0.0, dr = 0.0;
      unsigned mask = 0u;
      int istart = -1;
      for (int k = k0; k < k1; ++k) {
        const int pos = S.fperm[k];
        const uint32_t ix = S.ridx[pos];
        const int i = ix & 0xff, j = (ix >> 8) & 0xff;
        const bool sw = i > j;
        const double2 d = ldrec(S.rec, pos, 16);
        double2 Xi[6], Xj[6], C[6];
        if (!sw) {
          load_lo(S.rec, pos, Xi);
          load_hi(S.rec, pos, Xj);
        } else {
          load_hi(S.rec, pos, Xi);
          load_lo(S.rec, pos, Xj);
        }
        load_ex(S.rec, pos, C);
        const double2 r = ldrec(S.rec, pos, 15);
        double bj[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          ai[c] = fma(d.x, Xi[c].x, fma(d.y, Xi[c].y, ai[c]));
          ae[c] = fma(d.x, C[c].x, fma(d.y, C[c].y, ae[c]));
          bj[c] = d.x * Xj[c].x + d.y * Xj[c].y;
        }
        dd = fma(d.x, d.x, fma(d.y, d.y, dd));
        dr = fma(d.x, r.x, fma(d.y, r.y, dr));
        row[3 * j] = make_double2(bj[0], bj[1]);
        row[3 * j + 1] = make_double2(bj[2], bj[3]);
        row[3 * j + 2] = make_double2(bj[4], bj[5]);
        mask |= (1u << j) | (1u << i);
        istart = i;
      }
      if (accum && istart < 0) continue;  // a later part only touches the features it has factors of
      if (istart >= 0) {
        put2(&row[3 * istart], ai[0], ai[1], accum);
        put2(&row[3 * istart + 1], ai[2], ai[3], accum);
        put2(&row[3 * istart + 2], ai[4], ai[5], accum);
      }
      put2(&row[3 * E], ae[0], ae[1], accum);
      put2(&row[3 * E + 1], ae[2], ae[3], accum);
      put2(&row[3 * E + 2], ae[4], ae[5], accum);
      mask |= 1u << E;
      if (!accum)
        for (int b = 0; b < P; ++b)
          if (!((mask >> b) & 1u)) {
            row[3 * b] = make_double2(0.0, 0.0);
            row[3 * b + 1] = make_double2(0.0, 0.0);
            row[3 * b + 2] = make_double2(0.0, 0.0);
          }
      put(&Hll[l], dd, accum);
      put(&bl[l], dr, accum);
    }
    __syncthreads();  // shared memory is reused by the next part / window
    VIML_TICK(5);
    }  // parts
  }
  if (A.dbg && tid == 0)
    for (int k = 0; k < 6; ++k) A.dbg[blockIdx.x * 6 + k] = tph[k];
#undef VIML_TICK
}

}  // namespace fused
