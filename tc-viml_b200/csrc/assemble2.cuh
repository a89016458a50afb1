// assemble2.cuh — fused "evaluate + loss-correct + J^T J / J^T r" for batches of windows (mode B of viml_linearize_batch).
// Included by linearize_kernels.cu after eval_point / eval_line.
//
// What the reference does per window (marginalization_factor.cpp:3-69, :141-172): evaluate every factor into heap
// Jacobian blocks, then 4 pthreads scatter J_i^T J_j into private dense pos x pos matrices.  Here Jacobians never
// leave the SM and there is no CTA-wide phase structure: a window is cut into TASKS that warps execute on their own.
//
//   plan_kernel (indices only)   features are ordered by anchor pose; a task is a run of consecutive features
//       (<= 24 features, <= ~180 factors, cut at anchor boundaries); inside a task the factors are ordered by (anchor i, observing frame j) and
//       every (i, j) SEGMENT is padded to an even length.  The plan is a list of 32-bit slots
//       (factor id | feature slot | i | j) plus one record per task.
//   assemble_kernel, one CTA (4 warps) per window, three CTAs per SM.  A warp takes a task and streams its slots
//   32 at a time:
//     1. lane = factor: residual, Jacobians, Cauchy correction in registers (eval_point).
//     2. landmark row of the factor's feature: the d^T J_j block is unique to the factor and goes straight to
//        H_lp; d^T J_i, d^T J_ex, d^T d, d^T r are summed over the feature's factors in a warp-private
//        shared-memory row (all factors of a feature belong to one task: plain read-modify-write, no atomics).
//     3. pose part on the FP64 tensor pipe: the factor's 16 distinct columns U = [A | a_rot | b_rot | J_ex | r]
//        (the pose-j translation block is -A) go to a warp-private stage; two factors of one segment form one
//        K = 4 step of mma.sync.m8n8k4.f64 and the upper triangle of the 16 x 16 Gram matrix U^T U costs three
//        DMMA per step.  Every block of H_pp the two factors touch — (i,i), (j,j), (i,j), (i,ex), (j,ex), (ex,ex),
//        b_i, b_j, b_ex — is a signed sub-block of that Gram matrix.
//     4. at the end of a segment the j-role entries and the (i,j) block are added to the window's block-upper
//        accumulator in shared memory (shared-memory FP64 add = compare-and-swap loop; balanced over the
//        lanes through a 192-double patch and a compile-time destination table: 63 j-role + 33 (i,j) entries = three
//        rounds, the (i,j) block's symmetric top-left 3x3 travelling as its upper triangle); the i-role entries keep
//        accumulating in registers until the anchor changes, the extrinsic's own block and gradient until the warp
//        has no task left.
//   When the window's tasks are done the CTA expands the block-upper accumulator to the full symmetric D x D
//   matrix in the (now dead) warp work areas and writes it with ONE cp.async.bulk shared->global (TMA bulk store).
//
// Precondition of the fused path (it is the reference's factor structure, estimator.cpp:1735-1770): every feature
// has ONE anchor pose i, each (feature, j) pair occurs once, i != j, indices in range.  plan_kernel checks it; a
// window that violates it is zero-filled here and assembled by irregular_kernel with global atomics instead.
// The order in which segments reach the shared accumulator is not fixed: H_pp / b_p are reproducible to rounding
// (~1e-16 relative), not bit for bit; H_lp, H_ll, b_l and the per-factor outputs are bit-reproducible.
#pragma once

namespace stream {

constexpr int PMAX = 12;       // poses per window on the fused path (4-bit i/j, 12-bit observation masks)
constexpr int FMAXP = 4096;    // features per window the plan kernel's shared tables are sized for
constexpr int PT = 64;         // plan_kernel threads (warp 0 orders the features, warp 1 the line factors)
constexpr int PT_MAX = 512;    // ... for windows with many features: more warps write the tasks' slots
constexpr int TASK_T = 180;    // point factors per task at most (+ one feature's worth): an EuRoC anchor group fits whole (88 / 110 / 140 / 180: 0.446 / 0.426 / 0.415 / 0.414 ms)
constexpr int TASK_F = 24;     // features per task
constexpr int LTASK = 32;      // line slots per line task
constexpr int AW = 4;          // warps per assemble CTA (3 CTAs x 4 warps per SM = 3 warps per scheduler at <= 168 registers)
constexpr int LACC_W = 14;     // landmark row: d^T J_i (6) | d^T J_ex (6) | d^T d | d^T r
constexpr int STAGE_D = 1024;  // stage: 32 records x 32 doubles
constexpr int LACC_D = TASK_F * LACC_W;
constexpr int WORK_D = STAGE_D + LACC_D + 192;   // doubles of private work area per warp: stage | landmark rows | patch

struct PlanPtrs {
  int4* hdr;          // [W] {n_tasks, n_lslots, irregular, 0}
  int4* tasks;        // {slot_begin, n_slots, feat_begin, n_feats}; window w starts at task_base(w)
  uint32_t* slots;    // [2 NP] factor id (16) | feature slot in task (5) | i (4) | j (4); id 0xffff = padding
  uint32_t* finfo;    // [W F] anchor-sorted features: feature (16) | observation mask (12) | anchor (4)
  uint32_t* lslots;   // [2 NL] line id (16) | frame (8)
  int* any_irregular;
};

// tasks of a window: at most nf/64 cut by the factor limit, F/TASK_F by the feature limit, PMAX+1 by anchor boundaries
__host__ __device__ inline int task_base(int a0_rel, int w, int F) { return a0_rel / 64 + w * (F / TASK_F + PMAX + 4); }

// ------------------------------------------------------------------------------------------------ plan
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// One CTA per window.  Dynamic shared memory: fmask[F] u32 | fanchor[F] u32 | cexcl[F+1] u32 | ford[F] u16 | fid[F*P] u16
__global__ void __launch_bounds__(PT_MAX) plan_kernel(LinearizeArgs A, PlanPtrs PL) {
  extern __shared__ __align__(16) unsigned char plan_raw[];
  __shared__ int gcount[PMAX + 2], grun[PMAX + 2], lcount[PMAX + 1], lrun[PMAX + 1];
  __shared__ int s_bad, s_big, s_ntasks, s_nlslots;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, w = blockIdx.x;
  const int P = A.P, F = A.F, NTH = (int)blockDim.x;   // 64 threads for EuRoC-sized windows, PT_MAX for many-feature ones
  uint32_t* fmask = reinterpret_cast<uint32_t*>(plan_raw);
  uint32_t* fanchor = fmask + F;
  uint32_t* cexcl = fanchor + F;
  uint16_t* ford = reinterpret_cast<uint16_t*>(cexcl + F + 1);
  uint16_t* fid = ford + ((F + 1) & ~1);
  const int a0 = A.pf_window_offset[w], nf = A.pf_window_offset[w + 1] - a0;
  const int b0 = A.NL > 0 ? A.lf_window_offset[w] : 0, nl = A.NL > 0 ? A.lf_window_offset[w + 1] - b0 : 0;
  const int a0r = a0 - (int)A.pf_begin, b0r = b0 - (int)A.lf_begin;
  for (int f = tid; f < F; f += NTH) fmask[f] = 0u, fanchor[f] = 0xffu;
  for (int e = tid; e < F * P; e += NTH) fid[e] = 0xffff;
  if (tid < PMAX + 2) gcount[tid] = 0;
  if (tid < PMAX + 1) lcount[tid] = 0;
  if (tid == 0) s_bad = s_big = (nf > 65534 || nl > 65534 || nf < 0 || nl < 0) ? 1 : 0, s_ntasks = 0, s_nlslots = 0;
  __syncthreads();
  const int nfs = s_big ? 0 : nf;   // (s_bad itself is written by the loop below: not read before the next barrier)
  for (int k = tid; k < nfs; k += NTH) {
    const uint32_t ix = A.pf_idx[a0 + k];
    const int i = ix & 0xff, j = (ix >> 8) & 0xff, l = ix >> 16;
    if (i >= P || j >= P || l >= F || i == j) {
      s_bad = 1;
      continue;
    }
    const uint32_t old = atomicOr(&fmask[l], 1u << j);
    const uint32_t an = atomicCAS(&fanchor[l], 0xffu, (uint32_t)i);
    if (((old >> j) & 1u) || (an != 0xffu && an != (uint32_t)i)) s_bad = 1;
    fid[l * P + j] = (uint16_t)k;
  }
  __syncthreads();
  if (s_bad) {   // zero-filled by assemble_kernel, assembled by irregular_kernel
    if (tid == 0) {
      PL.hdr[w] = make_int4(0, 0, 1, 0);
      atomicOr(PL.any_irregular, 1);
    }
    return;
  }
  uint32_t* lsl = PL.lslots + 2 * (size_t)b0r;
  if (warp == 0) {
    // features ordered by anchor (stable in the feature index); features without a factor come last
    for (int f0 = 0; f0 < F; f0 += 32) {
      const int f = f0 + lane;
      const int key = f < F ? (fanchor[f] == 0xffu ? P : (int)fanchor[f]) : 64 + lane;
      const unsigned m = __match_any_sync(0xffffffffu, key);
      if (f < F && (m & ((1u << lane) - 1u)) == 0u) gcount[key] += __popc(m);
      __syncwarp();
    }
    {
      const int c = lane <= P ? gcount[lane] : 0;
      const int inc = warp_incl_scan(c, lane);
      if (lane <= P) grun[lane] = inc - c;
    }
    __syncwarp();
    for (int f0 = 0; f0 < F; f0 += 32) {
      const int f = f0 + lane;
      const int key = f < F ? (fanchor[f] == 0xffu ? P : (int)fanchor[f]) : 64 + lane;
      const unsigned m = __match_any_sync(0xffffffffu, key);
      const int rank = __popc(m & ((1u << lane) - 1u));
      if (f < F) ford[grun[key] + rank] = (uint16_t)f;
      __syncwarp();
      if (f < F && rank == 0) grun[key] += __popc(m);
      __syncwarp();
    }
    // exclusive prefix of the factor counts in that order
    int carry = 0;
    for (int p0 = 0; p0 < F; p0 += 32) {
      const int p = p0 + lane;
      const int c = p < F ? __popc(fmask[ford[p]]) : 0;
      const int inc = warp_incl_scan(c, lane);
      if (p < F) cexcl[p] = (uint32_t)(carry + inc - c);
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) cexcl[F] = (uint32_t)carry;
    __syncwarp();
    // greedy cut into tasks: as many consecutive features as fit TASK_T factors, at most TASK_F, at least one — but a task
    // ends where the anchor changes whenever the rest of the anchor group fits.  A task that holds a WHOLE anchor group is the
    // only one that ever touches the (anchor, j) blocks of H_pp: it is marked exclusive and stores them without atomics.
    const int tb = task_base(a0r, w, F);
    int pos = 0, t = 0;
    bool at_group_start = true;
    while (pos < F) {
      const int p = pos + lane;
      const int c = p < F ? (int)(cexcl[p + 1] - cexcl[p]) : 0x10000;
      const int inc = warp_incl_scan(c, lane);
      const unsigned okm = __ballot_sync(0xffffffffu, inc <= TASK_T && p < F && lane < TASK_F);
      int n = okm == 0xffffffffu ? 32 : __ffs(~okm) - 1;
      n = max(n, 1);
      n = min(n, F - pos);
      const int an = p < F ? (fanchor[ford[p]] == 0xffu ? P : (int)fanchor[ford[p]]) : -1;
      const unsigned same = __ballot_sync(0xffffffffu, an == __shfl_sync(0xffffffffu, an, 0));
      const int na = same == 0xffffffffu ? 33 : __ffs(~same) - 1;   // leading run of the first feature's anchor (33: longer than the view)
      int excl = 0;
      if (na <= n) {
        n = na;
        excl = at_group_start ? 1 : 0;
        at_group_start = true;
      } else {
        at_group_start = false;
      }
      if (lane == 0) PL.tasks[tb + t] = make_int4(2 * (int)cexcl[pos], 0, pos, n | (excl << 16));
      pos += n, ++t;
    }
    if (lane == 0) s_ntasks = t;
  } else if (warp == 1 && nl > 0) {
    // line factors ordered by frame, every frame's run padded to an even length
    bool bad = false;
    for (int k0 = 0; k0 < nl; k0 += 32) {
      const int k = k0 + lane;
      int key = k < nl ? A.lf_frame[b0 + k] : 64 + lane;
      if (k < nl && (key < 0 || key >= P)) bad = true, key = 96 + lane;
      const unsigned m = __match_any_sync(0xffffffffu, key);
      if (k < nl && key < P && (m & ((1u << lane) - 1u)) == 0u) lcount[key] += __popc(m);
      __syncwarp();
    }
    if (__any_sync(0xffffffffu, bad)) {
      if (lane == 0) s_bad = 1;
    } else {
      const int c = lane < P ? ((lcount[lane] + 1) & ~1) : 0;
      const int inc = warp_incl_scan(c, lane);
      if (lane < P) {
        lrun[lane] = inc - c;
        if (lcount[lane] & 1) lsl[inc - 1] = 0xffffu | ((uint32_t)lane << 16);   // padding slot of this frame
      }
      if (lane == 31) s_nlslots = inc;
      __syncwarp();
      for (int k0 = 0; k0 < nl; k0 += 32) {
        const int k = k0 + lane;
        const int key = k < nl ? A.lf_frame[b0 + k] : 64 + lane;
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(m & ((1u << lane) - 1u));
        if (k < nl) lsl[lrun[key] + rank] = (uint32_t)k | ((uint32_t)key << 16);
        __syncwarp();
        if (k < nl && rank == 0) lrun[key] += __popc(m);
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) {
      PL.hdr[w] = make_int4(0, 0, 1, 0);
      atomicOr(PL.any_irregular, 1);
    }
    return;
  }
  // slots of every task: lane = feature of the task; segments (i, j) in order, padded to even length
  const int tb = task_base(a0r, w, F), nt = s_ntasks;
  uint32_t* sl_w = PL.slots + 2 * (size_t)a0r;
  for (int t = warp; t < nt; t += NTH / 32) {
    int4 tk = PL.tasks[tb + t];
    const int p = tk.z + lane;
    const bool on = lane < (tk.w & 0xffff);
    const int l = on ? ford[p] : 0;
    const uint32_t m = on ? fmask[l] : 0u;
    const int an = (on && m) ? (int)fanchor[l] : 0xff;
    if (on) PL.finfo[(size_t)w * F + p] = (uint32_t)l | (m << 16) | ((uint32_t)(an & 15) << 28);
    uint32_t* sl = sl_w + tk.x;
    int off = 0;
    const unsigned lt = (1u << lane) - 1u;
    // anchors present in this task (usually one), and for each the observing frames present: only those are visited
    unsigned im = __reduce_or_sync(0xffffffffu, an < P ? 1u << an : 0u);
    while (im) {
      const int i = __ffs(im) - 1;
      im &= im - 1;
      unsigned jm = __reduce_or_sync(0xffffffffu, an == i ? m : 0u);
      while (jm) {
        const int j = __ffs(jm) - 1;
        jm &= jm - 1;
        const bool mine = an == i && ((m >> j) & 1u);
        const unsigned b = __ballot_sync(0xffffffffu, mine);
        const int n = __popc(b);
        const uint32_t key = ((uint32_t)i << 21) | ((uint32_t)j << 25);
        if (mine) sl[off + __popc(b & lt)] = (uint32_t)fid[l * P + j] | ((uint32_t)lane << 16) | key;
        if ((n & 1) && lane == 0) sl[off + n] = 0xffffu | key;
        off += (n + 1) & ~1;
      }
    }
    if (lane == 0) PL.tasks[tb + t].y = off;
  }
  if (tid == 0) PL.hdr[w] = make_int4(nt, s_nlslots, 0, 0);
}

// ------------------------------------------------------------------------------------------------ scatter tables
// The Gram matrix of U = [A(0-2) | a_rot(3-5) | b_rot(6-8) | Z(9-14) | r(15)] lives in three 8x8 tiles (rows 0-7 x
// cols 0-7, rows 0-7 x cols 8-15, rows 8-15 x cols 8-15).  A table entry sends patch[src] (tile*64 + row*8 + col) to
// one destination: kind selects the block (lo = min(i,j) role, hi role, extrinsic), off the entry inside it.
enum Kind { K_LL = 0, K_HH, K_LH, K_LE, K_HE, K_EE, K_BL, K_BH, K_BE, K_NKIND };

struct Tables {
  uint32_t seg[128];   // per segment: hi-role blocks (63 entries) and the (lo,hi) block (33: its symmetric 3x3 as upper triangle) = 3 rounds
  uint32_t lo[64];     // per anchor change: lo-role blocks (i,i), (i,ex), b_i (63 entries = two rounds)
  uint32_t ee[32];     // once per warp: (ex,ex) and b_ex (27 entries), summed in registers over all of the warp's tasks
  uint32_t line[32];   // line factors of one frame: (p,p) upper + b_p (27 entries)
};

constexpr uint32_t tab_entry(int R, int C, int kind, int off, bool neg) {
  const int tile = (R < 8 && C < 8) ? 0 : (R < 8 ? 1 : 2);
  return 0x80000000u | (uint32_t)(tile * 64 + (R & 7) * 8 + (C & 7)) | ((uint32_t)kind << 8) | ((uint32_t)off << 12) |
         (neg ? (1u << 18) : 0u);
}

// Destination (kind, off) of Gram entry (R, C), R <= C, for the three roles; the tables are emitted ordered by destination
// (kind, then offset) so that the 32 lanes of a scatter round add to runs of consecutive doubles of at most three blocks:
// the shared-memory read-modify-write of a round is then (nearly) bank-conflict free.
constexpr int col_type(int c) { return c < 3 ? 0 : (c < 6 ? 1 : (c < 9 ? 2 : (c < 15 ? 3 : 4))); }   // A X B Z res
constexpr int col_sub(int c) { return c < 3 ? c : (c < 6 ? c - 3 : (c < 9 ? c - 6 : (c < 15 ? c - 9 : 0))); }

constexpr Tables make_tables() {
  Tables T{};
  int ns = 0, nlo = 0, nli = 0, nee = 0;
  // every (kind, off) in destination order; for each, the (R, C, sign) that feeds it
  // per-segment table: the blocks that always need the atomic add first (HH, HE, BH = 63 entries), the (lo, hi) block last
  // (33 entries — the three sub-diagonal entries of its symmetric top-left 3x3 are restored by the expansion — plain stores
  // for an exclusive task): 96 entries = three full rounds
  constexpr int kind_order[K_NKIND] = {K_LL, K_LE, K_BL, K_EE, K_BE, K_HH, K_HE, K_BH, K_LH};
  for (int ko = 0; ko < K_NKIND; ++ko)
    for (int off = 0; off < 36; ++off) {
      const int kind = kind_order[ko];
      const bool vec = kind >= K_BL;
      if (vec && off >= 6) continue;
      const int dr = vec ? off : off / 6, dc = vec ? 0 : off % 6;
      for (int R = 0; R < 16; ++R)
        for (int C = R; C < 16; ++C) {
          const int tr = col_type(R), tc = col_type(C), r = col_sub(R), c = col_sub(C);
          // contributions of Gram entry (R, C) — see the role algebra in the header of this file
          bool hit = false, neg = false;
          if (tr == 0 && tc == 0) {            // (A_r, A_c), r <= c
            if (kind == K_LL || kind == K_HH) hit = dr == r && dc == c;
            if (kind == K_LH) hit = dr == r && dc == c, neg = true;   // -A^T A is symmetric: upper triangle only (expansion mirrors it)
          } else if (tr == 0 && tc == 1) {     // (A_r, X_c)
            if (kind == K_LL) hit = dr == r && dc == 3 + c;
            if (kind == K_LH) hit = dr == 3 + c && dc == r, neg = true;
          } else if (tr == 0 && tc == 2) {     // (A_r, B_c)
            if (kind == K_LH) hit = dr == r && dc == 3 + c;
            if (kind == K_HH) hit = dr == r && dc == 3 + c, neg = true;
          } else if (tr == 0 && tc == 3) {     // (A_r, Z_c)
            if (kind == K_LE) hit = dr == r && dc == c;
            if (kind == K_HE) hit = dr == r && dc == c, neg = true;
          } else if (tr == 0 && tc == 4) {     // (A_r, res)
            if (kind == K_BL) hit = dr == r;
            if (kind == K_BH) hit = dr == r, neg = true;
          } else if (tr == 1 && tc == 1) {     // (X, X)
            if (kind == K_LL) hit = dr == 3 + r && dc == 3 + c;
          } else if (tr == 1 && tc == 2) {     // (X, B)
            if (kind == K_LH) hit = dr == 3 + r && dc == 3 + c;
          } else if (tr == 1 && tc == 3) {     // (X, Z)
            if (kind == K_LE) hit = dr == 3 + r && dc == c;
          } else if (tr == 1 && tc == 4) {
            if (kind == K_BL) hit = dr == 3 + r;
          } else if (tr == 2 && tc == 2) {     // (B, B)
            if (kind == K_HH) hit = dr == 3 + r && dc == 3 + c;
          } else if (tr == 2 && tc == 3) {     // (B, Z)
            if (kind == K_HE) hit = dr == 3 + r && dc == c;
          } else if (tr == 2 && tc == 4) {
            if (kind == K_BH) hit = dr == 3 + r;
          } else if (tr == 3 && tc == 3) {     // (Z, Z)
            if (kind == K_EE) hit = dr == r && dc == c;
          } else if (tr == 3 && tc == 4) {
            if (kind == K_BE) hit = dr == r;
          }
          if (!hit) continue;
          const bool per_segment = kind == K_HH || kind == K_LH || kind == K_HE || kind == K_BH;
          if (per_segment) {
            T.seg[ns++] = tab_entry(R, C, kind, off, neg);
          } else if (kind == K_EE || kind == K_BE) {
            T.ee[nee++] = tab_entry(R, C, kind, off, neg);
          } else {
            T.lo[nlo++] = tab_entry(R, C, kind, off, neg);
          }
        }
    }
  // line factors: columns 0..5 = J, 6 = residual (tile 0 only)
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) T.line[nli++] = tab_entry(r, c, K_LL, r * 6 + c, false);
  for (int r = 0; r < 6; ++r) T.line[nli++] = tab_entry(r, 6, K_BL, r, false);
  return T;
}
// read once per CTA with lane-indexed (coalesced) loads: global memory, not __constant__ (divergent constant reads replay)
__device__ const Tables g_tables = make_tables();
static_assert(make_tables().seg[95] != 0u && make_tables().seg[96] == 0u, "63 hi-role + 33 (lo, hi) per-segment destinations = three full rounds");
static_assert(make_tables().lo[62] != 0u && make_tables().lo[63] == 0u, "63 per-anchor destinations");
static_assert(make_tables().ee[26] != 0u && make_tables().ee[27] == 0u, "27 extrinsic destinations");
static_assert(make_tables().line[26] != 0u && make_tables().line[27] == 0u, "27 line destinations");


// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// index of block (br, bc), br <= bc, in the block-upper layout with NB blocks per side
__device__ __forceinline__ int blk(int br, int bc, int NB) { return br * NB - (br * (br - 1)) / 2 + (bc - br); }

// stage: 32 records x 16 units (unit = one column as (row 0, row 1), 16 bytes).  Unit c of record f sits at
// f*16 + (c ^ swz(f)) so that the record stores (STS.128, a quarter warp per wavefront) and the LDS.64 fragment loads
// (a half warp = two records per wavefront) are both conflict-free.
__device__ __forceinline__ int swz(int f) { return ((f & 1) << 2) ^ ((f >> 1) & 3); }

// One scatter round per table word: patch[src] goes to Hc[base(kind) + off] with a shared-memory FP64 add (CAS loop).
// Lane k < K_NKIND holds the base of destination kind k in `mybase`.  The table words are decoded once per kernel.
struct Dest {
  int src;        // double index in the patch, -1 = no entry
  int off;        // offset inside the destination block
  int kind;
  unsigned sign;  // 0x80000000 to negate
};
__device__ __forceinline__ Dest decode(uint32_t e) {
  Dest d;
  d.src = (e >> 31) ? (int)(e & 0xff) : -1;
  d.off = (e >> 12) & 63;
  d.kind = (e >> 8) & 15;
  d.sign = ((e >> 18) & 1u) << 31;
  return d;
}
template <int ROUNDS>
__device__ __forceinline__ void scatter(const double* __restrict__ patch, const Dest* tab, double* __restrict__ Hc,
                                        int mybase, bool store_lh = false) {
  // destinations and values of every round first (the shuffles and patch loads overlap), then the adds one after the other
  double v[ROUNDS];
  double* dst[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int base = __shfl_sync(0xffffffffu, mybase, tab[r].kind);
    dst[r] = Hc + base + tab[r].off;
    const double x = patch[tab[r].src >= 0 ? tab[r].src : 0];
    v[r] = __hiloint2double(__double2hiint(x) ^ (int)tab[r].sign, __double2loint(x));
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    if (tab[r].src >= 0) {
      if (store_lh && tab[r].kind == K_LH) *dst[r] = v[r];   // single writer, written once: no read-modify-write
      else atomicAdd(dst[r], v[r]);
    }
  }
}
// (Measured and dropped, round 2: per-pose spin locks with plain read-modify-write under the lock — 0.72 vs 0.56 ms for
// the compare-and-swap adds; taking entries out with a 64-bit exchange + sentinel so that the rounds of a flush pipeline —
// 0.68 ms, the per-round address/value registers spill at the 168-register cap; prefetching the next chunk's factor inputs
// into registers — spills as well; prefetch.global.L2 of the inputs of this window and of the window that runs on the SM one CTA
// lifetime later, issued at CTA start — 0.507 vs 0.468 ms.
// Later in round 2, against 0.465 ms: the accumulator in global memory (the window's own H_pp region, zeroed, red.global.add.f64
// fire-and-forget, read back before the expansion) — 0.470 ms, correct but no faster; one CTA-wide lock per flush (lane 0, 32-bit
// atomicCAS) with plain vector read-modify-write under it — 0.95 ms (ATOMS.CAS is far slower than the ATOMS.CAST.SPIN of the
// compiler's own FP64 add loop); the rounds' compare-and-swaps written out with atomicCAS so that two or four are in flight —
// 0.54 ms, same reason; the next chunk's factor inputs loaded right after the Jacobians leave the registers — 0.466 ms, no
// change; K-step loop unrolled by four — 0.463 ms.  Kept: shuffles and patch loads of all rounds hoisted above the first add
// (0.465 -> 0.457 ms); the (lo, hi) block flushed as 33 entries (its symmetric 3x3 as upper triangle) so that a segment
// flush is three rounds instead of four (0.457 -> 0.436 ms); (ex,ex) / b_ex summed per warp in registers (lo flush two rounds;
// no change on cfg 2).  Dropped: the (lo, hi) block of an exclusive task stored straight from the fragment registers (two rounds
// through the patch): 0.4415 vs 0.4356 ms.  profiles/assemble_knockouts_r2.md has the knock-out timings that say where the
// time goes.  After that: every global load of the CTA prologue issued before the first use (0.4356 -> ~0.427 ms); the inverse
// depth fetched once per task and handed to the factors by shuffle instead of a dependent load: 0.4397 vs 0.4356 ms, dropped.)

// Per-lane selectors of kind_base: which pose (0 = lo, 1 = hi, 2 = extrinsic) is the block row / column of kind `lane`.
//   kind      LL HH LH LE HE EE BL BH BE
//   row        0  1  0  0  1  2  0  1  2      col   0  1  1  2  2  2  -  -  -
__device__ __forceinline__ int kind_base(int rs, int cs, bool isb, int lo, int hi, int NB, int boff) {
  const int e = NB - 1;
  const int x = rs == 0 ? lo : (rs == 1 ? hi : e), y = cs == 0 ? lo : (cs == 1 ? hi : e);
  return isb ? boff + 6 * x : blk(x, y, NB) * 36;
}

__device__ __forceinline__ void put_patch(double* __restrict__ patch, int lane, const double (&G)[6]) {
  // lane holds rows g = lane>>2, columns 2*(lane&3) + {0,1} of each tile
  double2* p2 = reinterpret_cast<double2*>(patch);
  const int o = (lane >> 2) * 4 + (lane & 3);
  p2[o] = make_double2(G[0], G[1]);
  p2[32 + o] = make_double2(G[2], G[3]);
  p2[64 + o] = make_double2(G[4], G[5]);
}

// ------------------------------------------------------------------------------------------------ assemble
template <bool MODE_A>
__global__ void __maxnreg__(168) assemble_kernel(LinearizeArgs A, PlanPtrs PL, int use_tma) {
  extern __shared__ __align__(16) unsigned char asm_raw[];
  __shared__ int s_next;
  constexpr int NT = AW * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, w = blockIdx.x;
  const int P = A.P, F = A.F, D = A.D, NB = P + 1;
  const int nblk = NB * (NB + 1) / 2, boff = nblk * 36;
  const int hc_n = (boff + D + 1) & ~1;
  const int cstride = P * kPoseCache + kExCache;
  double* __restrict__ Hc = reinterpret_cast<double*>(asm_raw);
  double* __restrict__ cache = Hc + hc_n;
  double* __restrict__ work0 = cache + cstride;
  double* __restrict__ stage = work0 + warp * WORK_D;
  double* __restrict__ lacc = stage + STAGE_D;
  double* __restrict__ patch = lacc + LACC_D;

  const int4 hdr = PL.hdr[w];
  const int a0 = A.pf_window_offset[w];
  const int b0 = A.NL > 0 ? A.lf_window_offset[w] : 0;
  double* __restrict__ Hpp = A.out.H_pp + (size_t)w * D * D;
  double* __restrict__ Hlp = A.out.H_lp + (size_t)w * F * D;
  double* __restrict__ Hll = A.out.H_ll + (size_t)w * F;
  double* __restrict__ bp = A.out.b_p + (size_t)w * D;
  double* __restrict__ bl = A.out.b_l + (size_t)w * F;
  // every global load of the prologue is issued before the first one is needed (one exposed latency instead of three): the
  // window's pose cache (up to three 16-byte units per thread) and the scatter tables
  const double2* __restrict__ gc = reinterpret_cast<const double2*>(A.cache + (size_t)w * cstride);
  double2 cpre[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) cpre[q] = tid + q * NT < cstride / 2 ? gc[tid + q * NT] : make_double2(0.0, 0.0);
  uint32_t traw[7];
#pragma unroll
  for (int r = 0; r < 3; ++r) traw[r] = g_tables.seg[32 * r + lane];
#pragma unroll
  for (int r = 0; r < 2; ++r) traw[3 + r] = g_tables.lo[32 * r + lane];
  traw[5] = g_tables.ee[lane], traw[6] = g_tables.line[lane];
  if (hdr.z) {   // irregular window: zero here, irregular_kernel adds with global atomics
    for (int e = tid; e < D * D; e += NT) Hpp[e] = 0.0;
    for (int e = tid; e < F * D; e += NT) Hlp[e] = 0.0;
    for (int e = tid; e < F; e += NT) Hll[e] = 0.0, bl[e] = 0.0;
    for (int e = tid; e < D; e += NT) bp[e] = 0.0;
    return;
  }
  {
    double2* z = reinterpret_cast<double2*>(Hc);
    for (int e = tid; e < hc_n / 2; e += NT) z[e] = make_double2(0.0, 0.0);
    double2* c2 = reinterpret_cast<double2*>(cache);
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (tid + q * NT < cstride / 2) c2[tid + q * NT] = cpre[q];
    for (int e = tid + 3 * NT; e < cstride / 2; e += NT) c2[e] = gc[e];   // (P <= 12: never taken)
  }
  if (tid == 0) s_next = 0;
  Dest tab_seg[3], tab_lo[2], tab_ee[1], tab_line[1];
#pragma unroll
  for (int r = 0; r < 3; ++r) tab_seg[r] = decode(traw[r]);
#pragma unroll
  for (int r = 0; r < 2; ++r) tab_lo[r] = decode(traw[3 + r]);
  tab_ee[0] = decode(traw[5]);
  tab_line[0] = decode(traw[6]);
  const int kb_rs = lane < K_NKIND ? (0x24904 >> (2 * lane)) & 3 : 0, kb_cs = lane < K_NKIND ? (0xA94 >> (2 * lane)) & 3 : 0;
  const bool kb_isb = lane >= K_BL && lane < K_NKIND;
  __syncthreads();

  const int ntask = hdr.x, nlsl = hdr.y, nltask = (nlsl + LTASK - 1) / LTASK;
  const int tb = task_base(a0 - (int)A.pf_begin, w, F);
  const uint32_t* __restrict__ sl_w = PL.slots + 2 * (size_t)(a0 - (int)A.pf_begin);
  const uint32_t* __restrict__ lsl_w = PL.lslots + 2 * (size_t)(b0 - (int)A.lf_begin);
  // fragment geometry of mma.m8n8k4: lane (g, kk) holds column g of row kk&1 of factor kk>>1 (A and B operand alike)
  const int g = lane >> 2, kk = lane & 3;
  const int gx = g ^ ((kk >> 1) << 2);
  const double* __restrict__ frag0 = stage + (kk >> 1) * 32 + (kk & 1);   // record 2s + (kk>>1), row kk&1
  const unsigned full = 0xffffffffu;
  // landmark-row units of this lane: u = lane, and 32 + (lane & 7) in the tail pass (a row has 3 NB <= 39 units of 16 bytes)
  const int upr = 3 * NB;
  const int ru1 = 32 + (lane & 7), rb0 = (lane * 11) >> 5, rb1 = (ru1 * 43) >> 7;   // tail units: 8 lanes per feature
  const int rp0 = lane - 3 * rb0, rp1 = ru1 - 3 * rb1;

  double E[2] = {0.0, 0.0};   // tile 2 of the Gram matrix summed over every point factor this warp sees: (ex,ex) and b_ex
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&s_next, 1);
    t = __shfl_sync(full, t, 0);
    if (t >= ntask + nltask) break;
    if (t < ntask) {
      // ================================================================== point task
      const int4 tk = PL.tasks[tb + t];
      const int n_slots = tk.y, n_feats = tk.w & 0xffff;
      const bool excl = (tk.w >> 16) != 0;   // whole anchor group: the (lo, hi) blocks have no other writer
      const uint32_t fi = lane < n_feats ? PL.finfo[(size_t)w * F + tk.z + lane] : 0u;
      {
        double2* z = reinterpret_cast<double2*>(lacc);
        for (int q = lane; q < LACC_D / 2; q += 32) z[q] = make_double2(0.0, 0.0);
      }
      __syncwarp();
      const uint32_t* __restrict__ sl = sl_w + tk.x;
      double G[6] = {0, 0, 0, 0, 0, 0}, R[4] = {0, 0, 0, 0};
      uint32_t nsw = lane < n_slots ? sl[lane] : 0xffffu;
      for (int c0 = 0; c0 < n_slots; c0 += 32) {
        const uint32_t sw_ = nsw;
        nsw = c0 + 32 + lane < n_slots ? sl[c0 + 32 + lane] : 0xffffu;
        const bool valid = (sw_ & 0xffffu) != 0xffffu;
        const int lf = (sw_ >> 16) & 31;
        {
          const int i = (sw_ >> 21) & 15, j = (sw_ >> 25) & 15;
          const int l = __shfl_sync(full, fi, lf) & 0xffff;
          PointJac J;
          if (valid) {
            const int64_t k = (int64_t)a0 + (sw_ & 0xffffu);
            const double4 ob = reinterpret_cast<const double4*>(A.pf_obs)[k];
            const double piz = A.pf_pts_i_z ? A.pf_pts_i_z[k] : 1.0;
            const double lam = A.inv_depth[(size_t)w * F + l];
            eval_point(A, cache, i, j, lam, ob.x, ob.y, piz, ob.z, ob.w, J);
            if (MODE_A) {
              if (A.out.pf_residual) reinterpret_cast<double2*>(A.out.pf_residual)[k] = make_double2(J.r[0], J.r[1]);
              if (A.out.pf_jac_pose_i) store_jac7(A.out.pf_jac_pose_i + 14 * k, J.a);
              if (A.out.pf_jac_pose_j) store_jac7(A.out.pf_jac_pose_j + 14 * k, J.b);
              if (A.out.pf_jac_ex) store_jac7(A.out.pf_jac_ex + 14 * k, J.c);
              if (A.out.pf_jac_feat) reinterpret_cast<double2*>(A.out.pf_jac_feat)[k] = make_double2(J.d[0], J.d[1]);
            }
            // the factor's own block of the landmark row: d^T J_j  (unique per (feature, j))
            double2* row = reinterpret_cast<double2*>(Hlp + (size_t)l * D + 6 * j);
            row[0] = make_double2(J.d[0] * J.b[0][0] + J.d[1] * J.b[1][0], J.d[0] * J.b[0][1] + J.d[1] * J.b[1][1]);
            row[1] = make_double2(J.d[0] * J.b[0][2] + J.d[1] * J.b[1][2], J.d[0] * J.b[0][3] + J.d[1] * J.b[1][3]);
            row[2] = make_double2(J.d[0] * J.b[0][4] + J.d[1] * J.b[1][4], J.d[0] * J.b[0][5] + J.d[1] * J.b[1][5]);
          } else {
#pragma unroll
            for (int c = 0; c < 6; ++c) J.a[0][c] = J.a[1][c] = J.b[0][c] = J.b[1][c] = J.c[0][c] = J.c[1][c] = 0.0;
            J.r[0] = J.r[1] = J.d[0] = J.d[1] = 0.0;
          }
          // record -> stage (the lo role is the smaller pose index)
          {
            const bool swp = i > j;
            const int sx = swz(lane);
            double2* rec = reinterpret_cast<double2*>(stage) + lane * 16;
#pragma unroll
            for (int c = 0; c < 6; ++c) rec[c ^ sx] = swp ? make_double2(J.b[0][c], J.b[1][c]) : make_double2(J.a[0][c], J.a[1][c]);
#pragma unroll
            for (int c = 0; c < 3; ++c)
              rec[(6 + c) ^ sx] = swp ? make_double2(J.a[0][3 + c], J.a[1][3 + c]) : make_double2(J.b[0][3 + c], J.b[1][3 + c]);
#pragma unroll
            for (int c = 0; c < 6; ++c) rec[(9 + c) ^ sx] = make_double2(J.c[0][c], J.c[1][c]);
            rec[15 ^ sx] = make_double2(J.r[0], J.r[1]);
          }
          // landmark row sums: lanes of one segment have distinct features; lanes that share a feature take turns
          {
            const unsigned grp = __match_any_sync(full, valid ? lf : 32 + lane);
            const int rank = __popc(grp & ((1u << lane) - 1u));
            const int rounds = __reduce_max_sync(full, valid ? __popc(grp) : 0);
            double v[LACC_W];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              v[c] = J.d[0] * J.a[0][c] + J.d[1] * J.a[1][c];
              v[6 + c] = J.d[0] * J.c[0][c] + J.d[1] * J.c[1][c];
            }
            v[12] = J.d[0] * J.d[0] + J.d[1] * J.d[1];
            v[13] = J.d[0] * J.r[0] + J.d[1] * J.r[1];
            double2* __restrict__ la = reinterpret_cast<double2*>(lacc + lf * LACC_W);
            for (int r = 0; r < rounds; ++r) {
              if (valid && rank == r) {
#pragma unroll
                for (int c = 0; c < LACC_W / 2; ++c) {
                  double2 o = la[c];
                  o.x += v[2 * c], o.y += v[2 * c + 1];
                  la[c] = o;
                }
              }
              __syncwarp();
            }
          }
        }
        __syncwarp();
        // segment ends: bit q of bnd = the slot after lane q belongs to another segment (or the task ends)
        const uint32_t key = sw_ >> 21;
        uint32_t nk = __shfl_down_sync(full, key, 1);
        const uint32_t nk31 = __shfl_sync(full, nsw >> 21, 0);
        if (lane == 31) nk = nk31;
        if (c0 + lane + 1 >= n_slots) nk = 0xffffffffu;
        const unsigned bnd = __ballot_sync(full, key != nk);
        const int nks = min(16, (n_slots - c0) >> 1);
        int s = 0;
        while (s < nks) {
          const unsigned rem = bnd >> (2 * s + 1);   // bit 2q: the segment ends after K-step s + q
          const int q_end = rem ? ((__ffs(rem) - 1) >> 1) : 32;
          const bool ends = s + q_end < nks;
          const int s_end = ends ? s + q_end + 1 : nks;
#pragma unroll 2
          for (; s < s_end; ++s) {
            const double* __restrict__ p = frag0 + s * 64 + 2 * (gx ^ (s & 3));
            const double u0 = p[0], u1 = p[16];
            dmma(G[0], G[1], u0, u0);
            dmma(G[2], G[3], u0, u1);
            dmma(G[4], G[5], u1, u1);
          }
          if (ends) {
            const int la_ = 2 * s - 2;   // first lane of the segment's last K-step
            const uint32_t kseg = __shfl_sync(full, key, la_), knext = __shfl_sync(full, nk, la_ + 1);
            const int si = kseg & 15, sj = (kseg >> 4) & 15;
            const int lo = min(si, sj), hi = max(si, sj);
            const int mybase = kind_base(kb_rs, kb_cs, kb_isb, lo, hi, NB, boff);
            put_patch(patch, lane, G);
            __syncwarp();
            scatter<3>(patch, tab_seg, Hc, mybase, excl);
#pragma unroll
            for (int q = 0; q < 4; ++q) R[q] += G[q], G[q] = 0.0;
            E[0] += G[4], E[1] += G[5], G[4] = G[5] = 0.0;
            const bool lo_ends = knext == 0xffffffffu || min((int)(knext & 15), (int)((knext >> 4) & 15)) != lo;
            if (lo_ends) {
              __syncwarp();
              {
                const double R6[6] = {R[0], R[1], R[2], R[3], 0.0, 0.0};   // tile 2 holds no lo-role entry
                put_patch(patch, lane, R6);
              }
              __syncwarp();
              scatter<2>(patch, tab_lo, Hc, mybase);
#pragma unroll
              for (int q = 0; q < 4; ++q) R[q] = 0.0;
            }
            __syncwarp();
          }
        }
        __syncwarp();   // the stage is rewritten by the next chunk
      }
      // rows of the task's features: anchor block, extrinsic block, structural zeros (the observing frames' blocks
      // were written by the factors), H_ll, b_l.  One feature per iteration, lane = 16-byte unit of the row.
      for (int f = 0; f < n_feats; ++f) {
        const uint32_t info = __shfl_sync(full, fi, f);
        const int l = info & 0xffff;
        const uint32_t m = (info >> 16) & 0xfffu;
        const int an = m ? (int)(info >> 28) : -1;
        double2* __restrict__ dst = reinterpret_cast<double2*>(Hlp + (size_t)l * D);
        const double2* __restrict__ la = reinterpret_cast<const double2*>(lacc + f * LACC_W);
        const bool ia = rb0 == an, ie = rb0 == P;
        double2 v = la[ia ? rp0 : 3 + rp0];
        if (!(ia || ie)) v = make_double2(0.0, 0.0);
        if ((ia || ie || !((m >> rb0) & 1u)) && lane < upr) dst[lane] = v;
      }
      // units 32 .. upr-1 (at most 7): four features per step, eight lanes each
      if (upr > 32) {
        const int fs = lane >> 3;
        for (int f0 = 0; f0 < n_feats; f0 += 4) {
          const int f = f0 + fs;
          const uint32_t info = __shfl_sync(full, fi, f & 31);
          const int l = info & 0xffff;
          const uint32_t m = (info >> 16) & 0xfffu;
          const int an = m ? (int)(info >> 28) : -1;
          const bool ia = rb1 == an, ie = rb1 == P;
          if (f < n_feats && ru1 < upr) {
            double2 v = reinterpret_cast<const double2*>(lacc + f * LACC_W)[ia ? rp1 : 3 + rp1];
            if (!(ia || ie)) v = make_double2(0.0, 0.0);
            if (ia || ie || !((m >> rb1) & 1u)) reinterpret_cast<double2*>(Hlp + (size_t)l * D)[ru1] = v;
          }
        }
      }
      if (lane < n_feats) {
        const int l = fi & 0xffff;
        Hll[l] = lacc[lane * LACC_W + 12];
        bl[l] = lacc[lane * LACC_W + 13];
      }
      __syncwarp();
    } else {
      // ================================================================== line task
      const int t0 = (t - ntask) * LTASK, n_slots = min(LTASK, nlsl - t0);
      const uint32_t sw_ = lane < n_slots ? lsl_w[t0 + lane] : 0xffffu;
      const bool valid = (sw_ & 0xffffu) != 0xffffu;
      const int frame = (sw_ >> 16) & 0xff;
      {
        LineJac J;
        if (valid) {
          const int64_t k = (int64_t)b0 + (sw_ & 0xffffu);
          double g9[9];
#pragma unroll
          for (int c = 0; c < 9; ++c) g9[c] = A.lf_geom[(size_t)c * A.NL_stride + k];
          eval_line(A, cache, frame, g9, J);
          if (MODE_A) {
            if (A.out.lf_residual) reinterpret_cast<double2*>(A.out.lf_residual)[k] = make_double2(J.r[0], J.r[1]);
            if (A.out.lf_jac_pose) store_jac7(A.out.lf_jac_pose + 14 * k, J.a);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 6; ++c) J.a[0][c] = J.a[1][c] = 0.0;
          J.r[0] = J.r[1] = 0.0;
        }
        const int sx = swz(lane);
        double2* rec = reinterpret_cast<double2*>(stage) + lane * 16;
#pragma unroll
        for (int c = 0; c < 6; ++c) rec[c ^ sx] = make_double2(J.a[0][c], J.a[1][c]);
        rec[6 ^ sx] = make_double2(J.r[0], J.r[1]);
        rec[7 ^ sx] = make_double2(0.0, 0.0);
      }
      __syncwarp();
      const uint32_t key = sw_ >> 16;
      uint32_t nk = __shfl_down_sync(full, key, 1);
      if (lane + 1 >= n_slots) nk = 0xffffffffu;
      const unsigned bnd = __ballot_sync(full, key != nk);
      const int nks = n_slots >> 1;
      double G[6] = {0, 0, 0, 0, 0, 0};
      int s = 0;
      while (s < nks) {
        const unsigned rem = bnd >> (2 * s + 1);
        const int s_end = min(nks, s + ((__ffs(rem) - 1) >> 1) + 1);   // a line task always ends on a segment end
        for (; s < s_end; ++s) {
          const double u0 = frag0[s * 64 + 2 * (gx ^ (s & 3))];
          dmma(G[0], G[1], u0, u0);
        }
        const int fr = __shfl_sync(full, key, 2 * s - 2) & 0xff;
        const int mybase = kind_base(kb_rs, kb_cs, kb_isb, fr, fr, NB, boff);
        put_patch(patch, lane, G);
        __syncwarp();
        scatter<1>(patch, tab_line, Hc, mybase);
        G[0] = G[1] = 0.0;
        __syncwarp();
      }
    }
  }
  {
    // the extrinsic's own block and gradient, once per warp
    const int mybase = kind_base(kb_rs, kb_cs, kb_isb, 0, 0, NB, boff);
    const double E6[6] = {0.0, 0.0, 0.0, 0.0, E[0], E[1]};
    __syncwarp();
    put_patch(patch, lane, E6);
    __syncwarp();
    scatter<1>(patch, tab_ee, Hc, mybase);
  }
  __syncthreads();
  // expand the block-upper accumulator to the full symmetric matrix (+ b_p behind it) in the warps' work areas
  double* __restrict__ Hf = work0;
  const bool staged = D * D + D <= AW * WORK_D;   // P = 12 does not fit: written straight from the accumulator
  const bool aligned16 = use_tma != 0;     // the launcher's check of the caller's H_pp / b_p pointers
  if (!staged) {
    Hf = Hpp;
    use_tma = 0;
  }
  if (!staged && !aligned16) {             // rare: large P and an 8-byte aligned output — scalar stores
    for (int e = tid; e < D * D; e += NT) {
      const int r = e / D, c = e - r * D;
      const int br = r / 6, bc = c / 6, rr = r - 6 * br, cc = c - 6 * bc;
      const int lo = min(br, bc), hi = max(br, bc);
      const bool tr = br > bc || (br == bc && rr > cc);
      int a = tr ? cc : rr, b = tr ? rr : cc;   // entry (a, b) of the upper block (lo, hi)
      if (lo != hi && hi < NB - 1 && a < 3 && b < a) { const int t = a; a = b, b = t; }   // symmetric 3x3 of a pose-pose block
      Hpp[e] = Hc[blk(lo, hi, NB) * 36 + a * 6 + b];
    }
  } else {
    // block by block: the 18 column pairs of an upper block go out as they are, and — for an off-diagonal block — once more
    // transposed into the mirrored block (D is even, a pair never straddles a 6x6 block); a diagonal block is symmetrised
    double2* __restrict__ Hf2 = reinterpret_cast<double2*>(Hf);
    const int Dh = D / 2;
    int br = 0, row0 = 0, rown = NB;   // block row br holds the blocks row0 .. row0 + rown - 1 of the block-upper layout
    for (int e = tid; e < nblk * 18; e += NT) {
      const int q = (e * 3641) >> 16, k = e - 18 * q;   // e < 18 * 91
      const int rr = (k * 11) >> 5, cp = k - 3 * rr, cc = 2 * cp;
      while (q >= row0 + rown) row0 += rown, --rown, ++br;
      const int bc = br + (q - row0);
      const double* __restrict__ sb = Hc + q * 36;
      if (br != bc) {
        // a pose-pose block holds its symmetric top-left 3x3 (-A^T A) as the upper triangle only
        const bool pp = bc < NB - 1;
        auto at = [&](int a, int b) { return sb[(pp && a < 3 && b < a) ? b * 6 + a : a * 6 + b]; };
        Hf2[(6 * br + rr) * Dh + 3 * bc + cp] = make_double2(at(rr, cc), at(rr, cc + 1));
        Hf2[(6 * bc + rr) * Dh + 3 * br + cp] = make_double2(at(cc, rr), at(cc + 1, rr));
      } else {
        Hf2[(6 * br + rr) * Dh + 3 * bc + cp] =
            make_double2(rr <= cc ? sb[rr * 6 + cc] : sb[cc * 6 + rr], rr <= cc + 1 ? sb[rr * 6 + cc + 1] : sb[(cc + 1) * 6 + rr]);
      }
    }
  }
  if (staged) {
    for (int e = tid; e < D; e += NT) Hf[D * D + e] = Hc[boff + e];
  } else {
    for (int e = tid; e < D; e += NT) bp[e] = Hc[boff + e];
  }
  if (use_tma) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(Hf);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Hpp), "r"(s0), "r"((uint32_t)(D * D * 8))
                   : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(bp), "r"(s0 + (uint32_t)(D * D * 8)),
                   "r"((uint32_t)(D * 8))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else if (staged) {
    __syncthreads();
    for (int e = tid; e < D * D; e += NT) Hpp[e] = Hf[e];
    for (int e = tid; e < D; e += NT) bp[e] = Hf[D * D + e];
  }
}

// Windows the plan rejected (a feature with two anchors, a repeated (feature, j), i == j, indices out of range):
// every factor adds its blocks with global atomics onto the zeros assemble_kernel wrote.  Exits at once when the
// batch has no such window.
template <bool MODE_A>
__global__ void __launch_bounds__(128) irregular_kernel(LinearizeArgs A, PlanPtrs PL) {
  if (!*PL.any_irregular) return;
  const int cstride = A.P * kPoseCache + kExCache;
  for (int w = blockIdx.x; w < A.W; w += gridDim.x) {
    if (!PL.hdr[w].z) continue;
    const double* cw = A.cache + (size_t)w * cstride;
    const int a0 = A.pf_window_offset[w], a1 = A.pf_window_offset[w + 1];
    for (int64_t k = a0 + threadIdx.x; k < a1; k += blockDim.x) {
      const uint32_t pk = A.pf_idx[k];
      const int i = pk & 0xff, j = (pk >> 8) & 0xff, f = pk >> 16;
      if (i >= A.P || j >= A.P || f >= A.F) continue;   // out-of-range indices are dropped, never dereferenced
      const double4 ob = reinterpret_cast<const double4*>(A.pf_obs)[k];
      const double piz = A.pf_pts_i_z ? A.pf_pts_i_z[k] : 1.0;
      PointJac J;
      eval_point(A, cw, i, j, A.inv_depth[(size_t)w * A.F + f], ob.x, ob.y, piz, ob.z, ob.w, J);
      if (MODE_A) {
        if (A.out.pf_residual) reinterpret_cast<double2*>(A.out.pf_residual)[k] = make_double2(J.r[0], J.r[1]);
        if (A.out.pf_jac_pose_i) store_jac7(A.out.pf_jac_pose_i + 14 * k, J.a);
        if (A.out.pf_jac_pose_j) store_jac7(A.out.pf_jac_pose_j + 14 * k, J.b);
        if (A.out.pf_jac_ex) store_jac7(A.out.pf_jac_ex + 14 * k, J.c);
        if (A.out.pf_jac_feat) reinterpret_cast<double2*>(A.out.pf_jac_feat)[k] = make_double2(J.d[0], J.d[1]);
      }
      point_atomics(A, w, i, j, f, J, true, false);   // lanes diverge here (continue above): no warp reduction
    }
    if (A.NL > 0) {
      const int b0 = A.lf_window_offset[w], b1 = A.lf_window_offset[w + 1];
      for (int64_t k = b0 + threadIdx.x; k < b1; k += blockDim.x) {
        const int frame = A.lf_frame[k];
        if (frame < 0 || frame >= A.P) continue;
        double g9[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) g9[c] = A.lf_geom[(size_t)c * A.NL_stride + k];
        LineJac J;
        eval_line(A, cw, frame, g9, J);
        if (MODE_A) {
          if (A.out.lf_residual) reinterpret_cast<double2*>(A.out.lf_residual)[k] = make_double2(J.r[0], J.r[1]);
          if (A.out.lf_jac_pose) store_jac7(A.out.lf_jac_pose + 14 * k, J.a);
        }
        line_atomics(A, w, frame, J);
      }
    }
  }
}

}  // namespace stream
