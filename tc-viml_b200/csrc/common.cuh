// common.cuh — context, error handling and launch declarations shared by the CUDA translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "viml.h"

#define VIML_TRY_CUDA(ctx, expr)                                                              \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                      \
      return VIML_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

// Growable device buffer; contents are not preserved across grow().
struct DeviceArena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  cudaError_t reserve(size_t bytes) {
    used = 0;
    if (bytes <= cap) return cudaSuccess;
    if (base) cudaFree(base);
    base = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + (1u << 20);
    cudaError_t e = cudaMalloc((void**)&base, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <class T>
  T* take(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = (T*)(base + used);
    used += bytes;
    return p;
  }
  static size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = used = 0;
  }
};

struct viml_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t copy_stream2 = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  cudaStream_t aux_stream = nullptr;            // plan_kernel runs here, beside prep_windows_kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  viml_config cfg{};
  std::string err;
  int64_t launches = 0;
  // association decision threshold: angle > angle_th  <=>  |dot| < cos_th  
  double cos_th = 0.0;
  int nan_angle_passes = 0;  // !(3.1415926 > angle_th)
  // prior map, six SoA planes of n_map doubles
  double* d_map = nullptr;
  int64_t n_map = 0;
  bool map_set = false;   // viml_set_map was called (an empty map is legal)
  // spatially sorted copy for the hierarchical cull: Morton order of segment mid-points, tiles of kMapTile
  // lines with a bounding sphere each (centre xyz + radius), and the original index of every sorted line
  double* d_map_sorted = nullptr;   // [6][n_map]
  int32_t* d_map_orig = nullptr;    // [n_map]
  double* d_tile_sphere = nullptr;  // [n_tiles][4]
  double* d_group_sphere = nullptr; // [ceil(n_tiles / 16)][4]: one sphere around 16 consecutive tile spheres
  int64_t n_tiles = 0;
  // FoV cache: one mask row per window slot (viml_fov_update / viml_fov_slide), sized for the current map
  uint32_t* d_fov_slots = nullptr;     // [VIML_FOV_SLOTS][fov_words]
  int32_t* d_fov_slot_count = nullptr; // [VIML_FOV_SLOTS]
  int64_t fov_words = 0;
  unsigned long long* d_assoc_stats = nullptr;  // {gate tests, gated pairs, overlap-scored, distance-scored} of the last association call
  DeviceArena in_arena, out_arena, scratch, scratch2, scratch3, gn_in, gn_out, s_full;
  bool schur_splitk = false;   // VIML_SCHUR_SPLITK=1: round-1 split-K Schur kernel instead of the TMA-pipelined one
  // pinned staging of the small-batch host path (one H2D, one D2H per call)
  char *h_stage_in = nullptr, *h_stage_out = nullptr;
  size_t h_stage_in_cap = 0, h_stage_out_cap = 0;
  void* nccl_lib = nullptr;
  // profiling (viml_profile_begin/end): event pairs per kernel id
  bool brute_cull = false;     // VIML_BRUTE_CULL=1: the literal all-pairs FoV sweep (roofline accounting, cross-check)
  bool force_generic = false;  // tests: route everything through the generic atomic kernels
  bool prof = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[VIML_NUM_KERNELS];
};

enum KernelId {
  K_PREP = 0, K_POINTS, K_LINES, K_ASSEMBLE, K_SCHUR, K_CAMPOSE, K_CULL, K_SCAN, K_FILL, K_PROJECT, K_MATCH,
  K_MARG, K_MICRO, K_PLAN, K_IRREGULAR, K_GN, K_COUNT
};
static_assert(K_COUNT <= VIML_NUM_KERNELS, "raise VIML_NUM_KERNELS");

// Brackets one kernel launch with profiling events when enabled and counts it.
struct LaunchScope {
  viml_ctx* ctx;
  cudaEvent_t stop = nullptr;
  cudaStream_t on;
  LaunchScope(viml_ctx* c, int id, cudaStream_t st = nullptr) : ctx(c), on(st ? st : c->stream) {
    ctx->launches++;
    if (ctx->prof) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, on);
      ctx->prof_events[id].emplace_back(a, b);
      stop = b;
    }
  }
  ~LaunchScope() {
    if (stop) cudaEventRecord(stop, on);
  }
};

// Per-pose cache written by prep_windows_kernel and read by the factor kernels (42 doubles):
//   P[3]      position
//   R[9]      toRotationMatrix(q)               un-normalised (projection_factor.cpp:56-57)
//   Rinv[9]   toRotationMatrix(q.inverse())      == the map v -> Qj.inverse()*v (projection_factor.cpp:38)
//   M[9]      ric^T * R^T                        (projection_factor.cpp:81, :93)
//   Rl[9]     Ricn^T * Rn^T, Rn = toRotationMatrix(normalized(q))   (line_projection_factor.cpp:31-39)
//   tl[3]     -Rl*P - Ricn^T*Tic                 (line_projection_factor.cpp:40)
// Per-window extrinsic cache (24 doubles): tic[3], ric[9], ricinv[9], rtt[3] = ric^T*tic
constexpr int kMapTile = 64;   // lines per tile of the Morton-sorted map copy (measured on the cfg-3 sweep: 256 -> 0.234 ms, 64 -> 0.207, 32 -> 0.307)
constexpr int kPoseCache = 42;
constexpr int kExCache = 24;
constexpr int PC_P = 0, PC_R = 3, PC_RINV = 12, PC_M = 21, PC_RL = 30, PC_TL = 39;
constexpr int EC_TIC = 0, EC_RIC = 3, EC_RICINV = 12, EC_RTT = 21;

struct LinearizeArgs {  // device pointers only
  int W, P, F, D;
  int64_t NP, NL;          // factors covered by this launch
  int64_t pf_begin, lf_begin;  // first factor covered (chunked launches view a window range of one batch)
  int64_t NL_stride;       // plane stride of lf_geom (total line factors of the batch)
  const double* poses;
  const double* ex_pose;
  const double* inv_depth;
  const int32_t* pf_window_offset;
  const uint32_t* pf_idx;
  const double* pf_obs;
  const double* pf_pts_i_z;
  const int32_t* lf_window_offset;
  const int32_t* lf_frame;
  const double* lf_geom;
  double* cache;  // [W][P*kPoseCache + kExCache]
  viml_linearize_out out;
  double sqrt_info, cauchy_a, inv_cauchy_a2, fx, fy, cx, cy;
  uint32_t flags;
};

struct DenseArgs {  // device pointers only; viml_dense_factors
  int X;          // extra tangent columns behind the D pose / extrinsic columns
  int64_t ND;
  const int32_t* window_offset;
  const int64_t* row_offset;
  const int64_t* col_offset;
  const int64_t* jac_offset;
  const int32_t* col_index;
  const double* residual;
  const double* jacobian;
};

// linearize_kernels.cu
int viml_launch_linearize(viml_ctx* ctx, const LinearizeArgs& a);
int viml_launch_prep(viml_ctx* ctx, const LinearizeArgs& a);   // pose cache of the state in `a` only
// observation table -> per-factor pairs for the windows / factor range of `a` (a.pf_obs is the destination, absolute factor ids)
int viml_launch_expand_obs(viml_ctx* ctx, const LinearizeArgs& a, const void* feat_obs, const void* pf_obs_j, bool f32);
// line table -> the nine lf_geom planes for the line-factor range of `a` (a.lf_geom is the destination, absolute factor ids)
int viml_launch_expand_lines(viml_ctx* ctx, const LinearizeArgs& a, const int32_t* lf_map_index, const float* lf_seg2d);
int viml_launch_reduced(viml_ctx* ctx, int W, int D, const DenseArgs& dn, const double* S, const double* g, double* Sx, double* gx);
int viml_launch_gn_reduced_solve(viml_ctx* ctx, int W, int D, const DenseArgs& dn, const double* S, const double* g, double lambda,
                                 double* dx, int32_t* solved, double* cost);
int viml_launch_gn_solve(viml_ctx* ctx, int W, int Dx, double lambda, const double* Sx, const double* gx, double* dx, int32_t* solved,
                         double* cost);
int viml_launch_gn_update(viml_ctx* ctx, const LinearizeArgs& a, int X, const double* extra_in, const double* dx, const int32_t* solved,
                          double* o_poses, double* o_ex, double* o_dep, double* o_extra);
int viml_launch_cost(viml_ctx* ctx, const LinearizeArgs& a, const DenseArgs& dn, const double* dx, int slot, double* cost);
// schur_kernels.cu
int viml_launch_schur(viml_ctx* ctx, int W, int F, int D, const double* H_pp, const double* H_lp, const double* H_ll,
                      const double* b_p, const double* b_l, double* S, double* g, double eps);
double viml_dmma_peak_tflops(viml_ctx* ctx);
// marg_kernels.cu
int viml_launch_marginalize(viml_ctx* ctx, int K, int pos, int m, double eps, const double* A, const double* b,
                            double* A_schur, double* b_schur, double* lin_jac, double* lin_res);

struct AssocArgs {  // device pointers only
  int Pq, L;
  int64_t N;
  const double* map;  // [6][N] original order
  const double* map_sorted;  // [6][N] Morton order
  const int32_t* map_orig;   // [N]
  const double* tile_sphere; // [n_tiles][4]
  const double* group_sphere;// [ceil(n_tiles / 16)][4]
  int64_t n_tiles;
  const double* cull_poses;
  const double* match_poses;  // may alias cull_poses
  const double* ex_pose;
  const double* cull_ex_pose;  // may alias ex_pose
  const double* lines2d;
  const int32_t* n_lines2d;  // nullable
  int32_t* match_index;
  float* err;
  double* projected;
  int32_t* fov_count;   // always valid (scratch if the caller did not ask)
  int32_t* fov_index;   // nullable
  int32_t fov_capacity;
  uint32_t* fov_mask;   // always valid
  int64_t words;        // ceil(N/32)
  unsigned long long* stats;  // {gate tests, gated, overlap-scored, distance-scored}
  bool cached;                // VIML_FOV_CACHED: fov_mask / fov_count are given (cached lists), no cull
};
// associate_kernels.cu (compiled with -fmad=false)
int viml_launch_associate(viml_ctx* ctx, const AssocArgs& a);
// The same in two phases, so that a host can pipeline pose chunks without a synchronisation per chunk: phase 1 (camera poses, FoV
// cull, list offsets for ALL poses of `a`; one device->host read of the offsets), phase 2 (list fill, projection, match) for the
// poses [p0, p1) of a view `v` of `a` whose per-pose pointers are advanced to pose p0.
struct AssocPlan {
  void *cull = nullptr, *match = nullptr;   // camera poses [Pq]
  int64_t* off = nullptr;                   // device list offsets [Pq + 1]
  std::vector<int64_t> h_off;               // the same on the host
  int32_t* list = nullptr;
  void *seg = nullptr, *abc = nullptr, *aux = nullptr, *dir = nullptr, *len = nullptr, *rec = nullptr, *flen = nullptr;
  int Pq = 0;
};
int viml_assoc_phase1(viml_ctx* ctx, const AssocArgs& a, AssocPlan* plan);
int viml_assoc_phase2(viml_ctx* ctx, const AssocArgs& v, const AssocPlan& plan, int p0, int p1);
// triangulate_kernels.cu
int viml_launch_triangulate(viml_ctx* ctx, int P, int64_t NF, const double* poses, const double* ex, const int32_t* fwin,
                            const int32_t* start, const int64_t* off, const double* pts, double init_depth, double* depth);
int viml_launch_track_gate(viml_ctx* ctx, int n_tracks, const int32_t* track_offset, const int32_t* line_index, uint8_t* credible_line,
                           uint8_t* credible_matching);
int viml_launch_divcheck(viml_ctx* ctx, const double* a, const double* b, int64_t n, unsigned long long* mismatches);
