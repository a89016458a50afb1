// viml_api.cu — the C-ABI of include/viml.h: context lifetime, host<->device staging, entry points.
// No CPU fallback anywhere: every compute path ends in a kernel launch on the context's device.
#include <dlfcn.h>

#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <ctime>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace {

thread_local std::string g_create_error;

// c* = min{ c in [0,1] : acos(c) <= angle_th } with the host libm (the reference's acos, estimator.cpp:608).
double cos_threshold(double angle_th) {
  if (!(std::acos(1.0) <= angle_th)) return INFINITY;
  if (std::acos(0.0) <= angle_th) return 0.0;
  uint64_t lo, hi;
  double dlo = 0.0, dhi = 1.0;
  std::memcpy(&lo, &dlo, 8);
  std::memcpy(&hi, &dhi, 8);
  while (hi - lo > 1) {
    const uint64_t mid = lo + (hi - lo) / 2;
    double dm;
    std::memcpy(&dm, &mid, 8);
    if (std::acos(dm) <= angle_th) hi = mid; else lo = mid;
  }
  double r;
  std::memcpy(&r, &hi, 8);
  return r;
}

int fail(viml_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

}  // namespace

static void drop_prof_events(viml_ctx* ctx) {
  for (auto& v : ctx->prof_events) {
    for (auto& pr : v) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    v.clear();
  }
}

extern "C" {

int viml_abi_version(void) { return VIML_ABI_VERSION; }

int viml_create(viml_ctx** out, const viml_config* cfg, int device) {
  if (!out || !cfg) return VIML_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return VIML_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VIML_ERR_NO_DEVICE;
  if (prop.major != 10) return VIML_ERR_NO_DEVICE;  // the library only carries sm_100a code
  if (cudaSetDevice(device) != cudaSuccess) return VIML_ERR_NO_DEVICE;
  viml_ctx* ctx = new viml_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cfg = *cfg;
  ctx->cos_th = cos_threshold(cfg->angle_th);
  ctx->nan_angle_passes = !(3.1415926 > cfg->angle_th);
  if (const char* e = getenv("VIML_FORCE_GENERIC")) ctx->force_generic = e[0] == '1';  // test hook
  if (const char* e = getenv("VIML_BRUTE_CULL")) ctx->brute_cull = e[0] == '1';
  if (const char* e = getenv("VIML_SCHUR_SPLITK")) ctx->schur_splitk = e[0] == '1';
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return VIML_ERR_CUDA;
  }
  *out = ctx;
  return VIML_OK;
}

void viml_destroy(viml_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  drop_prof_events(ctx);
  ctx->in_arena.release();
  ctx->out_arena.release();
  if (ctx->h_stage_in) cudaFreeHost(ctx->h_stage_in);
  if (ctx->h_stage_out) cudaFreeHost(ctx->h_stage_out);
  ctx->scratch.release();
  ctx->scratch2.release();
  ctx->scratch3.release();
  ctx->gn_in.release();
  ctx->gn_out.release();
  ctx->s_full.release();
  if (ctx->d_map) cudaFree(ctx->d_map);
  if (ctx->d_map_sorted) cudaFree(ctx->d_map_sorted);
  if (ctx->d_map_orig) cudaFree(ctx->d_map_orig);
  if (ctx->d_tile_sphere) cudaFree(ctx->d_tile_sphere);
  if (ctx->d_group_sphere) cudaFree(ctx->d_group_sphere);
  if (ctx->d_assoc_stats) cudaFree(ctx->d_assoc_stats);
  if (ctx->d_fov_slots) cudaFree(ctx->d_fov_slots);
  if (ctx->d_fov_slot_count) cudaFree(ctx->d_fov_slot_count);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->copy_stream2) cudaStreamDestroy(ctx->copy_stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->nccl_lib) dlclose(ctx->nccl_lib);
  delete ctx;
}

const char* viml_last_error(const viml_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int viml_sync(viml_ctx* ctx) {
  if (!ctx) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VIML_OK;
}

void* viml_stream(viml_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int viml_host_alloc(void** p, size_t bytes) {
  if (!p) return VIML_ERR_INVALID;
  return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? VIML_OK : VIML_ERR_CUDA;
}
int viml_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? VIML_OK : VIML_ERR_CUDA; }

int viml_device_alloc(viml_ctx* ctx, void** p, size_t bytes) {
  if (!ctx || !p) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  VIML_TRY_CUDA(ctx, cudaMalloc(p, bytes));
  return VIML_OK;
}
int viml_device_free(viml_ctx* ctx, void* p) {
  if (!ctx) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  VIML_TRY_CUDA(ctx, cudaFree(p));
  return VIML_OK;
}
int viml_memcpy_h2d(viml_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return VIML_OK;
}
int viml_memcpy_d2h(viml_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return VIML_OK;
}

int64_t viml_kernel_launches(const viml_ctx* ctx) { return ctx ? ctx->launches : 0; }


int viml_profile_begin(viml_ctx* ctx) {
  if (!ctx) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  drop_prof_events(ctx);
  ctx->prof = true;
  return VIML_OK;
}

int viml_profile_end(viml_ctx* ctx, double* ms, int64_t* n) {
  if (!ctx) return VIML_ERR_INVALID;
  ctx->prof = false;
  VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < VIML_NUM_KERNELS; ++k) {
    double tot = 0.0;
    for (auto& pr : ctx->prof_events[k]) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) tot += t;
    }
    if (ms) ms[k] = tot;
    if (n) n[k] = (int64_t)ctx->prof_events[k].size();
  }
  drop_prof_events(ctx);
  return VIML_OK;
}

const char* viml_kernel_name(int id) {
  static const char* names[VIML_NUM_KERNELS] = {"prep_windows", "linearize_points", "linearize_lines", "assemble_hb",
                                                "schur_landmarks", "assoc_cam_pose", "assoc_cull", "assoc_scan",
                                                "assoc_fill_list", "assoc_project", "assoc_match", "marginalize_dense",
                                                "microbench", "plan_windows", "assemble_irregular", "gn_step"};
  return (id >= 0 && id < VIML_NUM_KERNELS) ? names[id] : "";
}

// ---- map -------------------------------------------------------------------------------------
int viml_set_map(viml_ctx* ctx, const double* lines, int64_t n) {
  if (!ctx || n < 0 || (n > 0 && !lines)) return VIML_ERR_INVALID;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  auto drop = [&](auto*& ptr) {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
  };
  drop(ctx->d_map), drop(ctx->d_map_sorted), drop(ctx->d_map_orig), drop(ctx->d_tile_sphere), drop(ctx->d_group_sphere);
  drop(ctx->d_fov_slots), drop(ctx->d_fov_slot_count);   // the cached FoV lists index the old map
  ctx->fov_words = 0;
  ctx->n_map = 0, ctx->n_tiles = 0;
  ctx->map_set = true;
  if (n == 0) return VIML_OK;
  // AoS rows [sx sy sz ex ey ez] (parameters.cpp:50-59) -> six SoA planes
  std::vector<double> soa((size_t)6 * n);
  for (int64_t j = 0; j < n; ++j)
    for (int c = 0; c < 6; ++c) soa[(size_t)c * n + j] = lines[6 * j + c];
  VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_map, soa.size() * sizeof(double)));
  VIML_TRY_CUDA(ctx, cudaMemcpy(ctx->d_map, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
  // Morton order of the segment mid-points (21 bits per axis over the bounding box); one-time ingest cost.
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t j = 0; j < n; ++j)
    for (int c = 0; c < 3; ++c) {
      const double m = 0.5 * (lines[6 * j + c] + lines[6 * j + 3 + c]);
      if (std::isfinite(m)) lo[c] = std::min(lo[c], m), hi[c] = std::max(hi[c], m);
    }
  auto spread = [](uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
  };
  // one scale for the three axes: cubic cells, so a flat map (x, y extent >> z extent) is split in x and y first
  double span = 0.0;
  for (int c = 0; c < 3; ++c)
    if (hi[c] > lo[c]) span = std::max(span, hi[c] - lo[c]);
  if (!(span > 0.0)) span = 1.0;
  std::vector<uint64_t> key(n);
  for (int64_t j = 0; j < n; ++j) {
    uint64_t k = 0;
    for (int c = 0; c < 3; ++c) {
      const double m = 0.5 * (lines[6 * j + c] + lines[6 * j + 3 + c]);
      double t = std::isfinite(m) ? (m - lo[c]) / span : 0.0;
      t = std::min(1.0, std::max(0.0, t));
      k |= spread((uint64_t)(t * 2097151.0)) << c;
    }
    key[j] = k;
  }
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
  std::vector<double> sorted((size_t)6 * n);
  for (int64_t k = 0; k < n; ++k)
    for (int c = 0; c < 6; ++c) sorted[(size_t)c * n + k] = lines[6 * (int64_t)order[k] + c];
  const int64_t nt = (n + kMapTile - 1) / kMapTile;
  std::vector<double> sph((size_t)nt * 4);
  for (int64_t t = 0; t < nt; ++t) {
    const int64_t k0 = t * kMapTile, k1 = std::min<int64_t>(n, k0 + kMapTile);
    double blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool finite = true;
    for (int64_t k = k0; k < k1; ++k)
      for (int c = 0; c < 6; ++c) {
        const double v = sorted[(size_t)c * n + k];
        finite &= std::isfinite(v);
        blo[c % 3] = std::min(blo[c % 3], v), bhi[c % 3] = std::max(bhi[c % 3], v);
      }
    double ctr[3], r2 = 0.0;
    for (int c = 0; c < 3; ++c) ctr[c] = 0.5 * (blo[c] + bhi[c]);
    for (int64_t k = k0; k < k1; ++k)
      for (int e = 0; e < 2; ++e) {
        double d2 = 0.0;
        for (int c = 0; c < 3; ++c) {
          const double d = sorted[(size_t)(3 * e + c) * n + k] - ctr[c];
          d2 += d * d;
        }
        r2 = std::max(r2, d2);
      }
    // a tile with a non-finite coordinate is never rejected (radius = +inf): the exact test decides
    sph[4 * t] = ctr[0], sph[4 * t + 1] = ctr[1], sph[4 * t + 2] = ctr[2];
    sph[4 * t + 3] = finite ? std::sqrt(r2) * (1.0 + 1e-12) : INFINITY;
  }
  VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_map_sorted, sorted.size() * sizeof(double)));
  VIML_TRY_CUDA(ctx, cudaMemcpy(ctx->d_map_sorted, sorted.data(), sorted.size() * sizeof(double), cudaMemcpyHostToDevice));
  VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_map_orig, (size_t)n * sizeof(int32_t)));
  VIML_TRY_CUDA(ctx, cudaMemcpy(ctx->d_map_orig, order.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
  VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_tile_sphere, sph.size() * sizeof(double)));
  VIML_TRY_CUDA(ctx, cudaMemcpy(ctx->d_tile_sphere, sph.data(), sph.size() * sizeof(double), cudaMemcpyHostToDevice));
  // second level: one sphere around every 16 consecutive tile spheres (consecutive Morton tiles are neighbours)
  const int64_t ngrp = (nt + 15) / 16;
  std::vector<double> gsp((size_t)std::max<int64_t>(ngrp, 1) * 4, 0.0);
  for (int64_t gq = 0; gq < ngrp; ++gq) {
    const int64_t ta = gq * 16, tb = std::min<int64_t>(nt, ta + 16);
    double c[3] = {0, 0, 0}, r = 0.0;
    bool fin = true;
    for (int64_t t = ta; t < tb; ++t) {
      fin &= std::isfinite(sph[4 * t + 3]) && std::isfinite(sph[4 * t]) && std::isfinite(sph[4 * t + 1]) && std::isfinite(sph[4 * t + 2]);
      for (int k = 0; k < 3; ++k) c[k] += sph[4 * t + k] / (double)(tb - ta);
    }
    for (int64_t t = ta; t < tb && fin; ++t) {
      double d2 = 0.0;
      for (int k = 0; k < 3; ++k) d2 += (sph[4 * t + k] - c[k]) * (sph[4 * t + k] - c[k]);
      r = std::max(r, std::sqrt(d2) * (1.0 + 1e-12) + sph[4 * t + 3]);
    }
    gsp[4 * gq] = fin ? c[0] : 0.0, gsp[4 * gq + 1] = fin ? c[1] : 0.0, gsp[4 * gq + 2] = fin ? c[2] : 0.0;
    gsp[4 * gq + 3] = fin ? r * (1.0 + 1e-12) : INFINITY;   // never rejected when a member is not finite
  }
  VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_group_sphere, gsp.size() * sizeof(double)));
  VIML_TRY_CUDA(ctx, cudaMemcpy(ctx->d_group_sphere, gsp.data(), gsp.size() * sizeof(double), cudaMemcpyHostToDevice));
  ctx->n_map = n;
  ctx->n_tiles = nt;
  return VIML_OK;
}

// ---- linearisation -----------------------------------------------------------------------------
int viml_linearize_batch(viml_ctx* ctx, const viml_window_batch* in, const viml_linearize_out* out, uint32_t flags) {
  if (!ctx) return VIML_ERR_INVALID;
  if (!in || !out) return fail(ctx, VIML_ERR_INVALID, "null batch or output struct");
  const int W = in->n_windows, P = in->poses_per_window, F = in->feats_per_window;
  const int64_t NP = in->n_point_factors, NL = in->n_line_factors;
  if (W < 0 || P < 1 || P > 255 || F < 0 || F > 65535 || NP < 0 || NL < 0)
    return fail(ctx, VIML_ERR_INVALID, "window batch sizes out of range");
  if (!(flags & (VIML_OUT_RESIDUAL_JACOBIAN | VIML_OUT_HB | VIML_OUT_SCHUR)))
    return fail(ctx, VIML_ERR_INVALID, "no output mode requested");
  if (W == 0) return VIML_OK;
  const bool obs_f32 = !in->pf_obs && !(in->feat_obs && in->pf_obs_j) && in->feat_obs_f32 && in->pf_obs_j_f32;
  const bool obs_table = !in->pf_obs && ((in->feat_obs && in->pf_obs_j) || obs_f32);   // observations as a per-feature table
  const void* h_fobs = obs_f32 ? (const void*)in->feat_obs_f32 : (const void*)in->feat_obs;
  const void* h_obsj = obs_f32 ? (const void*)in->pf_obs_j_f32 : (const void*)in->pf_obs_j;
  const size_t ob = obs_f32 ? 8 : 16;   // bytes of one {x, y}
  const bool line_table = NL > 0 && !in->lf_geom && in->lf_map_index && in->lf_seg2d_f32;   // line factors as (map line, 2D segment)
  if (!in->poses || !in->ex_pose || (F > 0 && !in->inv_depth) || !in->pf_window_offset ||
      (NP > 0 && (!in->pf_idx || (!in->pf_obs && !obs_table))) ||
      (NL > 0 && (!in->lf_window_offset || !in->lf_frame || (!in->lf_geom && !line_table))))
    return fail(ctx, VIML_ERR_INVALID, "null input array");
  if (line_table && (!ctx->map_set || ctx->n_map <= 0))
    return fail(ctx, VIML_ERR_NOMAP, "line factors given by map index need a map (viml_set_map)");
  const int D = 6 * (P + 1);
  timespec ts_begin;
  clock_gettime(CLOCK_MONOTONIC, &ts_begin);
  const bool dev = (flags & VIML_PTRS_DEVICE) != 0;
  const bool wantA = (flags & VIML_OUT_RESIDUAL_JACOBIAN) != 0;
  const bool wantHB = (flags & VIML_OUT_HB) != 0, wantS = (flags & VIML_OUT_SCHUR) != 0;
  if (wantHB && !(out->H_pp && out->H_lp && out->H_ll && out->b_p && out->b_l))
    return fail(ctx, VIML_ERR_INVALID, "VIML_OUT_HB needs H_pp, H_lp, H_ll, b_p and b_l");
  if (wantS && !(out->S && out->g)) return fail(ctx, VIML_ERR_INVALID, "VIML_OUT_SCHUR needs S and g");
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;

  LinearizeArgs a{};
  a.W = W, a.P = P, a.F = F, a.D = D, a.NP = NP, a.NL = NL;
  a.pf_begin = 0, a.lf_begin = 0, a.NL_stride = NL;
  a.sqrt_info = ctx->cfg.sqrt_info, a.cauchy_a = ctx->cfg.cauchy_a;
  a.inv_cauchy_a2 = 1.0 / (ctx->cfg.cauchy_a * ctx->cfg.cauchy_a);
  a.fx = ctx->cfg.fx, a.fy = ctx->cfg.fy, a.cx = ctx->cfg.cx, a.cy = ctx->cfg.cy;
  a.flags = flags;

  const size_t n_pose = (size_t)W * P * 7, n_ex = (size_t)W * 7, n_dep = (size_t)W * F;
  const size_t szHpp = (size_t)W * D * D, szHlp = (size_t)W * F * D, szF = (size_t)W * F, szD = (size_t)W * D;
  // scratch: pose cache (+ H blocks when only the Schur complement is wanted)
  size_t scratch_bytes = DeviceArena::padded((size_t)W * (P * kPoseCache + kExCache) * sizeof(double)) +
                         DeviceArena::padded((size_t)W * sizeof(int));
  const bool scratchH = wantS && (!wantHB || !dev);
  if (scratchH && dev)
    scratch_bytes += DeviceArena::padded(szHpp * 8) + DeviceArena::padded(szHlp * 8) + 2 * DeviceArena::padded(szF * 8) +
                     DeviceArena::padded(szD * 8);
  VIML_TRY_CUDA(ctx, ctx->scratch.reserve(scratch_bytes));
  a.cache = ctx->scratch.take<double>((size_t)W * (P * kPoseCache + kExCache));

  if (dev) {
    a.poses = in->poses, a.ex_pose = in->ex_pose, a.inv_depth = in->inv_depth;
    a.pf_window_offset = in->pf_window_offset, a.pf_idx = in->pf_idx, a.pf_obs = in->pf_obs;
    a.pf_pts_i_z = in->pf_pts_i_z;
    a.lf_window_offset = in->lf_window_offset, a.lf_frame = in->lf_frame, a.lf_geom = in->lf_geom;
    a.out = *out;
    if ((obs_table && NP > 0) || line_table) {
      VIML_TRY_CUDA(ctx, ctx->in_arena.reserve(DeviceArena::padded((size_t)NP * 32) + DeviceArena::padded((size_t)NL * 72)));
      if (obs_table && NP > 0) {
        a.pf_obs = ctx->in_arena.take<double>((size_t)NP * 4);
        const int rc = viml_launch_expand_obs(ctx, a, h_fobs, h_obsj, obs_f32);
        if (rc != VIML_OK) return rc;
      }
      if (line_table) {
        a.lf_geom = ctx->in_arena.take<double>((size_t)NL * 9);
        const int rc = viml_launch_expand_lines(ctx, a, in->lf_map_index, in->lf_seg2d_f32);
        if (rc != VIML_OK) return rc;
      }
    }
    if (!wantA) a.out.pf_residual = a.out.pf_jac_pose_i = a.out.pf_jac_pose_j = a.out.pf_jac_ex = a.out.pf_jac_feat =
                    a.out.lf_residual = a.out.lf_jac_pose = nullptr;
    if (wantS && !wantHB) {
      a.out.H_pp = ctx->scratch.take<double>(szHpp);
      a.out.H_lp = ctx->scratch.take<double>(szHlp);
      a.out.H_ll = ctx->scratch.take<double>(szF);
      a.out.b_l = ctx->scratch.take<double>(szF);
      a.out.b_p = ctx->scratch.take<double>(szD);
    }
    return viml_launch_linearize(ctx, a);
  }

  // host pointers: the packed indices and CSR offsets are validated before anything is launched (device pointers cannot
  // be read here; on that path plan_kernel routes windows with bad indices to the generic kernel, which drops them)
  {
    const int32_t* po = in->pf_window_offset;
    bool ok = po[0] == 0 && po[W] == NP;
    for (int w = 0; w < W && ok; ++w) ok = po[w] <= po[w + 1];
    if (NL > 0) {
      const int32_t* lo = in->lf_window_offset;
      ok = ok && lo[0] == 0 && lo[W] == NL;
      for (int w = 0; w < W && ok; ++w) ok = lo[w] <= lo[w + 1];
    }
    if (!ok) return fail(ctx, VIML_ERR_INVALID, "window offsets are not a CSR of the factor arrays (offset[0] == 0, non-decreasing, offset[W] == n_factors)");
  }
  // packed indices: checked chunk by chunk inside the pipeline below (the host runs ahead of the asynchronous copies and
  // kernels, so the ~1 ms scan of a 2.3 M-factor batch hides behind the transfers of the previous chunk)
  auto indices_ok = [&](int64_t pa, int64_t pb, int64_t la, int64_t lb) {
    uint32_t bad = 0;
    const uint32_t uP = (uint32_t)P, uF = (uint32_t)F;
    for (int64_t k = pa; k < pb; ++k) {
      const uint32_t ix = in->pf_idx[k];
      bad |= (uint32_t)((ix & 0xffu) >= uP) | (uint32_t)(((ix >> 8) & 0xffu) >= uP) | (uint32_t)((ix >> 16) >= uF);
    }
    for (int64_t k = la; k < lb; ++k) bad |= (uint32_t)((uint32_t)in->lf_frame[k] >= uP);
    if (line_table)
      for (int64_t k = la; k < lb; ++k) bad |= (uint32_t)((uint64_t)(int64_t)in->lf_map_index[k] >= (uint64_t)ctx->n_map);
    return bad == 0;
  };
  // ---- host pointers: chunked pipeline  H2D(c+1) | kernels(c) | D2H(c-1)  on three streams ----
  // Device buffers hold the whole batch at the same absolute positions as the host arrays; a chunk is a range of
  // windows, its kernels run on a "view" (pointers advanced to the first window of the chunk, factor ranges
  // absolute).  PCIe is full duplex, so with pinned host memory the call approaches max(H2D, D2H) time.
  auto pad = [](size_t b) { return DeviceArena::padded(b); };
  size_t in_bytes = pad(n_pose * 8) + pad(n_ex * 8) + pad(n_dep * 8) + 2 * pad((size_t)(W + 1) * 4);
  in_bytes += pad((size_t)NP * 4) + pad((size_t)NP * 32) + pad((size_t)NP * 8);
  in_bytes += pad((size_t)NL * 4) + pad((size_t)NL * 72);
  if (obs_table) in_bytes += pad((size_t)NP * ob) + pad(n_dep * ob);
  if (line_table) in_bytes += pad((size_t)NL * 4) + pad((size_t)NL * 16);
  VIML_TRY_CUDA(ctx, ctx->in_arena.reserve(in_bytes));
  double* d_poses = ctx->in_arena.take<double>(n_pose);
  double* d_ex = ctx->in_arena.take<double>(n_ex);
  double* d_dep = ctx->in_arena.take<double>(n_dep);
  int32_t* d_poff = ctx->in_arena.take<int32_t>((size_t)W + 1);
  uint32_t* d_idx = ctx->in_arena.take<uint32_t>((size_t)NP);
  double* d_obs = ctx->in_arena.take<double>((size_t)NP * 4);
  double* d_z = in->pf_pts_i_z ? ctx->in_arena.take<double>((size_t)NP) : nullptr;
  int32_t* d_loff = NL > 0 ? ctx->in_arena.take<int32_t>((size_t)W + 1) : nullptr;
  int32_t* d_frame = NL > 0 ? ctx->in_arena.take<int32_t>((size_t)NL) : nullptr;
  double* d_geom = NL > 0 ? ctx->in_arena.take<double>((size_t)NL * 9) : nullptr;
  char* d_obsj = obs_table ? ctx->in_arena.take<char>((size_t)NP * ob) : nullptr;
  char* d_fobs = obs_table ? ctx->in_arena.take<char>(n_dep * ob) : nullptr;
  int32_t* d_lidx = line_table ? ctx->in_arena.take<int32_t>((size_t)NL) : nullptr;
  float* d_lseg = line_table ? ctx->in_arena.take<float>((size_t)NL * 4) : nullptr;
  a.poses = d_poses, a.ex_pose = d_ex, a.inv_depth = d_dep, a.pf_window_offset = d_poff, a.pf_idx = d_idx;
  a.pf_obs = d_obs, a.pf_pts_i_z = d_z, a.lf_window_offset = d_loff, a.lf_frame = d_frame, a.lf_geom = d_geom;

  struct Slot { double** dev; double* host; size_t per_window, per_pf, per_lf; };
  std::vector<Slot> slots;
  size_t out_bytes = 0;
  auto want = [&](double** dslot, double* host, size_t per_window, size_t per_pf, size_t per_lf, bool force) {
    if (host || force) {
      slots.push_back({dslot, host, per_window, per_pf, per_lf});
      out_bytes += pad((per_window * W + per_pf * NP + per_lf * NL) * 8);
    }
  };
  const bool needH = wantHB || wantS;
  if (wantA) {
    want(&a.out.pf_residual, out->pf_residual, 0, 2, 0, false);
    want(&a.out.pf_jac_pose_i, out->pf_jac_pose_i, 0, 14, 0, false);
    want(&a.out.pf_jac_pose_j, out->pf_jac_pose_j, 0, 14, 0, false);
    want(&a.out.pf_jac_ex, out->pf_jac_ex, 0, 14, 0, false);
    want(&a.out.pf_jac_feat, out->pf_jac_feat, 0, 2, 0, false);
    want(&a.out.lf_residual, out->lf_residual, 0, 0, 2, false);
    want(&a.out.lf_jac_pose, out->lf_jac_pose, 0, 0, 14, false);
  }
  if (needH) {
    want(&a.out.H_pp, wantHB ? out->H_pp : nullptr, (size_t)D * D, 0, 0, true);
    want(&a.out.H_lp, wantHB ? out->H_lp : nullptr, (size_t)F * D, 0, 0, true);
    want(&a.out.H_ll, wantHB ? out->H_ll : nullptr, (size_t)F, 0, 0, true);
    want(&a.out.b_p, wantHB ? out->b_p : nullptr, (size_t)D, 0, 0, true);
    want(&a.out.b_l, wantHB ? out->b_l : nullptr, (size_t)F, 0, 0, true);
  }
  if (wantS) {
    want(&a.out.S, out->S, (flags & VIML_S_PACKED) ? (size_t)D * (D + 1) / 2 : (size_t)D * D, 0, 0, true);
    want(&a.out.g, out->g, (size_t)D, 0, 0, true);
  }
  VIML_TRY_CUDA(ctx, ctx->out_arena.reserve(out_bytes));
  for (auto& sl : slots) *sl.dev = ctx->out_arena.take<double>(sl.per_window * W + sl.per_pf * NP + sl.per_lf * NL);

  // ---- small batches (a live sliding window: BASELINE configs[0]): latency, not bandwidth.  Every cudaMemcpyAsync costs
  // 5-10 us of driver and DMA set-up whatever its size (more from pageable memory), and this call has ~10 input and up to 14
  // output arrays: the inputs are gathered into ONE pinned block laid out like the device arena and go up in one copy, the
  // outputs come back in one copy and are handed out from the pinned block.
  if (W < 512 && ctx->in_arena.used + ctx->out_arena.used <= ((size_t)4 << 20) && !getenv("VIML_NO_STAGING")) {
    auto grow = [&](char** buf, size_t* cap, size_t need) -> cudaError_t {
      if (need <= *cap) return cudaSuccess;
      if (*buf) cudaFreeHost(*buf);
      *buf = nullptr, *cap = 0;
      const cudaError_t e = cudaHostAlloc((void**)buf, need + need / 2 + 4096, cudaHostAllocDefault);
      if (e == cudaSuccess) *cap = need + need / 2 + 4096;
      return e;
    };
    VIML_TRY_CUDA(ctx, grow(&ctx->h_stage_in, &ctx->h_stage_in_cap, ctx->in_arena.used));
    VIML_TRY_CUDA(ctx, grow(&ctx->h_stage_out, &ctx->h_stage_out_cap, ctx->out_arena.used));
    if (!indices_ok(0, NP, 0, NL))
      return fail(ctx, VIML_ERR_INVALID, "factor index out of range (pose index >= poses_per_window, feature >= feats_per_window, line frame >= poses_per_window or map index outside the map)");
    char* const ibase = ctx->in_arena.base;
    auto put = [&](const void* dptr, const void* src, size_t bytes) {
      if (bytes) memcpy(ctx->h_stage_in + ((const char*)dptr - ibase), src, bytes);
    };
    put(d_poses, in->poses, n_pose * 8), put(d_ex, in->ex_pose, n_ex * 8), put(d_dep, in->inv_depth, n_dep * 8);
    put(d_poff, in->pf_window_offset, (size_t)(W + 1) * 4), put(d_idx, in->pf_idx, (size_t)NP * 4);
    if (obs_table) put(d_obsj, h_obsj, (size_t)NP * ob), put(d_fobs, h_fobs, n_dep * ob);
    else put(d_obs, in->pf_obs, (size_t)NP * 32);
    if (d_z) put(d_z, in->pf_pts_i_z, (size_t)NP * 8);
    if (NL > 0) {
      put(d_loff, in->lf_window_offset, (size_t)(W + 1) * 4), put(d_frame, in->lf_frame, (size_t)NL * 4);
      if (line_table) put(d_lidx, in->lf_map_index, (size_t)NL * 4), put(d_lseg, in->lf_seg2d_f32, (size_t)NL * 16);
      else put(d_geom, in->lf_geom, (size_t)NL * 72);
    }
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(ibase, ctx->h_stage_in, ctx->in_arena.used, cudaMemcpyHostToDevice, st));
    int rc = obs_table ? viml_launch_expand_obs(ctx, a, d_fobs, d_obsj, obs_f32) : VIML_OK;
    if (rc == VIML_OK && line_table) rc = viml_launch_expand_lines(ctx, a, d_lidx, d_lseg);
    if (rc == VIML_OK) rc = viml_launch_linearize(ctx, a);
    if (rc != VIML_OK) return rc;
    // the device range that holds every slot the caller wants
    char *lo = nullptr, *hi = nullptr;
    for (auto& sl : slots)
      if (sl.host) {
        char* b = (char*)*sl.dev;
        char* e = b + (sl.per_window * W + sl.per_pf * NP + sl.per_lf * NL) * 8;
        lo = lo ? std::min(lo, b) : b, hi = hi ? std::max(hi, e) : e;
      }
    if (lo) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage_out, lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, st));
    VIML_TRY_CUDA(ctx, cudaStreamSynchronize(st));
    for (auto& sl : slots)
      if (sl.host) memcpy(sl.host, ctx->h_stage_out + ((char*)*sl.dev - lo), (sl.per_window * W + sl.per_pf * NP + sl.per_lf * NL) * 8);
    VIML_TRY_CUDA(ctx, cudaGetLastError());
    return VIML_OK;
  }

  cudaStream_t s_in = ctx->copy_stream, s_out = ctx->copy_stream2;
  cudaEvent_t tv0 = nullptr;
  if (getenv("VIML_E2E_TIMING")) {
    cudaEventCreate(&tv0);
    cudaEventRecord(tv0, s_in);
  }
  // offsets first (tiny, needed by every chunk); the previous call's work on `st` is already complete (synchronous API)
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(d_poff, in->pf_window_offset, (size_t)(W + 1) * 4, cudaMemcpyHostToDevice, s_in));
  if (NL > 0) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(d_loff, in->lf_window_offset, (size_t)(W + 1) * 4, cudaMemcpyHostToDevice, s_in));
  int nchunk = W >= 512 ? 4 : 1;   // measured on the 4096-window batch with the float32 table (75 MB up, 88 MB down): 2, 3, 4, 5, 6, 8, 12, 16 chunks -> 2.96, 2.81, 2.67, 2.75, 2.70, 2.79, 2.93, 2.96 ms
  if (const char* e = getenv("VIML_CHUNKS")) nchunk = std::max(1, std::min(W, atoi(e)));   // tuning hook
  std::vector<int> wt(nchunk, 1);       // relative chunk sizes
  if (const char* e = getenv("VIML_CHUNK_WEIGHTS")) {   // tuning hook: e.g. "1,2,4,4,5" (also sets the chunk count)
    std::vector<int> v;
    for (const char* q = e; *q;) {
      v.push_back(std::max(1, atoi(q)));
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
    if (!v.empty() && (int)v.size() <= W) wt = v, nchunk = (int)v.size();
  }
  std::vector<int> wb(nchunk + 1, 0);
  {
    int64_t tot = 0, run = 0;
    for (int x : wt) tot += x;
    for (int c = 0; c < nchunk; ++c) run += wt[c], wb[c + 1] = (int)((int64_t)W * run / tot);
  }
  // small per-window arrays: whole batch, one copy each
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(d_poses, in->poses, n_pose * 8, cudaMemcpyHostToDevice, s_in));
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(d_ex, in->ex_pose, n_ex * 8, cudaMemcpyHostToDevice, s_in));
  if (n_dep) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(d_dep, in->inv_depth, n_dep * 8, cudaMemcpyHostToDevice, s_in));
  if (NL > 0) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(d_frame, in->lf_frame, (size_t)NL * 4, cudaMemcpyHostToDevice, s_in));
  std::vector<cudaEvent_t> ev_in(nchunk), ev_k(nchunk);
  for (int c = 0; c < nchunk; ++c) {
    cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_k[c], cudaEventDisableTiming);
  }
  int rc = VIML_OK;
  auto h2d = [&](void* d, const void* h, size_t bytes) {
    if (bytes) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s_in);
  };
  // every upload is queued first (the copies depend on nothing but the caller's arrays), so the upload stream runs at link
  // speed while the host validates the indices of chunk c and launches its kernels
  for (int c = 0; c < nchunk; ++c) {
    const int w0 = wb[c], w1 = wb[c + 1];
    const int64_t pa = in->pf_window_offset[w0], pb = in->pf_window_offset[w1];
    const int64_t la = NL > 0 ? in->lf_window_offset[w0] : 0, lb = NL > 0 ? in->lf_window_offset[w1] : 0;
    // per chunk only the three large arrays (observations, packed indices, line geometry as ONE strided copy of its nine
    // planes); the small per-window arrays went up for the whole batch before the loop — every copy costs ~10-20 us of
    // DMA set-up however small it is, and 17 copies per chunk kept the upload stream busy for 6 ms on 125 MB
    h2d(d_idx + pa, in->pf_idx + pa, (size_t)(pb - pa) * 4);
    if (obs_table) {
      h2d(d_obsj + ob * pa, (const char*)h_obsj + ob * pa, (size_t)(pb - pa) * ob);
      h2d(d_fobs + (size_t)w0 * F * ob, (const char*)h_fobs + (size_t)w0 * F * ob, (size_t)(w1 - w0) * F * ob);
    } else {
      h2d(d_obs + 4 * pa, in->pf_obs + 4 * pa, (size_t)(pb - pa) * 32);
    }
    if (d_z) h2d(d_z + pa, in->pf_pts_i_z + pa, (size_t)(pb - pa) * 8);
    if (lb > la && line_table) {
      h2d(d_lidx + la, in->lf_map_index + la, (size_t)(lb - la) * 4);
      h2d(d_lseg + 4 * la, in->lf_seg2d_f32 + 4 * la, (size_t)(lb - la) * 16);
    } else if (lb > la) {
      cudaMemcpy2DAsync(d_geom + la, (size_t)NL * 8, in->lf_geom + la, (size_t)NL * 8, (size_t)(lb - la) * 8, 9,
                        cudaMemcpyHostToDevice, s_in);
    }
    cudaEventRecord(ev_in[c], s_in);
  }
  for (int c = 0; c < nchunk && rc == VIML_OK; ++c) {
    const int w0 = wb[c], w1 = wb[c + 1], Wc = w1 - w0;
    const int64_t pa = in->pf_window_offset[w0], pb = in->pf_window_offset[w1];
    const int64_t la = NL > 0 ? in->lf_window_offset[w0] : 0, lb = NL > 0 ? in->lf_window_offset[w1] : 0;
    if (!indices_ok(pa, pb, la, lb)) {   // bad indices never reach a kernel: this chunk and the following ones are not launched
      rc = fail(ctx, VIML_ERR_INVALID, "factor index out of range (pose index >= poses_per_window, feature >= feats_per_window, line frame >= poses_per_window or map index outside the map)");
      break;
    }
    cudaStreamWaitEvent(st, ev_in[c], 0);
    // view of windows [w0, w1)
    LinearizeArgs v = a;
    v.W = Wc, v.NP = pb - pa, v.NL = lb - la, v.pf_begin = pa, v.lf_begin = la;
    v.poses = a.poses + (size_t)w0 * P * 7, v.ex_pose = a.ex_pose + (size_t)w0 * 7, v.inv_depth = a.inv_depth + (size_t)w0 * F;
    v.pf_window_offset = a.pf_window_offset + w0;
    if (NL > 0) v.lf_window_offset = a.lf_window_offset + w0;
    v.cache = a.cache + (size_t)w0 * (P * kPoseCache + kExCache);
    for (auto& sl : slots) {
      double** vd = (double**)((char*)&v.out + ((char*)sl.dev - (char*)&a.out));
      *vd = *sl.dev + sl.per_window * w0;   // per-factor arrays stay absolute (indexed by the global factor id)
    }
    if (Wc > 0 && obs_table) rc = viml_launch_expand_obs(ctx, v, d_fobs + (size_t)w0 * F * ob, d_obsj, obs_f32);
    if (Wc > 0 && rc == VIML_OK && line_table) rc = viml_launch_expand_lines(ctx, v, d_lidx, d_lseg);
    if (Wc > 0 && rc == VIML_OK) rc = viml_launch_linearize(ctx, v);
    cudaEventRecord(ev_k[c], st);
    cudaStreamWaitEvent(s_out, ev_k[c], 0);
    for (auto& sl : slots)
      if (sl.host) {
        const size_t off = sl.per_window * w0 + sl.per_pf * pa + sl.per_lf * la;
        const size_t cnt = sl.per_window * Wc + sl.per_pf * (pb - pa) + sl.per_lf * (lb - la);
        if (cnt) cudaMemcpyAsync(sl.host + off, *sl.dev + off, cnt * 8, cudaMemcpyDeviceToHost, s_out);
      }
  }
  const bool e2e_timing = getenv("VIML_E2E_TIMING") != nullptr;
  timespec ts_q;
  cudaEvent_t tv[3] = {nullptr, nullptr, nullptr};
  if (e2e_timing) {
    clock_gettime(CLOCK_MONOTONIC, &ts_q);
    for (auto& e : tv) cudaEventCreate(&e);
    cudaEventRecord(tv[0], s_in), cudaEventRecord(tv[1], st), cudaEventRecord(tv[2], s_out);
  }
  cudaError_t e1 = cudaStreamSynchronize(s_out), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(s_in);
  if (e2e_timing) {   // diagnosis: how long the host took to queue the pipeline vs how long the device needed to drain it
    timespec ts_d;
    clock_gettime(CLOCK_MONOTONIC, &ts_d);
    float a = 0, b = 0, c2 = 0;
    cudaEventElapsedTime(&a, tv0, tv[0]), cudaEventElapsedTime(&b, tv0, tv[1]), cudaEventElapsedTime(&c2, tv0, tv[2]);
    fprintf(stderr, "[viml e2e] device: uploads done %.3f ms, kernels done %.3f ms, downloads done %.3f ms after the first upload was queued\n", a, b, c2);
    for (auto& e : tv) cudaEventDestroy(e);
    cudaEventDestroy(tv0);
    fprintf(stderr, "[viml e2e] queued after %.3f ms, drained after %.3f ms (%d chunks)\n",
            (ts_q.tv_sec - ts_begin.tv_sec) * 1e3 + (ts_q.tv_nsec - ts_begin.tv_nsec) * 1e-6,
            (ts_d.tv_sec - ts_begin.tv_sec) * 1e3 + (ts_d.tv_nsec - ts_begin.tv_nsec) * 1e-6, nchunk);
  }
  for (int c = 0; c < nchunk; ++c) cudaEventDestroy(ev_in[c]), cudaEventDestroy(ev_k[c]);
  if (rc != VIML_OK) return rc;
  VIML_TRY_CUDA(ctx, e1);
  VIML_TRY_CUDA(ctx, e2);
  VIML_TRY_CUDA(ctx, e3);
  VIML_TRY_CUDA(ctx, cudaGetLastError());
  return VIML_OK;
}

// ---- dense marginalisation -----------------------------------------------------------------------
int viml_marginalize_batch(viml_ctx* ctx, const viml_marg_batch* in, const viml_marg_out* out, uint32_t flags) {
  if (!ctx) return VIML_ERR_INVALID;
  if (!in || !out) return fail(ctx, VIML_ERR_INVALID, "null marg batch");
  const int K = in->n_problems, pos = in->pos, m = in->m, n = pos - m;
  if (K < 0 || pos < 1 || pos > 256 || m < 0 || n < 1 || !in->A || !in->b)
    return fail(ctx, VIML_ERR_INVALID, "marg batch sizes out of range (pos <= 256, n >= 1)");
  if (K == 0) return VIML_OK;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (flags & VIML_PTRS_DEVICE)
    return viml_launch_marginalize(ctx, K, pos, m, in->eps, in->A, in->b, out->A_schur, out->b_schur,
                                   out->linearized_jacobians, out->linearized_residuals);
  const size_t szA = (size_t)K * pos * pos, szb = (size_t)K * pos, szS = (size_t)K * n * n, szs = (size_t)K * n;
  auto pad = [](size_t b) { return DeviceArena::padded(b); };
  VIML_TRY_CUDA(ctx, ctx->in_arena.reserve(pad(szA * 8) + pad(szb * 8)));
  double* dA = ctx->in_arena.take<double>(szA);
  double* db = ctx->in_arena.take<double>(szb);
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(dA, in->A, szA * 8, cudaMemcpyHostToDevice, st));
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(db, in->b, szb * 8, cudaMemcpyHostToDevice, st));
  VIML_TRY_CUDA(ctx, ctx->out_arena.reserve(2 * pad(szS * 8) + 2 * pad(szs * 8)));
  double* dS = ctx->out_arena.take<double>(szS);
  double* ds = ctx->out_arena.take<double>(szs);
  double* dJ = ctx->out_arena.take<double>(szS);
  double* dr = ctx->out_arena.take<double>(szs);
  const bool lin = out->linearized_jacobians || out->linearized_residuals;
  int rc = viml_launch_marginalize(ctx, K, pos, m, in->eps, dA, db, dS, ds, lin ? dJ : nullptr, lin ? dr : nullptr);
  if (rc != VIML_OK) return rc;
  if (out->A_schur) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(out->A_schur, dS, szS * 8, cudaMemcpyDeviceToHost, st));
  if (out->b_schur) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(out->b_schur, ds, szs * 8, cudaMemcpyDeviceToHost, st));
  if (out->linearized_jacobians)
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(out->linearized_jacobians, dJ, szS * 8, cudaMemcpyDeviceToHost, st));
  if (out->linearized_residuals)
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(out->linearized_residuals, dr, szs * 8, cudaMemcpyDeviceToHost, st));
  VIML_TRY_CUDA(ctx, cudaStreamSynchronize(st));
  return VIML_OK;
}

// ---- association -----------------------------------------------------------------------------------
int viml_line_associate(viml_ctx* ctx, const viml_assoc_query* q, const viml_assoc_out* out, uint32_t flags) {
  if (!ctx) return VIML_ERR_INVALID;
  if (!q || !out) return fail(ctx, VIML_ERR_INVALID, "null query or output struct");
  if (!ctx->map_set) return fail(ctx, VIML_ERR_NOMAP, "viml_line_associate before viml_set_map (an empty map is legal, but it must be set)");
  const int Pq = q->n_poses, L = q->lines_per_pose;
  const bool cached = (flags & VIML_FOV_CACHED) != 0;
  if (Pq < 0 || L < 0 || (!cached && !q->cull_poses) || !q->ex_pose || (L > 0 && !q->lines2d))
    return fail(ctx, VIML_ERR_INVALID, "bad association query");
  if (cached) {
    if (Pq > VIML_FOV_SLOTS && !q->fov_slot) return fail(ctx, VIML_ERR_INVALID, "VIML_FOV_CACHED: more poses than window slots and no fov_slot map");
    if (!q->match_poses && !q->cull_poses) return fail(ctx, VIML_ERR_INVALID, "VIML_FOV_CACHED needs match_poses");
    for (int p = 0; p < Pq && q->fov_slot; ++p)
      if (q->fov_slot[p] < 0 || q->fov_slot[p] >= VIML_FOV_SLOTS) return fail(ctx, VIML_ERR_INVALID, "fov_slot out of range");
  }
  if (Pq == 0) return VIML_OK;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t N = ctx->n_map, words = (N + 31) / 32;
  const bool dev = (flags & VIML_PTRS_DEVICE) != 0;
  AssocArgs a{};
  a.Pq = Pq, a.L = L, a.N = N, a.map = ctx->d_map, a.words = words;
  a.cached = cached;
  // cached lists: the slots' mask rows and counts gathered into [Pq][words] / [Pq] behind the pose order of this query
  uint32_t* cmask = nullptr;
  int32_t* ccount = nullptr;
  if (cached) {
    VIML_TRY_CUDA(ctx, ctx->s_full.reserve(DeviceArena::padded((size_t)Pq * words * 4 + 4) + DeviceArena::padded((size_t)Pq * 4)));
    cmask = ctx->s_full.take<uint32_t>((size_t)Pq * words + 1);
    ccount = ctx->s_full.take<int32_t>((size_t)Pq);
    for (int p = 0; p < Pq; ++p) {
      const int sl = q->fov_slot ? q->fov_slot[p] : p;
      if (ctx->d_fov_slots && ctx->fov_words == words) {
        if (words) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(cmask + (size_t)p * words, ctx->d_fov_slots + (size_t)sl * words, (size_t)words * 4, cudaMemcpyDeviceToDevice, st));
        VIML_TRY_CUDA(ctx, cudaMemcpyAsync(ccount + p, ctx->d_fov_slot_count + sl, 4, cudaMemcpyDeviceToDevice, st));
      } else {   // no slot was ever updated for this map: empty lists
        if (words) VIML_TRY_CUDA(ctx, cudaMemsetAsync(cmask + (size_t)p * words, 0, (size_t)words * 4, st));
        VIML_TRY_CUDA(ctx, cudaMemsetAsync(ccount + p, 0, 4, st));
      }
    }
  }
  if (!ctx->d_assoc_stats) VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_assoc_stats, 32));
  VIML_TRY_CUDA(ctx, cudaMemsetAsync(ctx->d_assoc_stats, 0, 32, st));
  a.stats = ctx->d_assoc_stats;
  a.map_sorted = ctx->d_map_sorted, a.map_orig = ctx->d_map_orig, a.tile_sphere = ctx->d_tile_sphere, a.group_sphere = ctx->d_group_sphere, a.n_tiles = ctx->n_tiles;
  a.fov_capacity = out->fov_index ? out->fov_capacity : 0;
  auto pad = [](size_t b) { return DeviceArena::padded(b); };
  const size_t nq = (size_t)Pq * L;
  if (dev) {
    size_t sb = 0;
    if (!out->fov_mask) sb += pad((size_t)Pq * words * 4);
    if (!out->fov_count) sb += pad((size_t)Pq * 4);
    VIML_TRY_CUDA(ctx, ctx->out_arena.reserve(sb));
    a.match_poses = q->match_poses ? q->match_poses : q->cull_poses;
    a.cull_poses = q->cull_poses ? q->cull_poses : a.match_poses;
    a.ex_pose = q->ex_pose, a.lines2d = q->lines2d, a.n_lines2d = q->n_lines2d;
    a.cull_ex_pose = q->cull_ex_pose ? q->cull_ex_pose : q->ex_pose;
    a.match_index = out->match_index, a.err = out->err, a.projected = out->projected;
    a.fov_index = out->fov_index;
    a.fov_mask = out->fov_mask ? out->fov_mask : ctx->out_arena.take<uint32_t>((size_t)Pq * words);
    a.fov_count = out->fov_count ? out->fov_count : ctx->out_arena.take<int32_t>(Pq);
    if (cached) {
      if (words) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(a.fov_mask, cmask, (size_t)Pq * words * 4, cudaMemcpyDeviceToDevice, st));
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(a.fov_count, ccount, (size_t)Pq * 4, cudaMemcpyDeviceToDevice, st));
    }
    return viml_launch_associate(ctx, a);
  }
  // Host buffers: poses are independent, so a large query is run as a pipeline of pose chunks -- all uploads are queued
  // on the copy stream up front (one event per chunk), chunk c's kernels wait for its upload only, and its results go
  // back on the second copy stream while chunk c + 1 computes.
  VIML_TRY_CUDA(ctx, ctx->in_arena.reserve(4 * pad((size_t)Pq * 56) + pad(nq * 32) + pad((size_t)Pq * 4)));
  cudaStream_t cs = ctx->copy_stream, ds = ctx->copy_stream2;
  int C = Pq >= 512 ? 4 : 1;   // measured on the cfg-3 sweep: 2, 4, 6, 8, 12, 16 chunks -> 2.32, 2.19, 2.30, 2.36, 2.79, 2.76 ms
  if (const char* e = getenv("VIML_ASSOC_CHUNKS")) C = std::max(1, std::min(Pq, atoi(e)));   // tuning hook
  auto up = [&](const void* src, size_t bytes) -> void* {
    char* d = ctx->in_arena.take<char>(bytes);
    if (bytes) cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, cs);
    return d;
  };
  VIML_TRY_CUDA(ctx, cudaEventRecord(ctx->ev_a, st));   // the copy stream starts after whatever the caller queued before
  VIML_TRY_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_a, 0));
  a.cull_poses = (const double*)up(q->cull_poses ? q->cull_poses : q->match_poses, (size_t)Pq * 56);
  a.match_poses = q->match_poses ? (const double*)up(q->match_poses, (size_t)Pq * 56) : a.cull_poses;
  a.ex_pose = (const double*)up(q->ex_pose, (size_t)Pq * 56);
  a.cull_ex_pose = q->cull_ex_pose ? (const double*)up(q->cull_ex_pose, (size_t)Pq * 56) : a.ex_pose;
  a.n_lines2d = q->n_lines2d ? (const int32_t*)up(q->n_lines2d, (size_t)Pq * 4) : nullptr;
  double* d_l2d = ctx->in_arena.take<double>(nq * 4);
  a.lines2d = d_l2d;
  const size_t cap = a.fov_capacity;
  VIML_TRY_CUDA(ctx, ctx->out_arena.reserve(pad(nq * 4) + pad(nq * 12) + pad(nq * 32) + pad((size_t)Pq * 4) +
                                            pad((size_t)Pq * cap * 4) + pad((size_t)Pq * words * 4)));
  a.match_index = ctx->out_arena.take<int32_t>(nq);
  a.err = ctx->out_arena.take<float>(nq * 3);
  a.projected = ctx->out_arena.take<double>(nq * 4);
  a.fov_count = ctx->out_arena.take<int32_t>(Pq);
  a.fov_index = cap ? ctx->out_arena.take<int32_t>((size_t)Pq * cap) : nullptr;
  a.fov_mask = ctx->out_arena.take<uint32_t>((size_t)Pq * words);
  if (cached) {
    if (words) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(a.fov_mask, cmask, (size_t)Pq * words * 4, cudaMemcpyDeviceToDevice, st));
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(a.fov_count, ccount, (size_t)Pq * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (out->fov_index && cap)  // entries past fov_count keep the caller's content
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(a.fov_index, out->fov_index, (size_t)Pq * cap * 4, cudaMemcpyHostToDevice, cs));
  if (out->match_index && q->n_lines2d) {  // ragged queries past n_lines2d keep the caller's content
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(a.match_index, out->match_index, nq * 4, cudaMemcpyHostToDevice, cs));
    if (out->err) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(a.err, out->err, nq * 12, cudaMemcpyHostToDevice, cs));
    if (out->projected) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(a.projected, out->projected, nq * 32, cudaMemcpyHostToDevice, cs));
  }
  // phase 1 (camera poses, FoV cull, list offsets) runs once for all poses as soon as the small arrays are up — it does not need
  // the detected lines, whose upload it overlaps — and holds the call's only synchronisation; phase 2 then runs chunk by chunk
  VIML_TRY_CUDA(ctx, cudaEventRecord(ctx->ev_b, cs));
  VIML_TRY_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_b, 0));
  std::vector<cudaEvent_t> ev_in(C), ev_k(C);
  auto chunk = [&](int c, int& p0, int& p1) { p0 = (int)((int64_t)Pq * c / C), p1 = (int)((int64_t)Pq * (c + 1) / C); };
  for (int c = 0; c < C; ++c) {
    int p0, p1;
    chunk(c, p0, p1);
    cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_k[c], cudaEventDisableTiming);
    const size_t o = (size_t)p0 * L * 4, n = (size_t)(p1 - p0) * L * 4;
    if (n) cudaMemcpyAsync(d_l2d + o, q->lines2d + o, n * 8, cudaMemcpyHostToDevice, cs);
    cudaEventRecord(ev_in[c], cs);
  }
  AssocPlan plan;
  int rc = viml_assoc_phase1(ctx, a, &plan);
  auto down = [&](void* host, const void* d, size_t bytes) {
    if (host && bytes) cudaMemcpyAsync(host, d, bytes, cudaMemcpyDeviceToHost, ds);
  };
  for (int c = 0; c < C && rc == VIML_OK; ++c) {
    int p0, p1;
    chunk(c, p0, p1);
    cudaStreamWaitEvent(st, ev_in[c], 0);
    AssocArgs v = a;   // view of poses [p0, p1)
    v.Pq = p1 - p0;
    v.cull_poses = a.cull_poses + (size_t)p0 * 7, v.match_poses = a.match_poses + (size_t)p0 * 7;
    v.ex_pose = a.ex_pose + (size_t)p0 * 7, v.cull_ex_pose = a.cull_ex_pose + (size_t)p0 * 7;
    v.lines2d = a.lines2d + (size_t)p0 * L * 4, v.n_lines2d = a.n_lines2d ? a.n_lines2d + p0 : nullptr;
    v.match_index = a.match_index + (size_t)p0 * L, v.err = a.err + (size_t)p0 * L * 3, v.projected = a.projected + (size_t)p0 * L * 4;
    v.fov_count = a.fov_count + p0, v.fov_index = a.fov_index ? a.fov_index + (size_t)p0 * cap : nullptr;
    v.fov_mask = a.fov_mask + (size_t)p0 * words;
    if (v.Pq > 0) rc = viml_assoc_phase2(ctx, v, plan, p0, p1);
    cudaEventRecord(ev_k[c], st);
    cudaStreamWaitEvent(ds, ev_k[c], 0);
    const size_t np = (size_t)(p1 - p0), nqc = np * L;
    down(out->match_index ? out->match_index + (size_t)p0 * L : nullptr, v.match_index, nqc * 4);
    down(out->err ? out->err + (size_t)p0 * L * 3 : nullptr, v.err, nqc * 12);
    down(out->projected ? out->projected + (size_t)p0 * L * 4 : nullptr, v.projected, nqc * 32);
    down(out->fov_count ? out->fov_count + p0 : nullptr, v.fov_count, np * 4);
    down(out->fov_index ? out->fov_index + (size_t)p0 * cap : nullptr, v.fov_index, np * cap * 4);
    down(out->fov_mask ? out->fov_mask + (size_t)p0 * words : nullptr, v.fov_mask, np * words * 4);
  }
  cudaError_t e1 = cudaStreamSynchronize(ds), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(cs);
  for (int c = 0; c < C; ++c) cudaEventDestroy(ev_in[c]), cudaEventDestroy(ev_k[c]);
  if (rc != VIML_OK) return rc;
  VIML_TRY_CUDA(ctx, e1);
  VIML_TRY_CUDA(ctx, e2);
  VIML_TRY_CUDA(ctx, e3);
  VIML_TRY_CUDA(ctx, cudaGetLastError());
  return VIML_OK;
}

// ---- FoV cache ---------------------------------------------------------------------------------------
static int fov_cache_ready(viml_ctx* ctx) {
  const int64_t words = (ctx->n_map + 31) / 32;
  if (ctx->d_fov_slots && ctx->fov_words == words) return VIML_OK;
  if (ctx->d_fov_slots) cudaFree(ctx->d_fov_slots);
  ctx->d_fov_slots = nullptr;
  VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_fov_slots, (size_t)VIML_FOV_SLOTS * std::max<int64_t>(words, 1) * 4));
  if (!ctx->d_fov_slot_count) VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_fov_slot_count, VIML_FOV_SLOTS * 4));
  VIML_TRY_CUDA(ctx, cudaMemsetAsync(ctx->d_fov_slots, 0, (size_t)VIML_FOV_SLOTS * std::max<int64_t>(words, 1) * 4, ctx->stream));
  VIML_TRY_CUDA(ctx, cudaMemsetAsync(ctx->d_fov_slot_count, 0, VIML_FOV_SLOTS * 4, ctx->stream));
  ctx->fov_words = words;
  return VIML_OK;
}

int viml_fov_update(viml_ctx* ctx, int32_t slot, const double* pose, const double* ex_pose, int32_t* count) {
  if (!ctx) return VIML_ERR_INVALID;
  if (slot < 0 || slot >= VIML_FOV_SLOTS || !pose || !ex_pose) return fail(ctx, VIML_ERR_INVALID, "viml_fov_update: bad slot or null pose");
  if (!ctx->map_set) return fail(ctx, VIML_ERR_NOMAP, "viml_fov_update before viml_set_map");
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = fov_cache_ready(ctx);
  if (rc != VIML_OK) return rc;
  cudaStream_t st = ctx->stream;
  VIML_TRY_CUDA(ctx, ctx->in_arena.reserve(2 * DeviceArena::padded(56)));
  double* dp = ctx->in_arena.take<double>(7);
  double* de = ctx->in_arena.take<double>(7);
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(dp, pose, 56, cudaMemcpyHostToDevice, st));
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(de, ex_pose, 56, cudaMemcpyHostToDevice, st));
  AssocArgs a{};
  a.Pq = 1, a.L = 0, a.N = ctx->n_map, a.map = ctx->d_map, a.words = ctx->fov_words;
  a.map_sorted = ctx->d_map_sorted, a.map_orig = ctx->d_map_orig, a.tile_sphere = ctx->d_tile_sphere, a.group_sphere = ctx->d_group_sphere, a.n_tiles = ctx->n_tiles;
  if (!ctx->d_assoc_stats) VIML_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_assoc_stats, 32));
  a.stats = ctx->d_assoc_stats;
  a.cull_poses = a.match_poses = dp, a.ex_pose = a.cull_ex_pose = de;
  a.fov_mask = ctx->d_fov_slots + (size_t)slot * ctx->fov_words;
  a.fov_count = ctx->d_fov_slot_count + slot;
  rc = viml_launch_associate(ctx, a);
  if (rc != VIML_OK) return rc;
  if (count) {
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(count, a.fov_count, 4, cudaMemcpyDeviceToHost, st));
    VIML_TRY_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return VIML_OK;
}

int viml_fov_slide(viml_ctx* ctx, int32_t marginalize_old) {
  if (!ctx) return VIML_ERR_INVALID;
  if (!ctx->map_set) return fail(ctx, VIML_ERR_NOMAP, "viml_fov_slide before viml_set_map");
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  const int rc = fov_cache_ready(ctx);
  if (rc != VIML_OK) return rc;
  cudaStream_t st = ctx->stream;
  const size_t wb = (size_t)ctx->fov_words * 4;
  const int last = VIML_FOV_SLOTS - 1;
  auto row = [&](int s) { return ctx->d_fov_slots + (size_t)s * ctx->fov_words; };
  if (marginalize_old) {
    // estimator.cpp:2148 swaps slot i with i+1 for i = 0..WINDOW_SIZE-1 and :2160 then copies slot WINDOW_SIZE-1 into the
    // newest one: net effect slot i <- old slot i+1 (i < WINDOW_SIZE), newest slot unchanged.  Stream-ordered copies,
    // ascending, each reading a row that has not been overwritten yet.
    for (int i = 0; i < last; ++i) {
      if (wb) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(row(i), row(i + 1), wb, cudaMemcpyDeviceToDevice, st));
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(ctx->d_fov_slot_count + i, ctx->d_fov_slot_count + i + 1, 4, cudaMemcpyDeviceToDevice, st));
    }
  } else {
    if (wb) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(row(last - 1), row(last), wb, cudaMemcpyDeviceToDevice, st));   // :2218
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(ctx->d_fov_slot_count + last - 1, ctx->d_fov_slot_count + last, 4, cudaMemcpyDeviceToDevice, st));
  }
  return VIML_OK;
}

int viml_track_gate(viml_ctx* ctx, int32_t n_tracks, const int32_t* track_offset, const int32_t* line_index, uint8_t* credible_line,
                    uint8_t* credible_matching, uint32_t flags) {
  if (!ctx) return VIML_ERR_INVALID;
  if (n_tracks < 0 || (n_tracks > 0 && (!track_offset || !credible_matching))) return fail(ctx, VIML_ERR_INVALID, "viml_track_gate: bad arguments");
  if (!ctx->map_set) return fail(ctx, VIML_ERR_NOMAP, "viml_track_gate before viml_set_map");
  if (n_tracks == 0) return VIML_OK;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (flags & VIML_PTRS_DEVICE) return viml_launch_track_gate(ctx, n_tracks, track_offset, line_index, credible_line, credible_matching);
  const int64_t nobs = track_offset[n_tracks];
  if (nobs < 0 || (nobs > 0 && (!line_index || !credible_line))) return fail(ctx, VIML_ERR_INVALID, "viml_track_gate: null observation arrays");
  auto pad = [](size_t b) { return DeviceArena::padded(b); };
  VIML_TRY_CUDA(ctx, ctx->in_arena.reserve(pad((size_t)(n_tracks + 1) * 4) + pad((size_t)nobs * 4 + 4)));
  VIML_TRY_CUDA(ctx, ctx->out_arena.reserve(pad((size_t)nobs + 1) + pad((size_t)n_tracks)));
  int32_t* doff = ctx->in_arena.take<int32_t>((size_t)n_tracks + 1);
  int32_t* didx = ctx->in_arena.take<int32_t>((size_t)nobs + 1);
  uint8_t* dcl = ctx->out_arena.take<uint8_t>((size_t)nobs + 1);
  uint8_t* dcm = ctx->out_arena.take<uint8_t>((size_t)n_tracks);
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(doff, track_offset, (size_t)(n_tracks + 1) * 4, cudaMemcpyHostToDevice, st));
  if (nobs) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(didx, line_index, (size_t)nobs * 4, cudaMemcpyHostToDevice, st));
  const int rc = viml_launch_track_gate(ctx, n_tracks, doff, didx, dcl, dcm);
  if (rc != VIML_OK) return rc;
  if (nobs) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(credible_line, dcl, (size_t)nobs, cudaMemcpyDeviceToHost, st));
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(credible_matching, dcm, (size_t)n_tracks, cudaMemcpyDeviceToHost, st));
  VIML_TRY_CUDA(ctx, cudaStreamSynchronize(st));
  return VIML_OK;
}

int viml_selftest_division(viml_ctx* ctx, const double* a, const double* b, int64_t n, int64_t* mismatches) {
  if (!ctx) return VIML_ERR_INVALID;
  if (n < 0 || (n > 0 && (!a || !b)) || !mismatches) return fail(ctx, VIML_ERR_INVALID, "bad division self-test arguments");
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  VIML_TRY_CUDA(ctx, ctx->in_arena.reserve(2 * DeviceArena::padded((size_t)n * 8) + 256));
  double* da = ctx->in_arena.take<double>((size_t)n);
  double* db = ctx->in_arena.take<double>((size_t)n);
  unsigned long long* dm = ctx->in_arena.take<unsigned long long>(1);
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(da, a, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(db, b, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  VIML_TRY_CUDA(ctx, cudaMemsetAsync(dm, 0, 8, ctx->stream));
  const int rc = viml_launch_divcheck(ctx, da, db, n, dm);
  if (rc != VIML_OK) return rc;
  unsigned long long h = 0;
  VIML_TRY_CUDA(ctx, cudaMemcpyAsync(&h, dm, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *mismatches = (int64_t)h;
  return VIML_OK;
}

int viml_assoc_stats(viml_ctx* ctx, int64_t* gate_tests, int64_t* gated, int64_t* overlap_scored,
                     int64_t* distance_scored) {
  if (!ctx) return VIML_ERR_INVALID;
  unsigned long long h[4] = {0, 0, 0, 0};
  if (ctx->d_assoc_stats) {
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_assoc_stats, 32, cudaMemcpyDeviceToHost, ctx->stream));
    VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (gate_tests) *gate_tests = (int64_t)h[0];
  if (gated) *gated = (int64_t)h[1];
  if (overlap_scored) *overlap_scored = (int64_t)h[2];
  if (distance_scored) *distance_scored = (int64_t)h[3];
  return VIML_OK;
}

// ---- multi-GPU: partial H/b all-reduce for the single huge window ----------------------------------
int viml_allreduce_hb(viml_ctx* ctx, void* nccl_comm, double* buf, int64_t count) {
  if (!ctx) return VIML_ERR_INVALID;
  if (!nccl_comm || !buf || count < 0) return fail(ctx, VIML_ERR_INVALID, "bad all-reduce arguments");
  if (!ctx->nccl_lib) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nme : names) {
      ctx->nccl_lib = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
      if (ctx->nccl_lib) break;
    }
    if (!ctx->nccl_lib) return fail(ctx, VIML_ERR_UNSUPPORTED, "libnccl.so.2 not loadable");
  }
  // ncclResult_t ncclAllReduce(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
  using allreduce_t = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  auto fn = (allreduce_t)dlsym(ctx->nccl_lib, "ncclAllReduce");
  if (!fn) return fail(ctx, VIML_ERR_UNSUPPORTED, "ncclAllReduce not found");
  const int ncclFloat64 = 8, ncclSum = 0;
  const int rc = fn(buf, buf, (size_t)count, ncclFloat64, ncclSum, nccl_comm, ctx->stream);
  if (rc != 0) return fail(ctx, VIML_ERR_CUDA, "ncclAllReduce failed");
  return VIML_OK;
}

}  // extern "C"
