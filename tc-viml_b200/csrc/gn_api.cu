// gn_api.cu — C-ABI entry points for the reduced system of the whole window (dense-block factors: prior, IMU) and one
// Gauss-Newton / Levenberg-Marquardt iteration on it, plus the line_3d.txt map reader.  See include/viml.h.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

int fail(viml_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

// Host <-> device staging for the synchronous host-pointer flavour of an entry point: inputs are copied into one arena,
// outputs live in another and are copied back by finish().  With VIML_PTRS_DEVICE the caller's pointers pass through.
struct Stager {
  viml_ctx* ctx;
  bool dev;
  struct In { const void* host; size_t bytes; const void** slot; };
  struct Out { void* host; size_t bytes; void** slot; };
  std::vector<In> ins;
  std::vector<Out> outs;
  template <class T>
  void in(const T* host, size_t count, const T** slot) {
    *slot = host;
    if (!dev && host && count) ins.push_back({host, count * sizeof(T), reinterpret_cast<const void**>(slot)});
  }
  // an output (or scratch when host == nullptr / device mode without a caller buffer)
  template <class T>
  void out(T* host, size_t count, T** slot) {
    *slot = host;
    if ((!dev || !host) && count) outs.push_back({dev ? nullptr : host, count * sizeof(T), reinterpret_cast<void**>(slot)});
  }
  int commit() {
    size_t ib = 0, ob = 0;
    for (auto& e : ins) ib += DeviceArena::padded(e.bytes);
    for (auto& e : outs) ob += DeviceArena::padded(e.bytes);
    VIML_TRY_CUDA(ctx, ctx->gn_in.reserve(ib));
    VIML_TRY_CUDA(ctx, ctx->gn_out.reserve(ob));
    for (auto& e : ins) {
      char* d = ctx->gn_in.take<char>(e.bytes);
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(d, e.host, e.bytes, cudaMemcpyHostToDevice, ctx->stream));
      *e.slot = d;
    }
    for (auto& e : outs) *e.slot = ctx->gn_out.take<char>(e.bytes);
    return VIML_OK;
  }
  int finish() {
    if (dev) return VIML_OK;
    for (auto& e : outs)
      if (e.host) VIML_TRY_CUDA(ctx, cudaMemcpyAsync(e.host, *e.slot, e.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VIML_OK;
  }
};

int check_batch(viml_ctx* ctx, const viml_window_batch* in) {
  if (!in) return fail(ctx, VIML_ERR_INVALID, "null window batch");
  const int W = in->n_windows, P = in->poses_per_window, F = in->feats_per_window;
  if (W < 0 || P < 1 || P > 255 || F < 0 || F > 65535 || in->n_point_factors < 0 || in->n_line_factors < 0)
    return fail(ctx, VIML_ERR_INVALID, "window batch sizes out of range");
  if (W > 0 && (!in->poses || !in->ex_pose || (F > 0 && !in->inv_depth) || !in->pf_window_offset ||
                (in->n_point_factors > 0 && (!in->pf_idx || (!in->pf_obs && !(in->feat_obs && in->pf_obs_j) && !(in->feat_obs_f32 && in->pf_obs_j_f32)))) ||
                (in->n_line_factors > 0 && (!in->lf_window_offset || !in->lf_frame || (!in->lf_geom && !(in->lf_map_index && in->lf_seg2d_f32))))))
    return fail(ctx, VIML_ERR_INVALID, "null input array");
  return VIML_OK;
}

// Shared core of viml_reduced_system / viml_gn_step.
int run(viml_ctx* ctx, const viml_window_batch* in, const viml_dense_factors* dense, const double* extra_state,
        const viml_gn_options* opt, const viml_reduced_out* rout, const viml_gn_out* gout, uint32_t flags) {
  int rc = check_batch(ctx, in);
  if (rc != VIML_OK) return rc;
  const int W = in->n_windows, P = in->poses_per_window, F = in->feats_per_window, D = 6 * (P + 1);
  const int64_t NP = in->n_point_factors, NL = in->n_line_factors;
  const int X = dense ? dense->extra_dim : 0, Dx = D + X;
  const int64_t ND = dense ? dense->n_factors : 0;
  if (X < 0 || ND < 0 || Dx > 232) return fail(ctx, VIML_ERR_INVALID, "dense factors: extra_dim out of range (D + extra_dim <= 232)");
  if (ND > 0 && (!dense->window_offset || !dense->row_offset || !dense->col_offset || !dense->jac_offset || !dense->col_index ||
                 !dense->residual || !dense->jacobian))
    return fail(ctx, VIML_ERR_INVALID, "dense factors: null array");
  const bool step = gout != nullptr;
  if (step && !(gout->poses && gout->ex_pose && (F == 0 || gout->inv_depth) && gout->cost && gout->solved))
    return fail(ctx, VIML_ERR_INVALID, "viml_gn_step needs poses, ex_pose, inv_depth, cost and solved outputs");
  if (step && X > 0 && !gout->extra) return fail(ctx, VIML_ERR_INVALID, "viml_gn_step: extra_dim > 0 needs the extra output");
  if (!step && !(rout && rout->Sx && rout->gx)) return fail(ctx, VIML_ERR_INVALID, "viml_reduced_system needs Sx and gx");
  if (W == 0) return VIML_OK;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool dev = (flags & VIML_PTRS_DEVICE) != 0;
  int64_t n_rows = 0, n_cols = 0, n_jac = 0;
  if (ND > 0) {
    if (dev) {   // totals are the last prefix entries: three 8-byte reads from the device
      int64_t t[3];
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(&t[0], dense->row_offset + ND, 8, cudaMemcpyDeviceToHost, ctx->stream));
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(&t[1], dense->col_offset + ND, 8, cudaMemcpyDeviceToHost, ctx->stream));
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(&t[2], dense->jac_offset + ND, 8, cudaMemcpyDeviceToHost, ctx->stream));
      VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      n_rows = t[0], n_cols = t[1], n_jac = t[2];
    } else {
      n_rows = dense->row_offset[ND], n_cols = dense->col_offset[ND], n_jac = dense->jac_offset[ND];
      if (dense->window_offset[0] != 0 || dense->window_offset[W] != ND)
        return fail(ctx, VIML_ERR_INVALID, "dense factors: window_offset is not a CSR of the factor list");
      for (int64_t k = 0; k < n_cols; ++k)
        if (dense->col_index[k] < 0 || dense->col_index[k] >= Dx)
          return fail(ctx, VIML_ERR_INVALID, "dense factors: col_index out of range");
      for (int64_t k = 0; k < ND; ++k) {
        const int64_t n = dense->row_offset[k + 1] - dense->row_offset[k], c = dense->col_offset[k + 1] - dense->col_offset[k];
        if (n < 0 || c < 0 || dense->jac_offset[k + 1] - dense->jac_offset[k] != n * c)
          return fail(ctx, VIML_ERR_INVALID, "dense factors: jac_offset does not match rows x columns");
      }
    }
  }

  Stager st{ctx, dev};
  LinearizeArgs a{};
  a.W = W, a.P = P, a.F = F, a.D = D, a.NP = NP, a.NL = NL, a.pf_begin = 0, a.lf_begin = 0, a.NL_stride = NL;
  a.sqrt_info = ctx->cfg.sqrt_info, a.cauchy_a = ctx->cfg.cauchy_a, a.inv_cauchy_a2 = 1.0 / (ctx->cfg.cauchy_a * ctx->cfg.cauchy_a);
  a.fx = ctx->cfg.fx, a.fy = ctx->cfg.fy, a.cx = ctx->cfg.cx, a.cy = ctx->cfg.cy;
  a.flags = (flags & VIML_LOSS_CAUCHY) | VIML_OUT_HB | VIML_OUT_SCHUR;
  st.in(in->poses, (size_t)W * P * 7, &a.poses);
  st.in(in->ex_pose, (size_t)W * 7, &a.ex_pose);
  st.in(in->inv_depth, (size_t)W * F, &a.inv_depth);
  st.in(in->pf_window_offset, (size_t)W + 1, &a.pf_window_offset);
  st.in(in->pf_idx, (size_t)NP, &a.pf_idx);
  const bool obs_table = !in->pf_obs && NP > 0;   // observations as a per-feature table (viml.h): expanded on the device
  const bool obs_f32 = obs_table && !(in->feat_obs && in->pf_obs_j);
  const double *d_fobs = nullptr, *d_obsj = nullptr;
  const float *d_fobs32 = nullptr, *d_obsj32 = nullptr;
  double* d_obs = nullptr;
  if (obs_f32) {
    st.in(in->feat_obs_f32, (size_t)W * F * 2, &d_fobs32);
    st.in(in->pf_obs_j_f32, (size_t)NP * 2, &d_obsj32);
  } else if (obs_table) {
    st.in(in->feat_obs, (size_t)W * F * 2, &d_fobs);
    st.in(in->pf_obs_j, (size_t)NP * 2, &d_obsj);
  } else {
    st.in(in->pf_obs, (size_t)NP * 4, &a.pf_obs);
  }
  st.in(in->pf_pts_i_z, in->pf_pts_i_z ? (size_t)NP : 0, &a.pf_pts_i_z);
  st.in(in->lf_window_offset, NL > 0 ? (size_t)W + 1 : 0, &a.lf_window_offset);
  st.in(in->lf_frame, (size_t)NL, &a.lf_frame);
  const bool line_table = NL > 0 && !in->lf_geom;   // line factors as (map line, 2D segment): expanded on the device
  const int32_t* d_lidx = nullptr;
  const float* d_lseg = nullptr;
  double* d_geom = nullptr;
  if (line_table) {
    if (!dev)
      for (int64_t k = 0; k < NL; ++k)
        if (in->lf_map_index[k] < 0 || in->lf_map_index[k] >= ctx->n_map) return fail(ctx, VIML_ERR_INVALID, "line factor: map index outside the map");
    st.in(in->lf_map_index, (size_t)NL, &d_lidx);
    st.in(in->lf_seg2d_f32, (size_t)NL * 4, &d_lseg);
  } else {
    st.in(in->lf_geom, (size_t)NL * 9, &a.lf_geom);
  }
  const double* d_extra = nullptr;
  st.in(extra_state, extra_state ? (size_t)W * X : 0, &d_extra);
  DenseArgs dn{};
  dn.X = X, dn.ND = ND;
  if (ND > 0) {
    st.in(dense->window_offset, (size_t)W + 1, &dn.window_offset);
    st.in(dense->row_offset, (size_t)ND + 1, &dn.row_offset);
    st.in(dense->col_offset, (size_t)ND + 1, &dn.col_offset);
    st.in(dense->jac_offset, (size_t)ND + 1, &dn.jac_offset);
    st.in(dense->col_index, (size_t)n_cols, &dn.col_index);
    st.in(dense->residual, (size_t)n_rows, &dn.residual);
    st.in(dense->jacobian, (size_t)n_jac, &dn.jacobian);
  }
  // scratch (host == nullptr) and outputs
  double *cache = nullptr, *Sx = nullptr, *gx = nullptr, *dxv = nullptr, *cost = nullptr;
  double *o_poses = nullptr, *o_ex = nullptr, *o_dep = nullptr, *o_extra = nullptr;
  int32_t* solved = nullptr;
  st.out((double*)nullptr, (size_t)W * (P * kPoseCache + kExCache), &cache);
  if (obs_table) st.out((double*)nullptr, (size_t)NP * 4, &d_obs);
  if (line_table) st.out((double*)nullptr, (size_t)NL * 9, &d_geom);
  st.out((double*)nullptr, (size_t)W * D * D, &a.out.H_pp);
  st.out((double*)nullptr, (size_t)W * F * D, &a.out.H_lp);
  st.out((double*)nullptr, (size_t)W * F, &a.out.H_ll);
  st.out((double*)nullptr, (size_t)W * D, &a.out.b_p);
  st.out((double*)nullptr, (size_t)W * F, &a.out.b_l);
  st.out((double*)nullptr, (size_t)W * D * D, &a.out.S);
  st.out((double*)nullptr, (size_t)W * D, &a.out.g);
  st.out(rout ? rout->Sx : nullptr, (size_t)W * Dx * Dx, &Sx);
  st.out(rout ? rout->gx : nullptr, (size_t)W * Dx, &gx);
  if (step) {
    st.out(gout->dx, (size_t)W * Dx, &dxv);
    st.out(gout->cost, (size_t)W * 3, &cost);
    st.out(gout->solved, (size_t)W, &solved);
    st.out(gout->poses, (size_t)W * P * 7, &o_poses);
    st.out(gout->ex_pose, (size_t)W * 7, &o_ex);
    st.out(gout->inv_depth, (size_t)W * F, &o_dep);
    if (X > 0) st.out(gout->extra, (size_t)W * X, &o_extra);
  }
  rc = st.commit();
  if (rc != VIML_OK) return rc;
  a.cache = cache;
  if (obs_table) {
    a.pf_obs = d_obs;
    rc = obs_f32 ? viml_launch_expand_obs(ctx, a, d_fobs32, d_obsj32, true) : viml_launch_expand_obs(ctx, a, d_fobs, d_obsj, false);
    if (rc != VIML_OK) return rc;
  }
  if (line_table) {
    a.lf_geom = d_geom;
    rc = viml_launch_expand_lines(ctx, a, d_lidx, d_lseg);
    if (rc != VIML_OK) return rc;
  }
  rc = viml_launch_linearize(ctx, a);
  if (rc != VIML_OK) return rc;
  // a step that does not have to hand Sx / gx to the caller builds and solves the reduced system in one kernel
  bool fused = false;
  if (step && !(rout && (rout->Sx || rout->gx))) {
    rc = viml_launch_cost(ctx, a, dn, nullptr, 0, cost);
    if (rc != VIML_OK) return rc;
    rc = viml_launch_gn_reduced_solve(ctx, W, D, dn, a.out.S, a.out.g, opt ? opt->lambda : 0.0, dxv, solved, cost);
    if (rc == VIML_OK) fused = true;
    else if (rc != VIML_ERR_UNSUPPORTED) return rc;
  }
  if (!fused) {
    rc = viml_launch_reduced(ctx, W, D, dn, a.out.S, a.out.g, Sx, gx);
    if (rc != VIML_OK) return rc;
  }
  if (step) {
    rc = VIML_OK;
    if (!fused) {
      rc = viml_launch_cost(ctx, a, dn, nullptr, 0, cost);
      if (rc == VIML_OK) rc = viml_launch_gn_solve(ctx, W, Dx, opt ? opt->lambda : 0.0, Sx, gx, dxv, solved, cost);
    }
    if (rc == VIML_OK) rc = viml_launch_gn_update(ctx, a, X, d_extra, dxv, solved, o_poses, o_ex, o_dep, o_extra);
    if (rc != VIML_OK) return rc;
    LinearizeArgs a2 = a;
    a2.poses = o_poses, a2.ex_pose = o_ex, a2.inv_depth = o_dep;
    rc = viml_launch_prep(ctx, a2);
    if (rc == VIML_OK) rc = viml_launch_cost(ctx, a2, dn, dxv, 1, cost);
    if (rc != VIML_OK) return rc;
  }
  return st.finish();
}

}  // namespace

extern "C" {

int viml_reduced_system(viml_ctx* ctx, const viml_window_batch* in, const viml_dense_factors* dense, const viml_reduced_out* out,
                        uint32_t flags) {
  if (!ctx) return VIML_ERR_INVALID;
  return run(ctx, in, dense, nullptr, nullptr, out, nullptr, flags);
}

int viml_reduced_from_schur(viml_ctx* ctx, int32_t W, int32_t D, const double* S, const double* g, const viml_dense_factors* dense,
                            const viml_reduced_out* out, uint32_t flags) {
  if (!ctx) return VIML_ERR_INVALID;
  const int X = dense ? dense->extra_dim : 0, Dx = D + X;
  const int64_t ND = dense ? dense->n_factors : 0;
  if (W < 0 || D < 1 || X < 0 || ND < 0 || !S || !g || !out || !out->Sx || !out->gx)
    return fail(ctx, VIML_ERR_INVALID, "viml_reduced_from_schur: bad arguments");
  if (ND > 0 && (!dense->window_offset || !dense->row_offset || !dense->col_offset || !dense->jac_offset || !dense->col_index ||
                 !dense->residual || !dense->jacobian))
    return fail(ctx, VIML_ERR_INVALID, "dense factors: null array");
  if (W == 0) return VIML_OK;
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool dev = (flags & VIML_PTRS_DEVICE) != 0;
  int64_t n_rows = 0, n_cols = 0, n_jac = 0;
  if (ND > 0) {
    if (dev) {
      int64_t t[3];
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(&t[0], dense->row_offset + ND, 8, cudaMemcpyDeviceToHost, ctx->stream));
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(&t[1], dense->col_offset + ND, 8, cudaMemcpyDeviceToHost, ctx->stream));
      VIML_TRY_CUDA(ctx, cudaMemcpyAsync(&t[2], dense->jac_offset + ND, 8, cudaMemcpyDeviceToHost, ctx->stream));
      VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      n_rows = t[0], n_cols = t[1], n_jac = t[2];
    } else {
      n_rows = dense->row_offset[ND], n_cols = dense->col_offset[ND], n_jac = dense->jac_offset[ND];
      for (int64_t k = 0; k < n_cols; ++k)
        if (dense->col_index[k] < 0 || dense->col_index[k] >= Dx) return fail(ctx, VIML_ERR_INVALID, "dense factors: col_index out of range");
    }
  }
  Stager st{ctx, dev};
  const double *dS = nullptr, *dg = nullptr;
  st.in(S, (size_t)W * D * D, &dS);
  st.in(g, (size_t)W * D, &dg);
  DenseArgs dn{};
  dn.X = X, dn.ND = ND;
  if (ND > 0) {
    st.in(dense->window_offset, (size_t)W + 1, &dn.window_offset);
    st.in(dense->row_offset, (size_t)ND + 1, &dn.row_offset);
    st.in(dense->col_offset, (size_t)ND + 1, &dn.col_offset);
    st.in(dense->jac_offset, (size_t)ND + 1, &dn.jac_offset);
    st.in(dense->col_index, (size_t)n_cols, &dn.col_index);
    st.in(dense->residual, (size_t)n_rows, &dn.residual);
    st.in(dense->jacobian, (size_t)n_jac, &dn.jacobian);
  }
  double *Sx = nullptr, *gx = nullptr;
  st.out(out->Sx, (size_t)W * Dx * Dx, &Sx);
  st.out(out->gx, (size_t)W * Dx, &gx);
  int rc = st.commit();
  if (rc != VIML_OK) return rc;
  rc = viml_launch_reduced(ctx, W, D, dn, dS, dg, Sx, gx);
  if (rc != VIML_OK) return rc;
  return st.finish();
}

int viml_gn_step(viml_ctx* ctx, const viml_window_batch* in, const viml_dense_factors* dense, const double* extra_state,
                 const viml_gn_options* opt, const viml_gn_out* out, uint32_t flags) {
  if (!ctx) return VIML_ERR_INVALID;
  if (!out) return fail(ctx, VIML_ERR_INVALID, "null output struct");
  return run(ctx, in, dense, extra_state, opt, nullptr, out, flags);
}

int viml_triangulate_batch(viml_ctx* ctx, const viml_triangulate_in* in, double init_depth, double* depth, uint32_t flags) {
  if (!ctx) return VIML_ERR_INVALID;
  if (!in || !depth) return fail(ctx, VIML_ERR_INVALID, "viml_triangulate_batch: null arguments");
  const int W = in->n_windows, P = in->poses_per_window;
  const int64_t NF = in->n_features;
  if (W < 0 || P < 1 || NF < 0) return fail(ctx, VIML_ERR_INVALID, "viml_triangulate_batch: sizes out of range");
  if (NF == 0) return VIML_OK;
  if (!in->poses || !in->ex_pose || !in->feat_window || !in->start_frame || !in->obs_offset || !in->points)
    return fail(ctx, VIML_ERR_INVALID, "viml_triangulate_batch: null input array");
  VIML_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool dev = (flags & VIML_PTRS_DEVICE) != 0;
  int64_t nobs = 0;
  if (dev) {
    VIML_TRY_CUDA(ctx, cudaMemcpyAsync(&nobs, in->obs_offset + NF, 8, cudaMemcpyDeviceToHost, ctx->stream));
    VIML_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  } else {
    nobs = in->obs_offset[NF];
    for (int64_t l = 0; l < NF; ++l)
      if (in->feat_window[l] < 0 || in->feat_window[l] >= W) return fail(ctx, VIML_ERR_INVALID, "viml_triangulate_batch: feat_window out of range");
  }
  Stager st{ctx, dev};
  const double *dp = nullptr, *de = nullptr, *dpts = nullptr;
  const int32_t *dw = nullptr, *ds = nullptr;
  const int64_t* doff = nullptr;
  st.in(in->poses, (size_t)W * P * 7, &dp);
  st.in(in->ex_pose, (size_t)W * 7, &de);
  st.in(in->feat_window, (size_t)NF, &dw);
  st.in(in->start_frame, (size_t)NF, &ds);
  st.in(in->obs_offset, (size_t)NF + 1, &doff);
  st.in(in->points, (size_t)nobs * 3, &dpts);
  double* dd = nullptr;
  st.out(depth, (size_t)NF, &dd);
  int rc = st.commit();
  if (rc != VIML_OK) return rc;
  rc = viml_launch_triangulate(ctx, P, NF, dp, de, dw, ds, doff, dpts, init_depth, dd);
  if (rc != VIML_OK) return rc;
  return st.finish();
}

int viml_load_line_map(viml_ctx* ctx, const char* path, int64_t* n_lines) {
  if (!ctx) return VIML_ERR_INVALID;
  if (!path) return fail(ctx, VIML_ERR_INVALID, "null map path");
  std::ifstream f(path);
  if (!f.is_open()) return fail(ctx, VIML_ERR_INVALID, "cannot open the line map file");
  std::vector<double> rows;
  std::string fline;
  while (std::getline(f, fline)) {   // parameters.cpp:53-59: one Vector6d per text line
    std::istringstream iss(fline);
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6; ++k)
      if (!(iss >> v[k])) {   // fields after a failed extraction: 0 here, uninitialised in the reference
        for (int q = k; q < 6; ++q) v[q] = 0.0;
        break;
      }
    rows.insert(rows.end(), v, v + 6);
  }
  if (n_lines) *n_lines = (int64_t)(rows.size() / 6);
  return viml_set_map(ctx, rows.data(), (int64_t)(rows.size() / 6));
}

}  // extern "C"
