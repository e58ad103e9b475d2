// extern "C" surface of libnif_b200.so (see include/nif_b200.h).  Pure argument checking and
// dispatch; nothing here allocates, synchronises or touches the host copy of any tensor.
#include "nif_common.cuh"

bool nif_plan_uses_bf(const Plan& pl);
int nif_pack_impl(const Plan& pl, long long G, const float* w_h, const float* b_h, float* packed, cudaStream_t st);
int nif_forward_impl(const Plan& pl, long long G, long long B, const float* z, const float* x, int x_shared,
                     const float* packed, float* u, float* save, cudaStream_t st);
int nif_tangent_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed, int n_dir,
                     const float* zdot, const float* xdot, float* u, float* udot, float* save, cudaStream_t st);
int nif_sobolev_backward_impl(const Plan& pl, long long B, const float* z, const float* x, int n_dir, const float* zdot,
                              unsigned zdot_dirs, const float* xdot, const float* packed, const float* save, const float* du,
                              const float* dud, float* dw_h, float* db_h, float beta, float* dz, float* dzdot, float* ws,
                              cudaStream_t st);
int nif_given_w_impl(const Plan& pl, long long B, const float* x, const float* w, float* u, cudaStream_t st);
int nif_backward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                      const float* save, const float* du, float* dw_h, float* db_h, float beta, float* dz,
                      float* ws, cudaStream_t st, cudaEvent_t dz_ev);
int nif_mse_backward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                          const float* u, const float* save, const float* target, const float* sw, float inv_gb,
                          float* loss, float* dw_h, float* db_h, float beta, float* dz, float* ws, cudaStream_t st,
                          cudaEvent_t dz_ev);
int nif_adam_impl(long long n, float* p, const float* g, float* m, float* v, double lr, double b1, double b2,
                  double eps, long long t, float l1, float l2, float gs, cudaStream_t st);
GradWs nif_grad_ws_layout(const Plan& pl, long long B);

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#define NIF_REQUIRE_PTR(p)                                               \
  do {                                                                   \
    if (!(p) || !aligned16(p)) {                                         \
      nif_set_error("%s: argument `%s` is null or not 16-byte aligned", __func__, #p); \
      return NIF_E_BAD_ARG;                                              \
    }                                                                    \
  } while (0)
#define NIF_OPTIONAL_PTR(p)                                              \
  do {                                                                   \
    if ((p) && !aligned16(p)) {                                          \
      nif_set_error("%s: argument `%s` is not 16-byte aligned", __func__, #p); \
      return NIF_E_BAD_ARG;                                              \
    }                                                                    \
  } while (0)

extern "C" int nif_query_sizes(const nif_desc_t* d, int64_t B, nif_sizes_t* out) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (!out || B < 0) { nif_set_error("nif_query_sizes: bad argument"); return NIF_E_BAD_ARG; }
  out->po_dim = pl.P;
  out->n_layers = pl.Lm;
  out->np = pl.NP;
  out->packed_floats = pl.packed_floats;
  out->save_floats_per_row = 2LL * (pl.H + 1) * pl.NP;
  out->grad_ws_floats = nif_grad_ws_layout(pl, B).total;
  out->tile_rows = pl.NP == 128 ? 64 : 128;
  out->kernel_path = nif_plan_uses_tc(pl) ? 2 : (nif_plan_uses_bf(pl) ? 1 : 0);
  return NIF_OK;
}

extern "C" int nif_pack(const nif_desc_t* d, int64_t G, const float* w_h, const float* b_h, float* packed,
                        void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (G < 1) { nif_set_error("nif_pack: G=%lld", (long long)G); return NIF_E_BAD_ARG; }
  if (pl.K > 0) NIF_REQUIRE_PTR(w_h);
  if (pl.K > 0 && G != 1) { nif_set_error("nif_pack: G>1 requires K==0 (explicit weight vectors in b_h)"); return NIF_E_BAD_ARG; }
  NIF_REQUIRE_PTR(b_h);
  NIF_REQUIRE_PTR(packed);
  if (pl.tc && G != 1) { nif_set_error("nif_pack: the tensor-core image (dtype_compute=2) is built for G==1"); return NIF_E_BAD_ARG; }
  return nif_pack_impl(pl, G, w_h, b_h, packed, static_cast<cudaStream_t>(stream));
}

extern "C" int nif_forward(const nif_desc_t* d, int64_t G, int64_t B, const float* z, const float* x,
                           int32_t x_shared, const float* packed, float* u, float* save, void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (G < 1 || B < 0) { nif_set_error("nif_forward: G=%lld B=%lld", (long long)G, (long long)B); return NIF_E_BAD_ARG; }
  if (B == 0) return NIF_OK;
  if (pl.K > 0) NIF_REQUIRE_PTR(z);
  NIF_REQUIRE_PTR(x);
  NIF_REQUIRE_PTR(packed);
  NIF_REQUIRE_PTR(u);
  NIF_OPTIONAL_PTR(save);
  if (save && G != 1) { nif_set_error("nif_forward: the activation stash needs G==1"); return NIF_E_BAD_ARG; }
  return nif_forward_impl(pl, G, B, z, x, x_shared, packed, u, save, static_cast<cudaStream_t>(stream));
}

extern "C" int nif_forward_tangent(const nif_desc_t* d, int64_t B, const float* z, const float* x,
                                   const float* packed, int32_t n_dir, const float* zdot, const float* xdot,
                                   float* u, float* udot, void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (B < 0 || n_dir < 1 || n_dir > NIF_MAX_DIR) {
    nif_set_error("nif_forward_tangent: B=%lld n_dir=%d (max %d)", (long long)B, n_dir, NIF_MAX_DIR);
    return NIF_E_BAD_ARG;
  }
  if (B == 0) return NIF_OK;
  if (pl.K > 0) NIF_REQUIRE_PTR(z);
  NIF_REQUIRE_PTR(x);
  NIF_REQUIRE_PTR(packed);
  NIF_REQUIRE_PTR(u);
  NIF_REQUIRE_PTR(udot);
  NIF_OPTIONAL_PTR(zdot);
  NIF_OPTIONAL_PTR(xdot);
  return nif_tangent_impl(pl, B, z, x, packed, n_dir, zdot, xdot, u, udot, nullptr, static_cast<cudaStream_t>(stream));
}

int nif_tangent2_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                      const float* zdot, const float* xdot, const float* zddot, float* u, float* udot, float* uddot,
                      cudaStream_t st);
extern "C" int nif_forward_tangent2(const nif_desc_t* d, int64_t B, const float* z, const float* x,
                                    const float* packed, const float* zdot, const float* xdot, const float* zddot,
                                    float* u, float* udot, float* uddot, void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (B < 0) { nif_set_error("nif_forward_tangent2: B=%lld", (long long)B); return NIF_E_BAD_ARG; }
  if (B == 0) return NIF_OK;
  if (pl.K > 0) NIF_REQUIRE_PTR(z);
  NIF_REQUIRE_PTR(x); NIF_REQUIRE_PTR(packed); NIF_REQUIRE_PTR(u); NIF_REQUIRE_PTR(udot); NIF_REQUIRE_PTR(uddot);
  NIF_OPTIONAL_PTR(zdot); NIF_OPTIONAL_PTR(xdot); NIF_OPTIONAL_PTR(zddot);
  return nif_tangent2_impl(pl, B, z, x, packed, zdot, xdot, zddot, u, udot, uddot, static_cast<cudaStream_t>(stream));
}

extern "C" int nif_sobolev_query_dirs(const nif_desc_t* d, int64_t B, int32_t n_dir, int64_t* save_floats_per_row,
                                      int64_t* ws_floats) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (B < 0 || n_dir < 1 || n_dir > NIF_MAX_DIR) {
    nif_set_error("nif_sobolev_query: B=%lld n_dir=%d (max %d)", (long long)B, n_dir, NIF_MAX_DIR);
    return NIF_E_BAD_ARG;
  }
  if (save_floats_per_row) *save_floats_per_row = (2LL + 2 * n_dir) * (pl.H + 1) * pl.NP;
  if (ws_floats) *ws_floats = nif_grad_ws_layout(pl, B).total + 2LL * (pl.H + 1) * nif_tiled_rows(B) * pl.NP;
  return NIF_OK;
}

extern "C" int nif_sobolev_query(const nif_desc_t* d, int64_t B, int64_t* save_floats_per_row, int64_t* ws_floats) {
  return nif_sobolev_query_dirs(d, B, 1, save_floats_per_row, ws_floats);
}

extern "C" int nif_forward_tangent_save(const nif_desc_t* d, int64_t B, const float* z, const float* x,
                                        const float* packed, int32_t n_dir, const float* zdot, const float* xdot,
                                        float* u, float* udot, float* save, void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (B < 0 || n_dir < 1 || n_dir > NIF_MAX_DIR) {
    nif_set_error("nif_forward_tangent_save: B=%lld n_dir=%d (max %d)", (long long)B, n_dir, NIF_MAX_DIR);
    return NIF_E_BAD_ARG;
  }
  if (B == 0) return NIF_OK;
  if (pl.K > 0) NIF_REQUIRE_PTR(z);
  NIF_REQUIRE_PTR(x); NIF_REQUIRE_PTR(packed); NIF_REQUIRE_PTR(u); NIF_REQUIRE_PTR(udot); NIF_REQUIRE_PTR(xdot);
  NIF_REQUIRE_PTR(save); NIF_OPTIONAL_PTR(zdot);
  return nif_tangent_impl(pl, B, z, x, packed, n_dir, zdot, xdot, u, udot, save, static_cast<cudaStream_t>(stream));
}

extern "C" int nif_sobolev_backward_dirs(const nif_desc_t* d, int64_t B, const float* z, const float* x, int32_t n_dir,
                                         const float* zdot, uint32_t zdot_dirs, const float* xdot, const float* packed, const float* save,
                                         const float* du, const float* dudot, float* dw_h, float* db_h, float beta,
                                         float* dz, float* dzdot, float* ws, void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (B < 0 || n_dir < 1 || n_dir > NIF_MAX_DIR) {
    nif_set_error("nif_sobolev_backward: B=%lld n_dir=%d (max %d)", (long long)B, n_dir, NIF_MAX_DIR);
    return NIF_E_BAD_ARG;
  }
  if (pl.wide_last) {
    nif_set_error("nif_sobolev_backward: trunk plans have no reverse-over-forward pass");
    return NIF_E_UNSUPPORTED;
  }
  if (B == 0) return NIF_OK;
  if (pl.K > 0) { NIF_REQUIRE_PTR(z); NIF_REQUIRE_PTR(dw_h); NIF_REQUIRE_PTR(dz); }
  NIF_REQUIRE_PTR(x); NIF_REQUIRE_PTR(xdot); NIF_REQUIRE_PTR(packed); NIF_REQUIRE_PTR(save); NIF_REQUIRE_PTR(du);
  NIF_REQUIRE_PTR(dudot); NIF_REQUIRE_PTR(db_h); NIF_REQUIRE_PTR(ws);
  NIF_OPTIONAL_PTR(zdot);
  if (pl.K == 0) zdot = nullptr;
  if (zdot) NIF_REQUIRE_PTR(dzdot);
  return nif_sobolev_backward_impl(pl, B, z, x, n_dir, zdot, zdot_dirs, xdot, packed, save, du, dudot, dw_h, db_h, beta, dz,
                                   dzdot, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int nif_sobolev_backward(const nif_desc_t* d, int64_t B, const float* z, const float* x, const float* xdot,
                                    const float* packed, const float* save, const float* du, const float* dudot,
                                    float* dw_h, float* db_h, float beta, float* dz, float* ws, void* stream) {
  return nif_sobolev_backward_dirs(d, B, z, x, 1, nullptr, 0u, xdot, packed, save, du, dudot, dw_h, db_h, beta, dz, nullptr, ws,
                                   stream);
}

extern "C" int nif_forward_given_w(const nif_desc_t* d, int64_t B, const float* x, const float* w, float* u,
                                   void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (B < 0) { nif_set_error("nif_forward_given_w: B=%lld", (long long)B); return NIF_E_BAD_ARG; }
  if (B == 0) return NIF_OK;
  NIF_REQUIRE_PTR(x);
  NIF_REQUIRE_PTR(w);
  NIF_REQUIRE_PTR(u);
  return nif_given_w_impl(pl, B, x, w, u, static_cast<cudaStream_t>(stream));
}

extern "C" int nif_mse_backward_ev(const nif_desc_t* d, int64_t B, const float* z, const float* x, const float* packed,
                                   const float* u, const float* save, const float* target, const float* sample_weight,
                                   float inv_global_batch, float* loss, float* dw_h, float* db_h, float beta,
                                   float* dz, float* ws, void* dz_ready_event, void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (B < 0) { nif_set_error("nif_mse_backward: B=%lld", (long long)B); return NIF_E_BAD_ARG; }
  if (B == 0) return NIF_OK;
  if (pl.K > 0) { NIF_REQUIRE_PTR(z); NIF_REQUIRE_PTR(dw_h); NIF_REQUIRE_PTR(dz); }
  NIF_REQUIRE_PTR(x); NIF_REQUIRE_PTR(packed); NIF_REQUIRE_PTR(u); NIF_REQUIRE_PTR(save);
  NIF_REQUIRE_PTR(target); NIF_OPTIONAL_PTR(sample_weight);
  if (!loss) { nif_set_error("nif_mse_backward: loss is null"); return NIF_E_BAD_ARG; }
  NIF_REQUIRE_PTR(db_h); NIF_REQUIRE_PTR(ws);
  return nif_mse_backward_impl(pl, B, z, x, packed, u, save, target, sample_weight, inv_global_batch, loss, dw_h,
                               db_h, beta, dz, ws, static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(dz_ready_event));
}

extern "C" int nif_mse_backward(const nif_desc_t* d, int64_t B, const float* z, const float* x, const float* packed,
                                const float* u, const float* save, const float* target, const float* sample_weight,
                                float inv_global_batch, float* loss, float* dw_h, float* db_h, float beta,
                                float* dz, float* ws, void* stream) {
  return nif_mse_backward_ev(d, B, z, x, packed, u, save, target, sample_weight, inv_global_batch, loss, dw_h, db_h, beta, dz,
                             ws, nullptr, stream);
}

extern "C" int nif_backward(const nif_desc_t* d, int64_t B, const float* z, const float* x, const float* packed,
                            const float* save, const float* du, float* dw_h, float* db_h, float beta, float* dz,
                            float* ws, void* stream) {
  Plan pl;
  int rc = nif_make_plan(d, &pl);
  if (rc) return rc;
  if (B < 0) { nif_set_error("nif_backward: B=%lld", (long long)B); return NIF_E_BAD_ARG; }
  if (B == 0) return NIF_OK;
  if (pl.K > 0) { NIF_REQUIRE_PTR(z); NIF_REQUIRE_PTR(dw_h); NIF_REQUIRE_PTR(dz); }
  NIF_REQUIRE_PTR(x); NIF_REQUIRE_PTR(packed); NIF_REQUIRE_PTR(save); NIF_REQUIRE_PTR(du);
  NIF_REQUIRE_PTR(db_h); NIF_REQUIRE_PTR(ws);
  return nif_backward_impl(pl, B, z, x, packed, save, du, dw_h, db_h, beta, dz, ws,
                           static_cast<cudaStream_t>(stream), nullptr);
}

extern "C" int nif_adam_step(int64_t n, float* p, const float* g, float* m, float* v, double lr, double b1,
                             double b2, double eps, int64_t t, float l1, float l2, float g_scale, void* stream) {
  if (n < 0 || t < 1) { nif_set_error("nif_adam_step: n=%lld t=%lld", (long long)n, (long long)t); return NIF_E_BAD_ARG; }
  if (n == 0) return NIF_OK;
  NIF_REQUIRE_PTR(p); NIF_REQUIRE_PTR(g); NIF_REQUIRE_PTR(m); NIF_REQUIRE_PTR(v);
  return nif_adam_impl(n, p, g, m, v, lr, b1, b2, eps, t, l1, l2, g_scale, static_cast<cudaStream_t>(stream));
}

int nif_adam_dev_impl(long long n, float* p, const float* g, float* m, float* v, const float* alpha_dev, double b1,
                      double b2, double eps, float l1, float l2, float gs, cudaStream_t st);
extern "C" int nif_adam_step_dev(int64_t n, float* p, const float* g, float* m, float* v, const float* alpha_dev,
                                 double b1, double b2, double eps, float l1, float l2, float g_scale, void* stream) {
  if (n < 0) { nif_set_error("nif_adam_step_dev: n=%lld", (long long)n); return NIF_E_BAD_ARG; }
  if (n == 0) return NIF_OK;
  NIF_REQUIRE_PTR(p); NIF_REQUIRE_PTR(g); NIF_REQUIRE_PTR(m); NIF_REQUIRE_PTR(v);
  if (!alpha_dev) { nif_set_error("nif_adam_step_dev: alpha_dev is null"); return NIF_E_BAD_ARG; }
  return nif_adam_dev_impl(n, p, g, m, v, alpha_dev, b1, b2, eps, l1, l2, g_scale, static_cast<cudaStream_t>(stream));
}

// ---- ParameterNet trunk --------------------------------------------------------------------------------
int nif_make_trunk_plan(int pi, int K, int n_st, int l_st, int act, Plan* out);
int nif_trunk_forward_impl(const Plan& pl, long long B, const float* p_in, const float* theta, float* z, float* save,
                           float* packed, cudaStream_t st);
int nif_trunk_backward_impl(const Plan& pl, long long B, const float* p_in, const float* theta, const float* save,
                            const float* dz, float* g_theta, float beta, const float* packed, float* ws,
                            cudaStream_t st);

static int trunk_plan(const nif_trunk_desc_t* d, Plan* pl) {
  if (!d) { nif_set_error("null trunk descriptor"); return NIF_E_BAD_DESC; }
  return nif_make_trunk_plan(d->pi, d->latent, d->units, d->nlayers, d->act, pl);
}

bool nif_trunk_uses_tc(int pi, int K, int n, int l, int act);
long long nif_trunk_tc_packed_floats(int pi, int K, int n, int l, int act);
long long nif_trunk_tc_ws_floats(int pi, int K, int n, int l, int act, long long B);
int nif_trunk_tc_forward_impl(int pi, int K, int n, int l, int act, long long B, const float* p_in, const float* theta,
                              float* z, float* save, float* packed, cudaStream_t st);
int nif_trunk_tc_backward_impl(int pi, int K, int n, int l, int act, long long B, const float* p_in, const float* save,
                               const float* dz, float* g_theta, float beta, const float* packed, float* ws,
                               cudaStream_t st);
static bool trunk_tc(const nif_trunk_desc_t* d) { return nif_trunk_uses_tc(d->pi, d->latent, d->units, d->nlayers, d->act); }

extern "C" int nif_trunk_query(const nif_trunk_desc_t* d, int64_t B, int64_t* n_theta, int64_t* save_floats_per_row,
                               int64_t* packed_floats, int64_t* ws_floats) {
  Plan pl;
  int rc = trunk_plan(d, &pl);
  if (rc) return rc;
  if (B < 0) { nif_set_error("nif_trunk_query: B=%lld", (long long)B); return NIF_E_BAD_ARG; }
  if (n_theta) *n_theta = pl.P;
  if (trunk_tc(d)) {  // tensor-core trunk kernels: 64-wide tiled stash (rows rounded up to 64 by the caller)
    if (save_floats_per_row) *save_floats_per_row = 2LL * (pl.H + 1) * 64;
    if (packed_floats) *packed_floats = nif_trunk_tc_packed_floats(d->pi, d->latent, d->units, d->nlayers, d->act);
    if (ws_floats) *ws_floats = nif_trunk_tc_ws_floats(d->pi, d->latent, d->units, d->nlayers, d->act, B);
    return NIF_OK;
  }
  if (save_floats_per_row) *save_floats_per_row = 2LL * (pl.H + 1) * pl.NP;
  if (packed_floats) *packed_floats = pl.packed_floats;
  if (ws_floats) *ws_floats = nif_grad_ws_layout(pl, B).total;
  return NIF_OK;
}

extern "C" int nif_trunk_kernel_path(const nif_trunk_desc_t* d) {
  Plan pl;
  int rc = trunk_plan(d, &pl);
  if (rc) return rc;
  return trunk_tc(d) ? 3 : 0;
}

extern "C" int nif_trunk_forward(const nif_trunk_desc_t* d, int64_t B, const float* p_in, const float* theta, float* z,
                                 float* save, float* packed, void* stream) {
  Plan pl;
  int rc = trunk_plan(d, &pl);
  if (rc) return rc;
  if (B < 0) { nif_set_error("nif_trunk_forward: B=%lld", (long long)B); return NIF_E_BAD_ARG; }
  if (B == 0) return NIF_OK;
  NIF_REQUIRE_PTR(p_in); NIF_REQUIRE_PTR(theta); NIF_REQUIRE_PTR(z); NIF_OPTIONAL_PTR(save); NIF_REQUIRE_PTR(packed);
  if (trunk_tc(d))
    return nif_trunk_tc_forward_impl(d->pi, d->latent, d->units, d->nlayers, d->act, B, p_in, theta, z, save, packed,
                                     static_cast<cudaStream_t>(stream));
  return nif_trunk_forward_impl(pl, B, p_in, theta, z, save, packed, static_cast<cudaStream_t>(stream));
}

extern "C" int nif_trunk_backward(const nif_trunk_desc_t* d, int64_t B, const float* p_in, const float* theta,
                                  const float* save, const float* dz, float* g_theta, float beta, const float* packed,
                                  float* ws, void* stream) {
  Plan pl;
  int rc = trunk_plan(d, &pl);
  if (rc) return rc;
  if (B < 0) { nif_set_error("nif_trunk_backward: B=%lld", (long long)B); return NIF_E_BAD_ARG; }
  if (B == 0) return NIF_OK;
  NIF_REQUIRE_PTR(p_in); NIF_REQUIRE_PTR(theta); NIF_REQUIRE_PTR(save); NIF_REQUIRE_PTR(dz);
  NIF_REQUIRE_PTR(g_theta); NIF_REQUIRE_PTR(packed); NIF_REQUIRE_PTR(ws);
  if (trunk_tc(d))
    return nif_trunk_tc_backward_impl(d->pi, d->latent, d->units, d->nlayers, d->act, B, p_in, save, dz, g_theta, beta,
                                      packed, ws, static_cast<cudaStream_t>(stream));
  return nif_trunk_backward_impl(pl, B, p_in, theta, save, dz, g_theta, beta, packed, ws,
                                 static_cast<cudaStream_t>(stream));
}

// ---- per-kernel timing (bench.py's kernel table) -------------------------------------------------------------------
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
int g_nif_prof_on = 0;
namespace {
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof;
}
void nif_prof_push(const char* name, cudaStream_t st, bool begin) {
  if (begin) {
    ProfRec r;
    r.name = name;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    g_prof.push_back(r);
  } else if (!g_prof.empty()) {
    cudaEventRecord(g_prof.back().e1, st);
  }
}
extern "C" int nif_profile_begin(void) {
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_nif_prof_on = 1;
  return NIF_OK;
}
// Closes the profile (blocks until the recorded work is done) and writes one line per kernel name,
// "name launches total_ms\n", in order of first launch, into `out` (NUL-terminated, truncated to cap).
extern "C" int nif_profile_end(char* out, int64_t cap) {
  g_nif_prof_on = 0;
  if (!out || cap < 1) { nif_set_error("nif_profile_end: null buffer"); return NIF_E_BAD_ARG; }
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<long long> cnt;
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) t = 0.f;
    size_t i = 0;
    for (; i < names.size(); ++i) if (names[i] == r.name) break;
    if (i == names.size()) { names.push_back(r.name); ms.push_back(0.0); cnt.push_back(0); }
    ms[i] += t; cnt[i] += 1;
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  std::string s;
  char line[256];
  for (size_t i = 0; i < names.size(); ++i) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", names[i].c_str(), cnt[i], ms[i]);
    s += line;
  }
  strncpy(out, s.c_str(), (size_t)cap - 1);
  out[cap - 1] = 0;
  return NIF_OK;
}

// ---- CRC-32C for the TFRecord framing (host code, slicing-by-8) ----------------------------------------------------
static uint32_t g_crc_tab[8][256];
static bool g_crc_ready = false;
static void crc32c_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);  // reflected Castagnoli polynomial
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xFFu];
  g_crc_ready = true;
}
extern "C" uint32_t nif_crc32c(const void* data, uint64_t n, uint32_t crc) {
  if (!g_crc_ready) crc32c_init();  // idempotent: a racing second initialisation writes the same values
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = ~crc;
  while (n >= 8) {
    const uint32_t lo = c ^ ((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24);
    c = g_crc_tab[7][lo & 0xFFu] ^ g_crc_tab[6][(lo >> 8) & 0xFFu] ^ g_crc_tab[5][(lo >> 16) & 0xFFu] ^ g_crc_tab[4][lo >> 24] ^
        g_crc_tab[3][p[4]] ^ g_crc_tab[2][p[5]] ^ g_crc_tab[1][p[6]] ^ g_crc_tab[0][p[7]];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_crc_tab[0][(c ^ *p++) & 0xFFu];
  return ~c;
}
