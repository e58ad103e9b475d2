// ParameterNet trunk (everything before the last linear layer) through the fused tile kernels.
//
// Reference: _call_parameter_net (nif/model.py:326-343) over  Dense(act) -> l_st x MLP_SimpleShortCut
// (nif/layers/mlp.py:148-160:  h + act(h W + b)) -> Dense(latent, linear)   (nif/model.py:176-216, 668-720).
// That is exactly the layer structure of the NIF-variant ShapeNet with one fixed (not per-row) weight vector,
// so the trunk is a Plan of that variant with K = 0 (zt = [1]) and Plan::wide_last (its last matrix is
// units x latent, too wide for the thin last-layer code, so it runs through the tile GEMM):
//   theta_t = [ W_first | W_hidden[0..l_st) | W_bottleneck | b_first | b_hidden[..] | b_bottleneck ]
// (the host keeps the trunk variables in this order, so theta_t and its gradient are views of the flat buffers).
// Forward = nif_pack + nif_fwd_kernel; reverse = nif_bwd_data_kernel + the batch-reduction weight kernels.
#include "nif_common.cuh"

void nif_plan_layout(Plan* p);
int nif_pack_impl(const Plan& pl, long long G, const float* w_h, const float* b_h, float* packed, cudaStream_t st);
int nif_forward_impl(const Plan& pl, long long G, long long B, const float* z, const float* x, int x_shared,
                     const float* packed, float* u, float* save, cudaStream_t st);
int nif_backward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                      const float* save, const float* du, float* dw_h, float* db_h, float beta, float* dz,
                      float* ws, cudaStream_t st, cudaEvent_t dz_ev);

int nif_make_trunk_plan(int pi, int K, int n_st, int l_st, int act, Plan* out) {
  if (pi < 1 || pi > NIF_MAX_SI || K < 1 || K > 64 || n_st < 1 || n_st > 64 || l_st < 0 || l_st > 64 ||
      act < 0 || act > NIF_ACT_SIGMOID || act == NIF_ACT_SINE) {
    nif_set_error("trunk (pi=%d, latent=%d, units=%d, layers=%d, act=%d) is outside the fused trunk kernels", pi, K, n_st,
                  l_st, act);
    return NIF_E_UNSUPPORTED;
  }
  Plan p = {};
  p.variant = NIF_VARIANT_NIF;
  p.act = act;
  p.si = pi; p.so = K; p.n = n_st; p.l = l_st; p.K = 0;
  p.omega0 = 1.0f;
  p.H = l_st;
  p.Lm = p.H + 2;
  p.NP = (n_st <= 32 && K <= 32) ? 32 : 64;  // the output is NP wide in the tile GEMM
  p.P = p.H * p.n * p.n + (p.si + p.so + 1 + p.H) * p.n + p.so;
  p.wide_last = 1;
  p.tc = 0;
  nif_plan_layout(&p);
  *out = p;
  return NIF_OK;
}

int nif_trunk_forward_impl(const Plan& pl, long long B, const float* p_in, const float* theta, float* z, float* save,
                           float* packed, cudaStream_t st) {
  if (B <= 0) return NIF_OK;
  int rc = nif_pack_impl(pl, 1, nullptr, theta, packed, st);
  if (rc) return rc;
  return nif_forward_impl(pl, 1, B, nullptr, p_in, 0, packed, z, save, st);
}

int nif_trunk_backward_impl(const Plan& pl, long long B, const float* p_in, const float* theta, const float* save,
                            const float* dz, float* g_theta, float beta, const float* packed, float* ws,
                            cudaStream_t st) {
  if (B <= 0) return NIF_OK;
  (void)theta;  // `packed` was built from it by the forward call of this step
  return nif_backward_impl(pl, B, nullptr, p_in, packed, save, dz, nullptr, g_theta, beta, nullptr, ws, st, nullptr);
}
