// ParameterNet trunk (everything before the last linear layer) as fused kernels.
//
// Reference: _call_parameter_net (nif/model.py:326-343) over  Dense(act) -> l_st x MLP_SimpleShortCut
// (nif/layers/mlp.py:148-160:  h + act(h W + b)) -> Dense(latent, linear)   (nif/model.py:176-216, 668-720).
// That is exactly the layer structure of the NIF-variant ShapeNet with a fixed (not per-row) weight vector, so
// the trunk reuses the Plan / column layout of that variant with K = 0:
//   theta_t = [ W_first | W_hidden[0..l_st) | W_bottleneck | b_first | b_hidden[..] | b_bottleneck ]
// (the host keeps the trunk variables in this order, so theta_t / its gradient are views of the flat buffers).
// Forward and the reverse data pass are one thread per row with all weights in shared memory (37 kFLOP/row,
// ~3 % of the step); the parameter gradients are the batch-reduction GEMMs of nif_bwd.cu with zt = [1].
#include "nif_tile.cuh"

struct TrunkArgs {
  long long B;
  const float *p_in, *theta, *dz;
  float *z, *save, *da;
};

// Shared-memory image of theta_t, padded so that every matrix row is 64 floats (16-byte aligned, zero padded):
//   W0 [si][64] | WH [H][64][64] | WL [64][SOP] | BIAS [H+2][64]          SOP = so rounded up to 4
struct TrunkSmem {
  int w0, wh, wl, bias, sop, total;
};
__host__ __device__ inline TrunkSmem trunk_smem_layout(const Plan& pl) {
  TrunkSmem t;
  t.sop = (pl.so + 3) / 4 * 4;
  t.w0 = 0;
  t.wh = t.w0 + pl.si * 64;
  t.wl = t.wh + pl.H * 64 * 64;
  t.bias = t.wl + 64 * t.sop;
  t.total = t.bias + (pl.H + 2) * 64;
  return t;
}
__device__ __forceinline__ void trunk_stage(float* ws, const TrunkSmem& t, const Plan& pl, const float* __restrict__ th) {
  const int n = pl.n, so = pl.so, H = pl.H;
  for (int e = threadIdx.x; e < t.total; e += blockDim.x) {
    float v = 0.f;
    if (e < t.wh) {
      const int i = e / 64, j = e % 64;
      if (j < n) v = __ldg(&th[plan_w_off(pl, 0) + i * n + j]);
    } else if (e < t.wl) {
      const int r = e - t.wh, m = r / 4096, i = (r % 4096) / 64, j = r % 64;
      if (i < n && j < n) v = __ldg(&th[plan_w_off(pl, m + 1) + i * n + j]);
    } else if (e < t.bias) {
      const int r = e - t.wl, i = r / t.sop, c = r % t.sop;
      if (i < n && c < so) v = __ldg(&th[plan_w_off(pl, H + 1) + i * so + c]);
    } else {
      const int r = e - t.bias, m = r / 64, j = r % 64;
      if (j < (m == H + 1 ? so : n)) v = __ldg(&th[plan_b_off(pl, m) + j]);
    }
    ws[e] = v;
  }
  __syncthreads();
}

// Both kernels keep the vector that is indexed by the (run-time) reduction index in shared memory
// (vec[i][thread], conflict free) and the other one in registers, so the reduction loop can stay rolled
// (a fully unrolled 64 x 64 body does not fit the instruction cache).
__global__ void __launch_bounds__(128) nif_trunk_fwd_kernel(const Plan pl, const TrunkArgs a, int save_on) {
  extern __shared__ __align__(16) float ws[];
  const TrunkSmem t = trunk_smem_layout(pl);
  trunk_stage(ws, t, pl, a.theta);
  float* hs = ws + t.total + threadIdx.x;  // hs[i * 128]: this row's layer input
  const int n = pl.n, si = pl.si, so = pl.so, H = pl.H, NP = pl.NP;
  // persistent over 128-row tiles: the weights are staged once per CTA
  for (long long b = blockIdx.x * 128LL + threadIdx.x; b < a.B; b += 128LL * gridDim.x) {
    float pre[64];
    {  // first layer
      const float* bb = ws + t.bias;
#pragma unroll
      for (int j = 0; j < 64; ++j) pre[j] = bb[j];
      for (int i = 0; i < si; ++i) {
        const float xi = __ldg(&a.p_in[b * si + i]);
        const float* W = ws + t.w0 + i * 64;
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(W + 4 * j4);
          pre[4 * j4] = fmaf(xi, w.x, pre[4 * j4]); pre[4 * j4 + 1] = fmaf(xi, w.y, pre[4 * j4 + 1]);
          pre[4 * j4 + 2] = fmaf(xi, w.z, pre[4 * j4 + 2]); pre[4 * j4 + 3] = fmaf(xi, w.w, pre[4 * j4 + 3]);
        }
      }
    }
    auto finish = [&](int m, bool residual) {
      float* sh = a.save + (long long)m * a.B * NP + b * NP;
      float* sd = a.save + (long long)(H + 1 + m) * a.B * NP + b * NP;
#pragma unroll
      for (int j4 = 0; j4 < 16; ++j4) {
        float fo[4], dd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * j4 + e;
          float f, d;
          act_fd(pl.act, pre[j], f, d);
          if (residual) f += hs[j * 128];
          if (j >= n) { f = 0.f; d = 0.f; }
          hs[j * 128] = f; fo[e] = f; dd[e] = d;
        }
        if (save_on && 4 * j4 < NP) {
          *reinterpret_cast<float4*>(sh + 4 * j4) = make_float4(fo[0], fo[1], fo[2], fo[3]);
          *reinterpret_cast<float4*>(sd + 4 * j4) = make_float4(dd[0], dd[1], dd[2], dd[3]);
        }
      }
    };
    finish(0, false);
    for (int m = 1; m <= H; ++m) {
      const float* bb = ws + t.bias + m * 64;
#pragma unroll
      for (int j = 0; j < 64; ++j) pre[j] = bb[j];
      const float* W = ws + t.wh + (m - 1) * 4096;
#pragma unroll 4
      for (int i = 0; i < 64; ++i) {
        const float hi = hs[i * 128];
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(W + i * 64 + 4 * j4);
          pre[4 * j4] = fmaf(hi, w.x, pre[4 * j4]); pre[4 * j4 + 1] = fmaf(hi, w.y, pre[4 * j4 + 1]);
          pre[4 * j4 + 2] = fmaf(hi, w.z, pre[4 * j4 + 2]); pre[4 * j4 + 3] = fmaf(hi, w.w, pre[4 * j4 + 3]);
        }
      }
      finish(m, true);
    }
    {  // bottleneck (linear): z[c] = b[c] + sum_i h[i] W[i][c]; pre[] is reused as the accumulator (so <= 64)
      const float* bb = ws + t.bias + (H + 1) * 64;
#pragma unroll
      for (int c = 0; c < 64; ++c) pre[c] = bb[c];
#pragma unroll 4
      for (int i = 0; i < 64; ++i) {
        const float hi = hs[i * 128];
        const float* W = ws + t.wl + i * t.sop;
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          if (4 * c4 < t.sop) {
            const float4 w = *reinterpret_cast<const float4*>(W + 4 * c4);
            pre[4 * c4] = fmaf(hi, w.x, pre[4 * c4]); pre[4 * c4 + 1] = fmaf(hi, w.y, pre[4 * c4 + 1]);
            pre[4 * c4 + 2] = fmaf(hi, w.z, pre[4 * c4 + 2]); pre[4 * c4 + 3] = fmaf(hi, w.w, pre[4 * c4 + 3]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 64; ++c)
        if (c < so) a.z[b * so + c] = pre[c];
    }
  }
}

// reverse data pass: da_m = dh_{m+1} * act'(pre_m),  dh_m = dh_{m+1} + da_m W_m^T  (shortcut), thread = row.
// Also copies the seed dz into stash slot H+1 (zero padded to NP columns): the operand of the last matrix's
// gradient in the batch-reduction GEMM (Plan::wide_last).
__global__ void __launch_bounds__(128) nif_trunk_bwd_kernel(const Plan pl, const TrunkArgs a) {
  extern __shared__ __align__(16) float ws[];
  const TrunkSmem t = trunk_smem_layout(pl);
  trunk_stage(ws, t, pl, a.theta);
  float* dhs = ws + t.total + threadIdx.x;  // dhs[i * 128]: this row's dh
  const int so = pl.so, H = pl.H, NP = pl.NP;
  for (long long b = blockIdx.x * 128LL + threadIdx.x; b < a.B; b += 128LL * gridDim.x) {
    float dp[64];  // first the seed dz (so <= 64 values), then da_m
    {
      float* dzl = a.da + (long long)(H + 1) * a.B * NP + b * NP;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) dp[4 * c4 + e] = (4 * c4 + e < so) ? __ldg(&a.dz[b * so + 4 * c4 + e]) : 0.f;
        if (4 * c4 < NP) *reinterpret_cast<float4*>(dzl + 4 * c4) = make_float4(dp[4 * c4], dp[4 * c4 + 1], dp[4 * c4 + 2], dp[4 * c4 + 3]);
      }
    }
    // dh_{H+1}[i] = sum_c W_L[i][c] dz[c]
#pragma unroll 4
    for (int i = 0; i < 64; ++i) {
      const float* W = ws + t.wl + i * t.sop;
      float s = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        if (4 * c4 < t.sop) {
          const float4 w = *reinterpret_cast<const float4*>(W + 4 * c4);
          s = fmaf(w.x, dp[4 * c4], fmaf(w.y, dp[4 * c4 + 1], fmaf(w.z, dp[4 * c4 + 2], fmaf(w.w, dp[4 * c4 + 3], s))));
        }
      }
      dhs[i * 128] = s;
    }
    for (int m = H; m >= 0; --m) {
      const float* sd = a.save + (long long)(H + 1 + m) * a.B * NP + b * NP;
      float* dag = a.da + (long long)m * a.B * NP + b * NP;
#pragma unroll
      for (int j4 = 0; j4 < 16; ++j4) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * j4 < NP) d = *reinterpret_cast<const float4*>(sd + 4 * j4);
        dp[4 * j4] = dhs[(4 * j4) * 128] * d.x; dp[4 * j4 + 1] = dhs[(4 * j4 + 1) * 128] * d.y;
        dp[4 * j4 + 2] = dhs[(4 * j4 + 2) * 128] * d.z; dp[4 * j4 + 3] = dhs[(4 * j4 + 3) * 128] * d.w;
        if (4 * j4 < NP) *reinterpret_cast<float4*>(dag + 4 * j4) = make_float4(dp[4 * j4], dp[4 * j4 + 1], dp[4 * j4 + 2], dp[4 * j4 + 3]);
      }
      if (m == 0) break;
      const float* W = ws + t.wh + (m - 1) * 4096;
#pragma unroll 4
      for (int i = 0; i < 64; ++i) {
        float s = dhs[i * 128];  // shortcut
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(W + i * 64 + 4 * j4);
          s = fmaf(w.x, dp[4 * j4], fmaf(w.y, dp[4 * j4 + 1], fmaf(w.z, dp[4 * j4 + 2], fmaf(w.w, dp[4 * j4 + 3], s))));
        }
        dhs[i * 128] = s;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
GradWs nif_grad_ws_layout(const Plan& pl, long long B);
int nif_weight_grads_impl(const Plan& pl, long long B, const float* z, const float* x, const float* save, const float* du,
                          float* dw_h, float* db_h, float beta, float* ws, cudaStream_t st);

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
  p.NP = (n_st <= 32 && K <= 32) ? 32 : 64;  // the seed dz is stashed NP wide for the last matrix's gradient
  p.wide_last = 1;
  p.P = p.H * p.n * p.n + (p.si + p.so + 1 + p.H) * p.n + p.so;
  p.tc = 0;
  p.KP = 2; p.NCH = 1; p.KZ = 16; p.LPC = 1; p.NLC = 1;
  if (((size_t)trunk_smem_layout(p).total + 64 * 128) * 4 > 200 * 1024) {
    nif_set_error("trunk weights do not fit in shared memory");
    return NIF_E_UNSUPPORTED;
  }
  *out = p;
  return NIF_OK;
}

static unsigned trunk_grid(long long B) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long tiles = (B + 127) / 128, g = 3LL * sms;
  return (unsigned)(tiles < g ? tiles : g);
}

int nif_trunk_forward_impl(const Plan& pl, long long B, const float* p_in, const float* theta, float* z, float* save,
                           cudaStream_t st) {
  if (B <= 0) return NIF_OK;
  TrunkArgs a = {};
  a.B = B; a.p_in = p_in; a.theta = theta; a.z = z; a.save = save;
  const size_t smem = ((size_t)trunk_smem_layout(pl).total + 64 * 128) * 4;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_trunk_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nif_trunk_fwd_kernel<<<trunk_grid(B), 128, smem, st>>>(pl, a, save ? 1 : 0);
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

int nif_trunk_backward_impl(const Plan& pl, long long B, const float* p_in, const float* theta, const float* save,
                            const float* dz, float* g_theta, float beta, float* ws, cudaStream_t st) {
  if (B <= 0) return NIF_OK;
  const GradWs w = nif_grad_ws_layout(pl, B);
  TrunkArgs a = {};
  a.B = B; a.p_in = p_in; a.theta = theta; a.dz = dz; a.save = const_cast<float*>(save); a.da = ws + w.da;
  const size_t smem = ((size_t)trunk_smem_layout(pl).total + 64 * 128) * 4;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_trunk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nif_trunk_bwd_kernel<<<trunk_grid(B), 128, smem, st>>>(pl, a);
  NIF_CUDA_CHECK(cudaGetLastError());
  // parameter gradients: batch-reduction GEMMs of the reverse pass with zt = [1] (K = 0)
  return nif_weight_grads_impl(pl, B, nullptr, p_in, save, dz, nullptr, g_theta, beta, ws, st);
}
