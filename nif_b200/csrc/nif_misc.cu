// Bandwidth-bound side kernels: explicit-weights ShapeNet (model_x_to_u_given_w), Adam, and the
// FP32-FMA peak probe used by the benchmark's roofline.
#include <cmath>
#include "nif_common.cuh"

// ---------------------------------------------------------------------------------------------------
// model_x_to_u_given_w (nif/model.py:435-464, 956-986): every row brings its own weight vector
// w[b, :] (reference column layout, P floats).  One warp per row streams that row's P floats exactly
// once, fully coalesced: the kernel is bound by HBM (4 P bytes per row).
// ---------------------------------------------------------------------------------------------------
template <int JPL>  // output columns per lane = ceil(n / 32)
__global__ void __launch_bounds__(256) nif_given_w_kernel(const Plan pl, long long B, const float* __restrict__ x,
                                                          const float* __restrict__ w, float* __restrict__ u) {
  __shared__ float hs_all[8][128];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* hs = hs_all[wid];
  const int n = pl.n, si = pl.si, so = pl.so, H = pl.H;
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long b = (long long)blockIdx.x * 8 + wid; b < B; b += nwarps) {
    const float* wr = w + b * pl.P;
    float hv[JPL], carry[JPL];
#pragma unroll
    for (int q = 0; q < JPL; ++q) { hv[q] = 0.f; carry[q] = 0.f; }
    // layer 0
    {
      const float om = plan_omega(pl, 0);
      const int bo = plan_b_off(pl, 0);
#pragma unroll
      for (int q = 0; q < JPL; ++q) {
        const int j = lane + 32 * q;
        if (j < n) {
          float s = 0.f;
          for (int i = 0; i < si; ++i) s = fmaf(__ldg(&x[b * si + i]), __ldg(&wr[i * n + j]), s);
          hv[q] = act_f(pl.act, fmaf(om, s, __ldg(&wr[bo + j])));
        }
      }
    }
    for (int m = 1; m <= H; ++m) {
      __syncwarp();
#pragma unroll
      for (int q = 0; q < JPL; ++q) {
        const int j = lane + 32 * q;
        if (j < n) hs[j] = hv[q];
      }
      __syncwarp();
      const float om = plan_omega(pl, m), alpha = plan_alpha(pl, m);
      const int res = plan_res(pl, m);
      const float* wm = wr + plan_w_off(pl, m);
      const int bo = plan_b_off(pl, m);
      float s[JPL];
#pragma unroll
      for (int q = 0; q < JPL; ++q) s[q] = 0.f;
      int i = 0;
      for (; i + 8 <= n; i += 8) {
        float wv[8][JPL];
#pragma unroll
        for (int e = 0; e < 8; ++e)
#pragma unroll
          for (int q = 0; q < JPL; ++q) {
            const int j = lane + 32 * q;
            wv[e][q] = (j < n) ? __ldg(&wm[(i + e) * n + j]) : 0.f;
          }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float hi = hs[i + e];
#pragma unroll
          for (int q = 0; q < JPL; ++q) s[q] = fmaf(hi, wv[e][q], s[q]);
        }
      }
      for (; i < n; ++i) {
        const float hi = hs[i];
#pragma unroll
        for (int q = 0; q < JPL; ++q) {
          const int j = lane + 32 * q;
          if (j < n) s[q] = fmaf(hi, __ldg(&wm[i * n + j]), s[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < JPL; ++q) {
        const int j = lane + 32 * q;
        if (j < n) {
          float o = alpha * act_f(pl.act, fmaf(om, s[q], __ldg(&wr[bo + j])));
          if (res == 1) o += hv[q];
          if (res == 2) carry[q] = hv[q];
          if (res == 3) o += 0.5f * carry[q];
          hv[q] = o;
        }
      }
    }
    // last layer
    {
      const float* wl = wr + plan_w_off(pl, H + 1);
      const int bo = plan_b_off(pl, H + 1);
      for (int c = 0; c < so; ++c) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < JPL; ++q) {
          const int j = lane + 32 * q;
          if (j < n) s = fmaf(hv[q], __ldg(&wl[j * so + c]), s);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) u[b * so + c] = s + __ldg(&wr[bo + c]);
      }
    }
  }
}

int nif_given_w_impl(const Plan& pl, long long B, const float* x, const float* w, float* u, cudaStream_t st) {
  if (B <= 0) return NIF_OK;
  long long nblk = (B + 7) / 8;
  if (nblk > 148 * 8) nblk = 148 * 8;
  const int jpl = (pl.n + 31) / 32;
  switch (jpl) {
    case 1: { NIF_PROF("nif_given_w_kernel", st); nif_given_w_kernel<1><<<(unsigned)nblk, 256, 0, st>>>(pl, B, x, w, u); } break;
    case 2: { NIF_PROF("nif_given_w_kernel", st); nif_given_w_kernel<2><<<(unsigned)nblk, 256, 0, st>>>(pl, B, x, w, u); } break;
    case 3: { NIF_PROF("nif_given_w_kernel", st); nif_given_w_kernel<3><<<(unsigned)nblk, 256, 0, st>>>(pl, B, x, w, u); } break;
    default: { NIF_PROF("nif_given_w_kernel", st); nif_given_w_kernel<4><<<(unsigned)nblk, 256, 0, st>>>(pl, B, x, w, u); } break;
  }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// ---------------------------------------------------------------------------------------------------
// Adam with tf.keras semantics (epsilon outside the bias correction); HBM-bound: 28 B / parameter.
// ---------------------------------------------------------------------------------------------------
// omb1 = 1 - beta_1, omb2 = 1 - beta_2 are formed in double on the host (Keras forms them in Python floats)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float alpha, float omb1, float omb2,
                                         float eps, float l1, float l2, float gs) {
  g *= gs;
  if (l1 != 0.f) g += l1 * (p > 0.f ? 1.f : (p < 0.f ? -1.f : 0.f));
  if (l2 != 0.f) g += 2.f * l2 * p;
  m += (g - m) * omb1;
  v += (g * g - v) * omb2;
  p -= alpha * m / (sqrtf(v) + eps);
}

// alpha_dev != nullptr: the step size comes from device memory (graph-replayed steps, nif_adam_step_dev)
__global__ void __launch_bounds__(256) nif_adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, float alpha,
                                                       const float* __restrict__ alpha_dev,
                                                       float b1, float b2, float eps, float l1, float l2, float gs) {
  if (alpha_dev) alpha = __ldg(alpha_dev);
  const long long n4 = n / 4;
  const long long stride = 256LL * gridDim.x;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, alpha, b1, b2, eps, l1, l2, gs);
    adam_one(pp.y, gg.y, mm.y, vv.y, alpha, b1, b2, eps, l1, l2, gs);
    adam_one(pp.z, gg.z, mm.z, vv.z, alpha, b1, b2, eps, l1, l2, gs);
    adam_one(pp.w, gg.w, mm.w, vv.w, alpha, b1, b2, eps, l1, l2, gs);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = n4 * 4 + blockIdx.x * 256LL + threadIdx.x; i < n; i += stride)
    adam_one(p[i], g[i], m[i], v[i], alpha, b1, b2, eps, l1, l2, gs);
}

int nif_adam_impl(long long n, float* p, const float* g, float* m, float* v, double lr, double b1, double b2,
                  double eps, long long t, float l1, float l2, float gs, cudaStream_t st) {
  if (n <= 0) return NIF_OK;
  const double alpha = lr * std::sqrt(1.0 - std::pow(b2, (double)t)) / (1.0 - std::pow(b1, (double)t));
  long long nblk = (n / 4 + 255) / 256;
  if (nblk < 1) nblk = 1;
  if (nblk > 148 * 8) nblk = 148 * 8;
  { NIF_PROF("nif_adam_kernel", st); nif_adam_kernel<<<(unsigned)nblk, 256, 0, st>>>(n, p, g, m, v, (float)alpha, nullptr, (float)(1.0 - b1),
                                                  (float)(1.0 - b2), (float)eps, l1, l2, gs); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

int nif_adam_dev_impl(long long n, float* p, const float* g, float* m, float* v, const float* alpha_dev, double b1,
                      double b2, double eps, float l1, float l2, float gs, cudaStream_t st) {
  if (n <= 0) return NIF_OK;
  long long nblk = (n / 4 + 255) / 256;
  if (nblk < 1) nblk = 1;
  if (nblk > 148 * 8) nblk = 148 * 8;
  { NIF_PROF("nif_adam_kernel", st); nif_adam_kernel<<<(unsigned)nblk, 256, 0, st>>>(n, p, g, m, v, 0.f, alpha_dev, (float)(1.0 - b1), (float)(1.0 - b2),
                                                  (float)eps, l1, l2, gs); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// ---------------------------------------------------------------------------------------------------
// FP32 FMA peak probe (the denominator of the CUDA-core roofline; MEASURED_PEAKS.json has none)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nif_fma_probe_kernel(int iters, float* out) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  const float b = 1.000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}

extern "C" int nif_measure_fp32_peak(double* tflops) {
  if (!tflops) return NIF_E_BAD_ARG;
  float* d = nullptr;
  NIF_CUDA_CHECK(cudaMalloc(&d, 4));
  int dev = 0, sms = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int iters = 4096, blocks = sms * 8;
  cudaEvent_t e0, e1;
  NIF_CUDA_CHECK(cudaEventCreate(&e0));
  NIF_CUDA_CHECK(cudaEventCreate(&e1));
  nif_fma_probe_kernel<<<blocks, 256>>>(iters, d);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    NIF_CUDA_CHECK(cudaEventRecord(e0));
    nif_fma_probe_kernel<<<blocks, 256>>>(iters, d);
    NIF_CUDA_CHECK(cudaEventRecord(e1));
    NIF_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    NIF_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * (double)blocks * 256.0 * iters * 8.0 * 16.0;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return NIF_OK;
}
