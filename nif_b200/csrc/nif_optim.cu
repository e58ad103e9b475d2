// Optimisers of nif/optimizers beyond Adam, on the flat fp32 parameter / gradient / slot buffers (HBM-bound, one launch):
//   AdaBeliefOptimizer  nif/optimizers/external_optimizers.py:321-628  (_resource_apply_dense :458-528)
//   Lion                nif/optimizers/external_optimizers.py:631-735  (_resource_apply_dense :681-702)
//   gradient centralisation  nif/optimizers/gtcf.py:7-67 (get_centralized_gradients :27-32)
// Every step-dependent scalar (bias corrections, warm-up learning rate, the rectification term) is formed on the host in
// double, as the reference forms them in the variable dtype once per step.
#include <cmath>
#include "nif_common.cuh"

// g_scale multiplies the gradient first (data parallel: 1 / world_size); l1 / l2: Keras kernel regularisers folded in
__global__ void __launch_bounds__(256) nif_adabelief_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                            float* __restrict__ m, float* __restrict__ v,
                                                            float* __restrict__ vhat, float lr_t, float b1, float b2,
                                                            float eps, float inv_bc1, float inv_bc2, float r_t, int mode,
                                                            float wd, float l1, float l2, float gs) {
  // mode: 0 = m_corr / (v_corr + eps)   1 = rectified, SMA above threshold: r_t * that   2 = rectified, below: m_corr
  const long long stride = 256LL * gridDim.x;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += stride) {
    float pp = p[i], gg = g[i] * gs;
    if (l1 != 0.f) gg += l1 * (pp > 0.f ? 1.f : (pp < 0.f ? -1.f : 0.f));
    if (l2 != 0.f) gg += 2.f * l2 * pp;
    const float mt = b1 * m[i] + (1.f - b1) * gg;
    const float d = gg - mt;
    float vt = b2 * v[i] + (1.f - b2) * d * d + eps;
    m[i] = mt;
    v[i] = vt;
    if (vhat) { vt = fmaxf(vhat[i], vt); vhat[i] = vt; }
    const float m_corr = mt * inv_bc1;
    const float v_corr = sqrtf(vt * inv_bc2);
    float upd = mode == 2 ? m_corr : m_corr / (v_corr + eps);
    if (mode == 1) upd *= r_t;
    if (wd != 0.f) upd += wd * pp;
    p[i] = pp - lr_t * upd;
  }
}

__global__ void __launch_bounds__(256) nif_lion_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float lr, float b1, float b2, float wd,
                                                       float l1, float l2, float gs) {
  const long long stride = 256LL * gridDim.x;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += stride) {
    float pp = p[i], gg = g[i] * gs;
    if (l1 != 0.f) gg += l1 * (pp > 0.f ? 1.f : (pp < 0.f ? -1.f : 0.f));
    if (l2 != 0.f) gg += 2.f * l2 * pp;
    const float mm = m[i];
    const float c = mm * b1 + gg * (1.f - b1);
    const float sgn = c > 0.f ? 1.f : (c < 0.f ? -1.f : 0.f);
    p[i] = pp - lr * (sgn + pp * wd);
    m[i] = mm * b2 + gg * (1.f - b2);
  }
}

// grad[rows][cols] -= mean over rows (every axis but the last of a rank >= 2 gradient); one thread per column
__global__ void __launch_bounds__(256) nif_centralize_kernel(long long rows, long long cols, float* __restrict__ g) {
  const long long c = blockIdx.x * 256LL + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (long long r = 0; r < rows; ++r) s += g[r * cols + c];
  const float mean = s / (float)rows;
  for (long long r = 0; r < rows; ++r) g[r * cols + c] -= mean;
}

static unsigned blocks_for(long long n) {
  long long nblk = (n + 255) / 256;
  if (nblk < 1) nblk = 1;
  if (nblk > 148 * 16) nblk = 148 * 16;
  return (unsigned)nblk;
}

extern "C" int nif_adabelief_step(int64_t n, float* p, const float* g, float* m, float* v, float* vhat, double lr_t,
                                  double b1, double b2, double eps, int64_t t, int32_t rectify, double sma_threshold,
                                  double weight_decay, float l1, float l2, float g_scale, void* stream) {
  if (n < 0 || t < 1) { nif_set_error("nif_adabelief_step: n=%lld t=%lld", (long long)n, (long long)t); return NIF_E_BAD_ARG; }
  if (n == 0) return NIF_OK;
  if (!p || !g || !m || !v) { nif_set_error("nif_adabelief_step: null buffer"); return NIF_E_BAD_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double b1p = std::pow(b1, (double)t), b2p = std::pow(b2, (double)t);
  const double sma_inf = 2.0 / (1.0 - b2) - 1.0;
  const double sma_t = sma_inf - 2.0 * (double)t * b2p / (1.0 - b2p);
  int mode = 0;
  double r_t = 1.0;
  if (rectify) {
    if (sma_t >= sma_threshold) {
      mode = 1;
      r_t = std::sqrt((sma_t - 4.0) / (sma_inf - 4.0) * (sma_t - 2.0) / (sma_inf - 2.0) * sma_inf / sma_t);
    } else {
      mode = 2;
    }
  }
  { NIF_PROF("nif_adabelief_kernel", st);
    nif_adabelief_kernel<<<blocks_for(n), 256, 0, st>>>(n, p, g, m, v, vhat, (float)lr_t, (float)b1, (float)b2, (float)eps,
                                                        (float)(1.0 / (1.0 - b1p)), (float)(1.0 / (1.0 - b2p)), (float)r_t, mode,
                                                        (float)weight_decay, l1, l2, g_scale); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

extern "C" int nif_lion_step(int64_t n, float* p, const float* g, float* m, double lr, double b1, double b2, double wd,
                             float l1, float l2, float g_scale, void* stream) {
  if (n < 0) { nif_set_error("nif_lion_step: n=%lld", (long long)n); return NIF_E_BAD_ARG; }
  if (n == 0) return NIF_OK;
  if (!p || !g || !m) { nif_set_error("nif_lion_step: null buffer"); return NIF_E_BAD_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { NIF_PROF("nif_lion_kernel", st);
    nif_lion_kernel<<<blocks_for(n), 256, 0, st>>>(n, p, g, m, (float)lr, (float)b1, (float)b2, (float)wd, l1, l2, g_scale); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

extern "C" int nif_centralize_gradient(int64_t rows, int64_t cols, float* g, void* stream) {
  if (rows < 0 || cols < 0) { nif_set_error("nif_centralize_gradient: rows=%lld cols=%lld", (long long)rows, (long long)cols); return NIF_E_BAD_ARG; }
  if (rows == 0 || cols == 0) return NIF_OK;
  if (!g) { nif_set_error("nif_centralize_gradient: null buffer"); return NIF_E_BAD_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { NIF_PROF("nif_centralize_kernel", st);
    nif_centralize_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(rows, cols, g); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Data-parallel update fused with its collective over NVSwitch multicast memory (NVLS): ONE kernel per step does
//   reduce-scatter   g = multimem.ld_reduce.add over every replica's gradient buffer, for this rank's slice only
//   Adam             on that slice (each rank keeps the moments of its own slice: the optimiser state is sharded)
//   all-gather       multimem.st of the updated parameters into every replica's parameter buffer
// replacing  all-reduce(gradient) -> Adam on every rank  (tf.distribute.MirroredStrategy in the reference, README.md:39-49;
// NCCL all-reduce + nif_adam_step in the plain data-parallel path).  Every parameter value is computed exactly once, so the
// replicas stay bit-identical by construction; each rank reads 1/N of the gradient and does 1/N of the update.
// The caller brackets the launch with cross-rank barriers (gradients complete before; parameters visible after).
// p_mc / g_mc: MULTICAST addresses of the symmetric parameter / gradient buffers; p, m, v: this rank's own buffers.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mc_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void adam_one_(float& p, float g, float& m, float& v, float alpha, float omb1, float omb2, float eps,
                                          float l1, float l2, float gs) {
  g *= gs;
  if (l1 != 0.f) g += l1 * (p > 0.f ? 1.f : (p < 0.f ? -1.f : 0.f));
  if (l2 != 0.f) g += 2.f * l2 * p;
  m += (g - m) * omb1;
  v += (g * g - v) * omb2;
  p -= alpha * m / (sqrtf(v) + eps);
}
__global__ void __launch_bounds__(256) nif_adam_multimem_kernel(long long i4_begin, long long i4_end, float* __restrict__ p_mc,
                                                                const float* __restrict__ g_mc, const float* __restrict__ p,
                                                                float* __restrict__ m, float* __restrict__ v, float alpha,
                                                                const float* __restrict__ alpha_dev, float omb1, float omb2,
                                                                float eps, float l1, float l2, float gs) {
  if (alpha_dev) alpha = __ldg(alpha_dev);
  const long long stride = 256LL * gridDim.x;
  for (long long i = i4_begin + blockIdx.x * 256LL + threadIdx.x; i < i4_end; i += stride) {
    const float4 gg = mc_ld_reduce_add(g_mc + 4 * i);
    float4 pp = reinterpret_cast<const float4*>(p)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one_(pp.x, gg.x, mm.x, vv.x, alpha, omb1, omb2, eps, l1, l2, gs);
    adam_one_(pp.y, gg.y, mm.y, vv.y, alpha, omb1, omb2, eps, l1, l2, gs);
    adam_one_(pp.z, gg.z, mm.z, vv.z, alpha, omb1, omb2, eps, l1, l2, gs);
    adam_one_(pp.w, gg.w, mm.w, vv.w, alpha, omb1, omb2, eps, l1, l2, gs);
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    mc_st(p_mc + 4 * i, pp);
  }
  __threadfence_system();  // the multicast stores are performed at every replica before the kernel retires
}

extern "C" int nif_adam_step_multimem(int64_t n, int32_t rank, int32_t world, float* p_mc, const float* g_mc, const float* p,
                                      float* m, float* v, double lr, const float* alpha_dev, double b1, double b2, double eps,
                                      int64_t t, float l1, float l2, float g_scale, void* stream) {
  if (n < 0 || (n & 3) || world < 1 || rank < 0 || rank >= world) {
    nif_set_error("nif_adam_step_multimem: n=%lld (must be a multiple of 4) rank=%d world=%d", (long long)n, rank, world);
    return NIF_E_BAD_ARG;
  }
  if (n == 0) return NIF_OK;
  if (!p_mc || !g_mc || !p || !m || !v) { nif_set_error("nif_adam_step_multimem: null buffer"); return NIF_E_BAD_ARG; }
  if (!alpha_dev && t < 1) { nif_set_error("nif_adam_step_multimem: t=%lld", (long long)t); return NIF_E_BAD_ARG; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n4 = n / 4, chunk = (n4 + world - 1) / world;
  const long long b = (long long)rank * chunk, e = b + chunk < n4 ? b + chunk : n4;
  if (e <= b) return NIF_OK;
  const double alpha = alpha_dev ? 0.0 : lr * std::sqrt(1.0 - std::pow(b2, (double)t)) / (1.0 - std::pow(b1, (double)t));
  long long nblk = (e - b + 255) / 256;
  if (nblk > 148 * 4) nblk = 148 * 4;
  { NIF_PROF("nif_adam_multimem_kernel", st);
    nif_adam_multimem_kernel<<<(unsigned)nblk, 256, 0, st>>>(b, e, p_mc, g_mc, p, m, v, (float)alpha, alpha_dev, (float)(1.0 - b1),
                                                             (float)(1.0 - b2), (float)eps, l1, l2, g_scale); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
