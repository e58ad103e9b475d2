// Elementwise stages of a WIDE ParameterNet trunk under the mixed_bfloat16 policy (units > 64: the tcgen05 trunk kernels
// of nif_trunk_tc.cu stop at 64).  The Dense products of such a trunk are plain library GEMMs on bf16 operands; everything
// between two GEMMs of a layer
//     h_out = h_in + act( float(y) + b ),   y = bf16(h_in) @ bf16(W)              (MLP_SimpleShortCut, nif/layers/mlp.py:148-160;
//                                                                                  first Dense: no h_in, nif/model.py:326-343)
// and of its reverse pass
//     t = dh_in + float(p)            p = pending product g_above @ W_above^T of the layer above (bf16), or absent
//     g = bf16( t * act'(float(y) + b) ),   db = sum_b t * act'(...)   (fp32),    dh_out = t  (the shortcut's share)
// is ONE kernel here instead of five to seven framework launches (cast, add, activation, add, cast; their backward
// twins), with the same rounding points: bf16 at the GEMM operands and products, fp32 everywhere else.  The
// pre-activation is not stored: it is recomputed from the bf16 product, which the forward keeps anyway.
#include <cuda_bf16.h>
#include "nif_common.cuh"

struct TrunkEwFwdArgs {
  long long B;
  int n, act;
  const __nv_bfloat16* y;
  const float *bias, *h_in;
  float* h_out;
  __nv_bfloat16* h_out_bf;
};

__device__ __forceinline__ void bf8_to_float(const uint4& q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {  // bf16 -> fp32 is a 16-bit shift
    v[2 * e] = __uint_as_float(w[e] << 16);
    v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 float8_to_bf(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    w[e] = *reinterpret_cast<const uint32_t*>(&p);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// one thread = 8 consecutive columns of one row per iteration; a warp reads 512 B of y and 1 KB of h_in per access
__global__ void __launch_bounds__(256) nif_trunk_ew_fwd_kernel(const TrunkEwFwdArgs a) {
  const int nv = a.n >> 3;
  const long long total = a.B * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % nv);
    const long long off = idx * 8;
    float y[8], b[8], o[8];
    bf8_to_float(*reinterpret_cast<const uint4*>(a.y + off), y);
    const float4 b0 = ldg4(a.bias + 8 * cv), b1 = ldg4(a.bias + 8 * cv + 4);
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float f, d;
      act_fd(a.act, y[e] + b[e], f, d);
      o[e] = f;
    }
    if (a.h_in) {
      const float4 h0 = ldg4(a.h_in + off), h1 = ldg4(a.h_in + off + 4);
      o[0] += h0.x; o[1] += h0.y; o[2] += h0.z; o[3] += h0.w; o[4] += h1.x; o[5] += h1.y; o[6] += h1.z; o[7] += h1.w;
    }
    *reinterpret_cast<float4*>(a.h_out + off) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(a.h_out + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
    if (a.h_out_bf) *reinterpret_cast<uint4*>(a.h_out_bf + off) = float8_to_bf(o);
  }
}

struct TrunkEwBwdArgs {
  long long B;
  int n, act;
  const __nv_bfloat16 *y, *pend;
  const float *bias, *dh_in;
  float* dh_out;
  __nv_bfloat16* g;
  float* db_part;  // [gridDim.x][n]
};

// block = (n / 8) column vectors x (256 / (n / 8)) row lanes; a block walks its rows with every thread keeping the
// running sums of its 8 columns, which the row lanes then add up in shared memory (deterministic, no atomics)
__global__ void __launch_bounds__(256) nif_trunk_ew_bwd_kernel(const TrunkEwBwdArgs a) {
  __shared__ float red[256 * 8];
  const int nv = a.n >> 3;
  const int lanes = 256 / nv;
  const int cv = threadIdx.x % nv, rl = threadIdx.x / nv;
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  float b[8];
  {
    const float4 b0 = ldg4(a.bias + 8 * cv), b1 = ldg4(a.bias + 8 * cv + 4);
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
  if (rl < lanes) {
    for (long long row = (long long)blockIdx.x * lanes + rl; row < a.B; row += (long long)gridDim.x * lanes) {
      const long long off = row * a.n + 8 * cv;
      float t[8], y[8], gq[8];
      if (a.dh_in) {  // (plain loads: dh_out may alias dh_in, so this is not read-only data)
        const float4 t0 = *reinterpret_cast<const float4*>(a.dh_in + off), t1 = *reinterpret_cast<const float4*>(a.dh_in + off + 4);
        t[0] = t0.x; t[1] = t0.y; t[2] = t0.z; t[3] = t0.w; t[4] = t1.x; t[5] = t1.y; t[6] = t1.z; t[7] = t1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = 0.f;
      }
      if (a.pend) {
        float p[8];
        bf8_to_float(*reinterpret_cast<const uint4*>(a.pend + off), p);
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] += p[e];
      }
      bf8_to_float(*reinterpret_cast<const uint4*>(a.y + off), y);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float f, d;
        act_fd(a.act, y[e] + b[e], f, d);
        gq[e] = t[e] * d;
        s[e] += gq[e];
      }
      *reinterpret_cast<uint4*>(a.g + off) = float8_to_bf(gq);
      if (a.dh_out) {
        *reinterpret_cast<float4*>(a.dh_out + off) = make_float4(t[0], t[1], t[2], t[3]);
        *reinterpret_cast<float4*>(a.dh_out + off + 4) = make_float4(t[4], t[5], t[6], t[7]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x * 8 + e] = s[e];
  __syncthreads();
  if (threadIdx.x < a.n) {  // column j = 8 cv' + e: add the row lanes
    const int cvj = threadIdx.x >> 3, e = threadIdx.x & 7;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[(l * nv + cvj) * 8 + e];
    a.db_part[(long long)blockIdx.x * a.n + threadIdx.x] = acc;
  }
}

// db[j] = sum over the blocks' partials (fixed order): a block = 32 columns x 32 groups of partials, a warp reads 128
// contiguous bytes per partial row; the group sums are added in shared memory
__global__ void __launch_bounds__(1024) nif_trunk_ew_db_kernel(int nblk, int n, const float* part, float* db) {
  __shared__ float red[32][33];
  const int c = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + c;
  float acc = 0.f;
  if (j < n)
    for (int k = grp; k < nblk; k += 32) acc += part[(long long)k * n + j];
  red[grp][c] = acc;
  __syncthreads();
  if (grp == 0 && j < n) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) t += red[q][c];
    db[j] = t;
  }
}

#define NIF_TRUNK_EW_BLOCKS 592  // 4 per SM

static bool ew_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
#define EW_REQUIRE(p)                                                                        \
  do {                                                                                       \
    if (!(p) || !ew_aligned16(p)) {                                                          \
      nif_set_error("%s: argument `%s` is null or not 16-byte aligned", __func__, #p);       \
      return NIF_E_BAD_ARG;                                                                  \
    }                                                                                        \
  } while (0)
#define EW_OPTIONAL(p)                                                                       \
  do {                                                                                       \
    if ((p) && !ew_aligned16(p)) {                                                           \
      nif_set_error("%s: argument `%s` is not 16-byte aligned", __func__, #p);               \
      return NIF_E_BAD_ARG;                                                                  \
    }                                                                                        \
  } while (0)

static int ew_check_shape(const char* fn, long long B, int n, int act) {
  if (B < 0 || n < 8 || n > 256 || (n & 7) || (256 % (n >> 3)) != 0) {
    nif_set_error("%s: B=%lld, n=%d (n must be a multiple of 8 that divides 2048, at most 256)", fn, B, n);
    return NIF_E_BAD_ARG;
  }
  if (act < NIF_ACT_LINEAR || act > NIF_ACT_SIGMOID || act == NIF_ACT_SINE) {
    nif_set_error("%s: activation %d is not a ParameterNet MLP activation", fn, act);
    return NIF_E_BAD_ARG;
  }
  return NIF_OK;
}

extern "C" int nif_trunk_ew_forward(int64_t B, int32_t n, int32_t act, const void* y_bf16, const float* bias,
                                    const float* h_in, float* h_out, void* h_out_bf16, void* stream) {
  const int rc = ew_check_shape(__func__, B, n, act);
  if (rc) return rc;
  if (B == 0) return NIF_OK;
  EW_REQUIRE(y_bf16); EW_REQUIRE(bias); EW_REQUIRE(h_out); EW_OPTIONAL(h_in); EW_OPTIONAL(h_out_bf16);
  TrunkEwFwdArgs a;
  a.B = B; a.n = n; a.act = act;
  a.y = static_cast<const __nv_bfloat16*>(y_bf16); a.bias = bias; a.h_in = h_in; a.h_out = h_out;
  a.h_out_bf = static_cast<__nv_bfloat16*>(h_out_bf16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = (B * (n >> 3) + 255) / 256;
  if (blocks > 8 * 148) blocks = 8 * 148;
  { NIF_PROF("nif_trunk_ew_fwd_kernel", st); nif_trunk_ew_fwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

extern "C" int nif_trunk_ew_ws_floats(int32_t n) { return NIF_TRUNK_EW_BLOCKS * n; }

extern "C" int nif_trunk_ew_backward(int64_t B, int32_t n, int32_t act, const void* y_bf16, const float* bias,
                                     const float* dh_in, const void* pend_bf16, float* dh_out, void* g_bf16, float* db,
                                     float* ws, void* stream) {
  const int rc = ew_check_shape(__func__, B, n, act);
  if (rc) return rc;
  if (B == 0) return NIF_OK;
  EW_REQUIRE(y_bf16); EW_REQUIRE(bias); EW_REQUIRE(g_bf16); EW_REQUIRE(db); EW_REQUIRE(ws);
  EW_OPTIONAL(dh_in); EW_OPTIONAL(pend_bf16); EW_OPTIONAL(dh_out);
  TrunkEwBwdArgs a;
  a.B = B; a.n = n; a.act = act;
  a.y = static_cast<const __nv_bfloat16*>(y_bf16); a.pend = static_cast<const __nv_bfloat16*>(pend_bf16);
  a.bias = bias; a.dh_in = dh_in; a.dh_out = dh_out; a.g = static_cast<__nv_bfloat16*>(g_bf16); a.db_part = ws;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lanes = 256 / (n >> 3);
  long long blocks = (B + lanes - 1) / lanes;
  if (blocks > NIF_TRUNK_EW_BLOCKS) blocks = NIF_TRUNK_EW_BLOCKS;
  { NIF_PROF("nif_trunk_ew_bwd_kernel", st); nif_trunk_ew_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  { NIF_PROF("nif_trunk_ew_db_kernel", st); nif_trunk_ew_db_kernel<<<(n + 31) / 32, 1024, 0, st>>>((int)blocks, n, ws, db); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
