// Fused forward on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-grade accuracy via FP16x3.
//
// Same mathematics as nif_fwd.cu:
//   pre[b][j] = sum_kappa zt[b][kappa] * ( omega * V[b][kappa][j] + C_m[kappa][j] ),   V = h @ M_m[kappa]
// with the shared-operand product V on the tensor cores:
//   A = h tile [128 rows x 64], per-row power-of-two scale, hi/lo fp16 split, written by the epilogue threads;
//   B = chunk of 2 latent coordinates [128 (kappa_l, j) x 64 (i)], per-slab scale, hi/lo split, streamed from the
//       packed image by cp.async.bulk (32 KB per chunk, 3 stages);
//   D = D1 [128 x 128] (hi*hi) and D2 [128 x 128] (cross terms), fp32 in TMEM.
// The per-row contraction over kappa (and the un-scaling) runs on the CUDA cores straight out of TMEM: thread r
// owns row r = TMEM lane r.
//
// One CTA per SM, persistent over PAIRS of 128-row tiles.  The two tiles of a pair share every staged weight
// chunk (halves the L2 traffic per row) and ping-pong on the tensor pipe: while the MMAs of one tile run, the
// other tile's epilogue warps drain its accumulators.
//   warps 0-3  epilogue of tile 0        warps 4-7  epilogue of tile 1
//   warp 8     MMA issuer (one elected lane) and TMEM owner (512 columns: 2 tiles x (D1 | D2))
//   warp 9     weight-stream producer
#include "nif_tc.cuh"

struct TcFwdArgs {
  long long B, total_pairs;
  const float *z, *x, *packed;
  float *u, *save;
};

#define TCF_THREADS 320
#define TCF_STAGES 3
#define TCF_STAGE_BYTES 32768u  // [hi | lo] x 16 KB

__host__ __device__ inline size_t tcf_smem_bytes(int KP, int si) {
  return 4 * (size_t)TC_TILE_BYTES + TCF_STAGES * (size_t)TCF_STAGE_BYTES + 2 * (size_t)KP * 128 * 4 +
         2 * (size_t)si * 128 * 4 + 256;
}

template <bool SAVE>
__global__ void __launch_bounds__(TCF_THREADS, 1) nif_tc_fwd_kernel(const Plan pl, const TcFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A_all = smem;                                  // tile t: hi at t*32K, lo at t*32K + 16K
  unsigned char* Bst = smem + 4 * TC_TILE_BYTES;                // [TCF_STAGES][hi | lo]
  float* zs_all = reinterpret_cast<float*>(Bst + TCF_STAGES * TCF_STAGE_BYTES);  // [2][KP][128]
  float* xs_all = zs_all + 2 * pl.KP * 128;                     // [2][si][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs_all + 2 * pl.si * 128);
  uint64_t* b_full = bars;                    // [TCF_STAGES]
  uint64_t* b_empty = bars + TCF_STAGES;      // [TCF_STAGES]
  uint64_t* t_full = bars + 2 * TCF_STAGES;   // [2]  accumulators of tile t ready
  uint64_t* t_empty = t_full + 2;             // [2]  accumulators of tile t drained
  uint64_t* a_ready = t_empty + 2;            // [2]  operand tile of tile t written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = pl.K, K1 = pl.K + 1, KP = pl.KP, NCH = pl.NCH, H = pl.H, n = pl.n, si = pl.si, so = pl.so;

  if (tid == 0) {
    for (int i = 0; i < TCF_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 128);
      mbar_init(&a_ready[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 8) tc_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  long long my_pairs = 0;
  if ((long long)blockIdx.x < a.total_pairs) my_pairs = (a.total_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == 9) {
    // ---------------- weight-stream producer ----------------
    if (lane == 0) {
      long long g = 0;
      const float* src0 = a.packed + pl.off_TCF;
      for (long long t = 0; t < my_pairs; ++t)
        for (int h = 0; h < H; ++h)
          for (int c = 0; c < NCH; ++c, ++g) {
            const int s = (int)(g % TCF_STAGES);
            mbar_wait(&b_empty[s], (uint32_t)(((g / TCF_STAGES) & 1) ^ 1));
            mbar_expect_tx(&b_full[s], TCF_STAGE_BYTES);
            bulk_g2s(Bst + s * TCF_STAGE_BYTES, src0 + ((long long)h * NCH + c) * NIF_TC_CHUNK_FLOATS,
                     TCF_STAGE_BYTES, &b_full[s]);
          }
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = tc_idesc_f16(128);
      uint64_t da_hi[2], da_lo[2];
      for (int t = 0; t < 2; ++t) {
        da_hi[t] = tc_make_desc(smem_u32(A_all + t * 2 * TC_TILE_BYTES));
        da_lo[t] = tc_make_desc(smem_u32(A_all + t * 2 * TC_TILE_BYTES + TC_TILE_BYTES));
      }
      long long g = 0, L = 0;
      for (long long p = 0; p < my_pairs; ++p)
        for (int h = 0; h < H; ++h, ++L)
          for (int c = 0; c < NCH; ++c, ++g) {
            const int s = (int)(g % TCF_STAGES);
            mbar_wait(&b_full[s], (uint32_t)((g / TCF_STAGES) & 1));
            const uint64_t db_hi = tc_make_desc(smem_u32(Bst + s * TCF_STAGE_BYTES));
            const uint64_t db_lo = tc_make_desc(smem_u32(Bst + s * TCF_STAGE_BYTES + TC_TILE_BYTES));
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (c == 0) mbar_wait(&a_ready[t], (uint32_t)(L & 1));
              mbar_wait(&t_empty[t], (uint32_t)((g & 1) ^ 1));
              tc_fence_after();
              const uint32_t d1 = tmem + (uint32_t)t * 256u;
              tc_mma_split_k64(d1, d1 + 128u, da_hi[t], da_lo[t], db_hi, db_lo, idesc);
              tc_commit(&t_full[t]);
            }
            tc_commit(&b_empty[s]);
          }
    }
  } else {
    // ---------------- epilogue warps: thread r <-> row r of tile wg <-> TMEM lane r ----------------
    const int wg = warp >> 2;  // tile of the pair
    const int r = tid & 127;
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)wg * 256u;
    unsigned char* A_hi = A_all + wg * 2 * TC_TILE_BYTES;
    unsigned char* A_lo = A_hi + TC_TILE_BYTES;
    float* zs = zs_all + wg * KP * 128;
    float* xs = xs_all + wg * si * 128;
    const float* C_all = a.packed + pl.off_C;
    const float* invB = a.packed + pl.off_TCS;
    const uint32_t row_off = (uint32_t)(r >> 3) * TC_SBO + (uint32_t)(r & 7) * 16u;
    long long g = 0;
    for (long long p = 0; p < my_pairs; ++p) {
      const long long row0 = ((blockIdx.x + p * gridDim.x) * 2 + wg) * 128;
      const long long b = row0 + r;
      const bool live = b < a.B;
      named_bar_sync(1 + wg, 128);  // this tile's threads are done with the previous zs / xs
      for (int idx = r; idx < 128 * K; idx += 128) {
        const int q = idx / K, kk = idx - q * K;
        zs[kk * 128 + q] = (row0 + q < a.B) ? __ldg(&a.z[(row0 + q) * K + kk]) : 0.f;
      }
      zs[K * 128 + r] = 1.f;
      for (int kk = K1; kk < KP; ++kk) zs[kk * 128 + r] = 0.f;
      for (int idx = r; idx < 128 * si; idx += 128) {
        const int q = idx / si, i = idx - q * si;
        xs[i * 128 + q] = (row0 + q < a.B) ? __ldg(&a.x[(row0 + q) * si + i]) : 0.f;
      }
      named_bar_sync(1 + wg, 128);

      float hcur[64];     // output of the layer being finished (input of the next one)
      float inv_a = 1.f;  // inverse of the power-of-two scale of this row's operand tile

      // activation / residual / stash for layer m given pre-activations in `pre`; result in hcur.
      auto finish_layer = [&](int m, float (&pre)[64]) {
        const int res = plan_res(pl, m);  // 0 or 1 on this path (res-blocks use the CUDA-core kernel)
        float* sh = a.save + (long long)m * a.B * 64 + b * 64;            // h_{m+1}
        float* sd = a.save + (long long)(H + 1 + m) * a.B * 64 + b * 64;  // d_m
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float hold[8];
          if (res == 1) {  // this layer's input, re-read from its operand tile ((hi + lo) * 2^-e)
            const uint4 ph = *reinterpret_cast<const uint4*>(A_hi + row_off + c * TC_LBO);
            const uint4 ql = *reinterpret_cast<const uint4*>(A_lo + row_off + c * TC_LBO);
            const uint32_t hw[4] = {ph.x, ph.y, ph.z, ph.w}, lw[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
              const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
              hold[2 * e] = (fh.x + fl.x) * inv_a;
              hold[2 * e + 1] = (fh.y + fl.y) * inv_a;
            }
          }
          float dch[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = 8 * c + e;
            float f, d;
            act_fd(pl.act, pre[j], f, d);
            float o = f;
            if (res == 1) o += hold[e];
            if (j >= n) { o = 0.f; d = 0.f; }
            hcur[j] = o;
            dch[e] = d;
          }
          if (SAVE && live) {
            *reinterpret_cast<float4*>(sh + 8 * c) = make_float4(hcur[8 * c], hcur[8 * c + 1], hcur[8 * c + 2], hcur[8 * c + 3]);
            *reinterpret_cast<float4*>(sh + 8 * c + 4) = make_float4(hcur[8 * c + 4], hcur[8 * c + 5], hcur[8 * c + 6], hcur[8 * c + 7]);
            *reinterpret_cast<float4*>(sd + 8 * c) = make_float4(dch[0], dch[1], dch[2], dch[3]);
            *reinterpret_cast<float4*>(sd + 8 * c + 4) = make_float4(dch[4], dch[5], dch[6], dch[7]);
          }
        }
      };

      // ---- layer 0 (si -> n) on the CUDA cores; weights are warp-uniform loads ----
      {
        float pre[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) pre[j] = 0.f;
        const float om = plan_omega(pl, 0);
        const float* M0 = a.packed + pl.off_M0;
        for (int kk = 0; kk < K1; ++kk) {
          const float zk = zs[kk * 128 + r];
          const float* cb = C_all + (long long)kk * 64;  // layer 0 bias rows
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float4 q = ldg4(cb + 4 * c);
            pre[4 * c] = fmaf(zk, q.x, pre[4 * c]); pre[4 * c + 1] = fmaf(zk, q.y, pre[4 * c + 1]);
            pre[4 * c + 2] = fmaf(zk, q.z, pre[4 * c + 2]); pre[4 * c + 3] = fmaf(zk, q.w, pre[4 * c + 3]);
          }
          for (int i = 0; i < si; ++i) {
            const float ai = zk * om * xs[i * 128 + r];
            const float* mw = M0 + ((long long)kk * si + i) * 64;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float4 q = ldg4(mw + 4 * c);
              pre[4 * c] = fmaf(ai, q.x, pre[4 * c]); pre[4 * c + 1] = fmaf(ai, q.y, pre[4 * c + 1]);
              pre[4 * c + 2] = fmaf(ai, q.z, pre[4 * c + 2]); pre[4 * c + 3] = fmaf(ai, q.w, pre[4 * c + 3]);
            }
          }
        }
        finish_layer(0, pre);
      }

      // ---- hidden layers on the tensor cores ----
      for (int m = 1; m <= H; ++m) {
        // operand tile for this layer.  Every MMA that read the previous contents has completed: this thread
        // observed the last t_full of the previous layer.
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) amax = fmaxf(amax, fabsf(hcur[j]));
        float sc_a;
        tc_row_scale(amax, sc_a, inv_a);
        tc_store_row_split(A_hi, A_lo, r, hcur, sc_a);
        fence_async_smem();
        mbar_arrive(&a_ready[wg]);
        const float om_inv = plan_omega(pl, m) * inv_a;
        const float* invBm = invB + (m - 1) * KP;
        float acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.f;
        for (int c = 0; c < NCH; ++c, ++g) {
          mbar_wait(&t_full[wg], (uint32_t)(g & 1));
          tc_fence_after();
#pragma unroll
          for (int kl = 0; kl < 2; ++kl) {
            const int kk = 2 * c + kl;
            const float zk = zs[kk * 128 + r];
            const float zo = zk * om_inv * __ldg(&invBm[kk]);
            const float* cb = C_all + ((long long)m * K1 + (kk < K1 ? kk : 0)) * 64;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {  // 32 columns at a time: D1 + D2, then the latent contraction
              float v1[32], v2[32];
              const uint32_t col = (uint32_t)(kl * 64 + hf * 32);
              tc_ld32(tm + col, v1);
              tc_ld32(tm + col + 128u, v2);
              tc_wait_ld();
              if (kk < K1) {
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                  const float4 cq = ldg4(cb + hf * 32 + 4 * q4);
                  const int j = hf * 32 + 4 * q4;
                  acc[j] = fmaf(zk, cq.x, fmaf(zo, v1[4 * q4] + v2[4 * q4], acc[j]));
                  acc[j + 1] = fmaf(zk, cq.y, fmaf(zo, v1[4 * q4 + 1] + v2[4 * q4 + 1], acc[j + 1]));
                  acc[j + 2] = fmaf(zk, cq.z, fmaf(zo, v1[4 * q4 + 2] + v2[4 * q4 + 2], acc[j + 2]));
                  acc[j + 3] = fmaf(zk, cq.w, fmaf(zo, v1[4 * q4 + 3] + v2[4 * q4 + 3], acc[j + 3]));
                }
              }
            }
          }
          tc_fence_before();
          mbar_arrive(&t_empty[wg]);
        }
        finish_layer(m, acc);
      }

      // ---- last layer (n -> so) on the CUDA cores ----
      {
        const float* ML = a.packed + pl.off_ML;
        const float* CL = C_all + (long long)(H + 1) * K1 * 64;
        float y[NIF_MAX_SO];
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c) y[c] = 0.f;
        for (int kk = 0; kk < K1; ++kk) {
          float sacc[NIF_MAX_SO];
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c) sacc[c] = (c < so) ? __ldg(&CL[(long long)kk * 64 + c]) : 0.f;
          const float* Mk = ML + (long long)kk * 64 * so;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            if (i < n) {
#pragma unroll
              for (int c = 0; c < NIF_MAX_SO; ++c)
                if (c < so) sacc[c] = fmaf(hcur[i], __ldg(&Mk[i * so + c]), sacc[c]);
            }
          }
          const float zk = zs[kk * 128 + r];
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c) y[c] = fmaf(zk, sacc[c], y[c]);
        }
        if (live) {
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c)
            if (c < so) a.u[b * so + c] = y[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 512);
}

template <bool SAVE>
static int launch_tcf(const Plan& pl, const TcFwdArgs& a, cudaStream_t st) {
  const size_t smem = tcf_smem_bytes(pl.KP, pl.si);
  auto kern = nif_tc_fwd_kernel<SAVE>;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = sms;
  if (grid > a.total_pairs) grid = a.total_pairs;
  if (grid < 1) return NIF_OK;
  kern<<<(unsigned)grid, TCF_THREADS, smem, st>>>(pl, a);
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// returns NIF_E_UNSUPPORTED (without setting an error) when the shape does not fit this kernel, so that the
// dispatcher can use the CUDA-core kernel instead
int nif_tc_forward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed, float* u,
                        float* save, cudaStream_t st) {
  if (!pl.tc || pl.NP != 64 || pl.H < 1 || pl.variant == NIF_VARIANT_SIREN_RES) return NIF_E_UNSUPPORTED;
  if (tcf_smem_bytes(pl.KP, pl.si) > 227 * 1024) return NIF_E_UNSUPPORTED;
  TcFwdArgs a;
  a.B = B;
  a.total_pairs = (B + 255) / 256;
  a.z = z; a.x = x; a.packed = packed; a.u = u; a.save = save;
  return save ? launch_tcf<true>(pl, a, st) : launch_tcf<false>(pl, a, st);
}
