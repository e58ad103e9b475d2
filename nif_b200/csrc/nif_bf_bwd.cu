// Reverse pass on the tensor cores with bf16 operands (dtype_compute = 1) -- the mirror image of nif_bf_fwd.cu.
//
// nif_bf_bwd_data_kernel, per 128-row tile, two threads per row (thread (r, half): row r, columns half*NP/2 ..):
//   prologue   dh_{H+1}[i] = sum_c du[c] * (zt @ BLT[c])[i]                       LT(c):  A = zt tile,       N = NP
//              dz[kappa]  += sum_c du[c] * ((h_{H+1} @ XL[c])[kappa] + CL[kappa][c])   LZ(c):  A = h_{H+1} tile,  N = KZ
//   m = H..0   da_m = dh_{m+1} * d_m   -> global (tiled) and the A tile
//              dz[kappa]  += (da_m @ BC[m])[kappa]                                 BC(m):  bias rows,         N = KZ
//       m >= 1 T = da_m @ WB[m-1][chunk];  dh_m[i] += omega zt[kappa] T[kappa][i];  dz[kappa] += omega sum_i T[kappa][i] h_m[i]
//       m == 0 dz[kappa]  += omega x[i] (da_0 @ B0[i])[kappa]                      B0(i):  first matrix,      N = KZ
// so every dz term, thin or not, is a tensor-core chunk over operand tiles that are in shared memory anyway.
// The halves of a row keep separate dz accumulators in shared memory ([2][K+1][128], thread-private entries: no atomics,
// deterministic) and are added when the row is written.
//
// nif_bf_bwd_weight_kernel: the batch-reduction GEMM dM_m[kappa][i][j] = omega sum_b zt[b][kappa] h_m[b][i] da_m[b][j]
// with both operands generated on the fly (MN-major bf16), accumulators resident in TMEM over the batch split.
#include "nif_bf.cuh"

NIF_TRACE_READER(nif_debug_read_trace_bfe)

struct BfBwdArgs {
  long long B, total_tiles;
  const float *z, *x, *packed, *save, *du;
  float* da;  // [(H+1)] tiled slots
  float* dz;  // [B][K]
  int nst;
};

#define BFB_THREADS 384

size_t nif_bfb_smem_fixed(const Plan& pl) {
  return (size_t)128 * pl.NP * 2 + (size_t)128 * pl.KZ * 2 + (size_t)2 * (pl.K + 1) * 128 * 4 + 256;
}

template <int NP>
__global__ void __launch_bounds__(BFB_THREADS, 1) nif_bf_bwd_data_kernel(const Plan pl, const BfBwdArgs a) {
  constexpr int CH = NP / 2;
  constexpr int CK = 128 / NP;
  constexpr uint32_t SBO_A = (NP / 8) * 128u;
  constexpr uint32_t MAIN_BYTES = 128u * NP * 2u;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A_tile = smem;
  unsigned char* Z_tile = smem + 128 * NP * 2;
  const uint32_t zbytes = 128u * (uint32_t)pl.KZ * 2u;
  const uint32_t sbo_z = (uint32_t)(pl.KZ / 8) * 128u;
  unsigned char* Bst = Z_tile + zbytes;
  float* dzs_all = reinterpret_cast<float*>(Bst + a.nst * BF_STAGE_BYTES);  // [2][K+1][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dzs_all + 2 * (pl.K + 1) * 128);
  uint64_t* b_full = bars;
  uint64_t* b_empty = bars + 8;
  uint64_t* t_full = bars + 16;
  uint64_t* t_empty = bars + 18;
  uint64_t* a_ready = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = pl.K, K1 = pl.K + 1, H = pl.H, si = pl.si, so = pl.so, KZ = pl.KZ;
  const int NCHW = bf_nchw(pl);
  const uint32_t nst = (uint32_t)a.nst;
  const uint32_t small_bytes = (uint32_t)NP * (uint32_t)KZ * 2u;

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 256); }
    mbar_init(&a_ready[0], 256);
    mbar_fence_init();
  }
  if (warp == 8) tc_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  long long my_tiles = 0;
  if ((long long)blockIdx.x < a.total_tiles) my_tiles = (a.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp >= 8) {
    tc_reg_dec<56>();
    if (warp == 9) {
      if (lane == 0) {  // weight-stream producer
        uint32_t s = 0, ph = 0;
        auto put = [&](const float* src, uint32_t bytes) {
          mbar_wait(&b_empty[s], ph ^ 1u);
          mbar_expect_tx(&b_full[s], bytes);
          bulk_g2s(Bst + s * BF_STAGE_BYTES, src, bytes, &b_full[s]);
          if (++s == nst) { s = 0; ph ^= 1u; }
        };
        const float* wb = a.packed + pl.off_WB;
        const float* wx = a.packed + pl.off_WX;
        auto small = [&](int T) { put(wx + (long long)T * bf_small_floats(pl), small_bytes); };
        for (long long t = 0; t < my_tiles; ++t) {
          for (int c = 0; c < so; ++c) small(bf_t_blt(pl, c));
          for (int c = 0; c < so; ++c) small(bf_t_xl(pl, c));
          for (int m = H; m >= 0; --m) {
            small(bf_t_bc(pl, m));
            if (m >= 1) {
              for (int c = 0; c < NCHW; ++c) put(wb + ((long long)(m - 1) * NCHW + c) * bf_chunk_floats(pl), MAIN_BYTES);
            } else {
              for (int i = 0; i < si; ++i) small(bf_t_b0(pl, i));
            }
          }
        }
      }
    } else if (warp == 8) {
      // MMA issuer: warp-uniform loop, one elected lane issues
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint64_t da = bf_make_desc(smem_u32(A_tile), SBO_A);
      const uint64_t dz = bf_make_desc(smem_u32(Z_tile), sbo_z);
      uint32_t g = 0, s = 0, ph = 0, ar = 0;
      auto chunk = [&](bool a_main, bool wait_a, int ksteps, uint32_t b_sbo, int N) {
        const uint32_t as = g & 1u;
        if (wait_a) { mbar_wait(&a_ready[0], ar & 1u); ++ar; }
        mbar_wait(&t_empty[as], ((g >> 1) & 1u) ^ 1u);
        mbar_wait(&b_full[s], ph);
        tc_fence_after();
        const uint64_t db = bf_make_desc(smem_u32(Bst + s * BF_STAGE_BYTES), b_sbo);
        const uint64_t dA = a_main ? da : dz;
        const uint32_t d = tmem_u + as * 128u;
        const uint32_t idesc = bf_idesc(N);
        if (tc_elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks)
            tc_mma_f16(d, dA + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : 0u);
          tc_commit(&t_full[as]);
          tc_commit(&b_empty[s]);
        }
        __syncwarp();
        ++g;
        if (++s == nst) { s = 0; ph ^= 1u; }
      };
      for (long long t = 0; t < my_tiles; ++t) {
        for (int c = 0; c < so; ++c) chunk(false, c == 0, KZ / 16, sbo_z, NP);   // LT(c): zt tile (+ h tile) published
        for (int c = 0; c < so; ++c) chunk(true, false, NP / 16, SBO_A, KZ);    // LZ(c)
        for (int m = H; m >= 0; --m) {
          chunk(true, true, NP / 16, SBO_A, KZ);                                // BC(m): da_m published
          if (m >= 1) {
            for (int c = 0; c < NCHW; ++c) chunk(true, false, NP / 16, SBO_A, 128);
          } else {
            for (int i = 0; i < si; ++i) chunk(true, false, NP / 16, SBO_A, KZ);
          }
        }
      }
    }
  } else {
    tc_reg_inc<224>();
    const int half = warp >> 2;
    const int r = tid & 127;
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t row_off = (uint32_t)(r >> 3) * SBO_A + (uint32_t)(r & 7) * 16u;
    const uint32_t zrow_off = (uint32_t)(r >> 3) * sbo_z + (uint32_t)(r & 7) * 16u;
    const long long slot_floats = bf_slot_floats(a.B, NP);
    float* dzs = dzs_all + (long long)half * K1 * 128 + r;  // dzs[kk * 128]: this thread's private accumulators
    const float* C_all = a.packed + pl.off_C;
    const float* CL = C_all + (long long)(H + 1) * K1 * pl.NP;
    const int kh = KZ / 2, k0 = half * kh;  // this thread's latent coordinates in the N = KZ chunks
    uint32_t g = 0;
    for (long long t = 0; t < my_tiles; ++t) {
      const long long b = (blockIdx.x + t * gridDim.x) * 128 + r;
      const bool live = b < a.B;
      const float* zrow = a.z + (live ? b : 0) * K;
      auto zt_at = [&](int kk) -> float { return kk < K ? (live ? __ldg(zrow + kk) : 0.f) : (kk == K ? 1.f : 0.f); };
      const long long trow = bf_tiled_row(b, NP) + (long long)(half * (CH / 4)) * 128;  // this thread's quads of a slot

      auto chunk_begin = [&]() -> uint32_t {
        mbar_wait(&t_full[g & 1u], (g >> 1) & 1u);
        tc_fence_after();
        return tm + (g & 1u) * 128u;
      };
      auto chunk_end = [&]() {
        tc_fence_before();
        mbar_arrive(&t_empty[g & 1u]);
        ++g;
      };
      // an N = KZ chunk: dzs[kappa] += coef * (D[kappa] + add[kappa]) for this thread's half of the latent coordinates
      auto drain_kz = [&](float coef, const float* add, int add_stride) {
        const uint32_t td = chunk_begin();
        for (int q = 0; q < kh; q += 8) {
          float v[8];
          tc_ld8(td + (uint32_t)(k0 + q), v);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int kk = k0 + q + e;
            if (kk < K1) dzs[kk * 128] += coef * (v[e] + (add ? __ldg(add + (long long)kk * add_stride) : 0.f));
          }
        }
        chunk_end();
      };

      // ---- prologue: operand tiles zt and h_{H+1}; private dz accumulators cleared ----
      for (int kk = 0; kk < K1; ++kk) dzs[kk * 128] = 0.f;
      {
        const int ng = KZ / 8;
        for (int c = half; c < ng; c += 2) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = bf_pack2(zt_at(8 * c + 2 * e), zt_at(8 * c + 2 * e + 1));
          *reinterpret_cast<uint4*>(Z_tile + zrow_off + c * 128) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        const float* hsrc = a.save + (long long)H * slot_floats + trow;  // h_{H+1}
#pragma unroll
        for (int c = 0; c < CH / 8; ++c) {
          float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
          if (live) { p0 = ldg4(hsrc + (2 * c) * 128); p1 = ldg4(hsrc + (2 * c + 1) * 128); }
          *reinterpret_cast<uint4*>(A_tile + row_off + (uint32_t)(half * (CH / 8) + c) * 128u) =
              make_uint4(bf_pack2(p0.x, p0.y), bf_pack2(p0.z, p0.w), bf_pack2(p1.x, p1.y), bf_pack2(p1.z, p1.w));
        }
        fence_async_smem();
        mbar_arrive(&a_ready[0]);
      }
      float dy[NIF_MAX_SO];
#pragma unroll
      for (int c = 0; c < NIF_MAX_SO; ++c) dy[c] = (c < so && live) ? __ldg(&a.du[b * so + c]) : 0.f;

      float acc[CH];  // dh_{m+1}, this thread's columns
      // ---- last matrix: dh_{H+1} ----
      for (int c = 0; c < so; ++c) {
        const uint32_t td = chunk_begin();
        float dyc = 0.f;
#pragma unroll
        for (int cc = 0; cc < NIF_MAX_SO; ++cc) if (cc == c) dyc = dy[cc];
#pragma unroll
        for (int q = 0; q < CH / 32; ++q) {
          float v[32];
          tc_ld32(td + (uint32_t)(half * CH + q * 32), v);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[q * 32 + e] = c == 0 ? dyc * v[e] : fmaf(dyc, v[e], acc[q * 32 + e]);
        }
        chunk_end();
      }
      // ---- last matrix: its dz terms ----
      for (int c = 0; c < so; ++c) {
        float dyc = 0.f;
#pragma unroll
        for (int cc = 0; cc < NIF_MAX_SO; ++cc) if (cc == c) dyc = dy[cc];
        drain_kz(dyc, CL + c, pl.NP);
      }

      // ---- layers H .. 0 ----
#pragma unroll 1
      for (int m = H; m >= 0; --m) {
        {  // da_m = dh_{m+1} * d_m -> global (tiled) and the operand tile
          const float* dsv = a.save + (long long)(H + 1 + m) * slot_floats + trow;
          float* dag = a.da + (long long)m * slot_floats + trow;
#pragma unroll
          for (int c = 0; c < CH / 8; ++c) {
            float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0;
            if (live) { d0 = ldg4(dsv + (2 * c) * 128); d1 = ldg4(dsv + (2 * c + 1) * 128); }
            const float v0 = acc[8 * c] * d0.x, v1 = acc[8 * c + 1] * d0.y, v2 = acc[8 * c + 2] * d0.z, v3 = acc[8 * c + 3] * d0.w;
            const float v4 = acc[8 * c + 4] * d1.x, v5 = acc[8 * c + 5] * d1.y, v6 = acc[8 * c + 6] * d1.z, v7 = acc[8 * c + 7] * d1.w;
            if (live) {
              *reinterpret_cast<float4*>(dag + (2 * c) * 128) = make_float4(v0, v1, v2, v3);
              *reinterpret_cast<float4*>(dag + (2 * c + 1) * 128) = make_float4(v4, v5, v6, v7);
            }
            *reinterpret_cast<uint4*>(A_tile + row_off + (uint32_t)(half * (CH / 8) + c) * 128u) =
                make_uint4(bf_pack2(v0, v1), bf_pack2(v2, v3), bf_pack2(v4, v5), bf_pack2(v6, v7));
          }
          fence_async_smem();
          mbar_arrive(&a_ready[0]);
        }
        // this layer's input h_m (this thread's columns) for the dz dot products: fetched before the bias-row chunk is
        // drained, so that the loads fly behind its wait
        float hm[CH];
        if (m >= 1) {
          const float* hsrc = a.save + (long long)(m - 1) * slot_floats + trow;
#pragma unroll
          for (int c = 0; c < CH / 4; ++c) {
            float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) q4 = ldg4(hsrc + c * 128);
            hm[4 * c] = q4.x; hm[4 * c + 1] = q4.y; hm[4 * c + 2] = q4.z; hm[4 * c + 3] = q4.w;
          }
        }
        drain_kz(1.f, nullptr, 0);  // bias rows of layer m
        if (m == 0) {
          const float om = plan_omega(pl, 0);
          for (int i = 0; i < si; ++i) drain_kz(om * (live ? __ldg(&a.x[b * si + i]) : 0.f), nullptr, 0);
          break;
        }
        if (plan_res(pl, m) != 1) {  // res == 1 (NIF hidden layer): dh_m = T + dh_{m+1}, keep acc
#pragma unroll
          for (int e = 0; e < CH; ++e) acc[e] = 0.f;
        }
        const float om = plan_omega(pl, m);
        float zn[CK];
#pragma unroll
        for (int kl = 0; kl < CK; ++kl) zn[kl] = om * zt_at(kl);
#pragma unroll 1
        for (int c = 0; c < NCHW; ++c) {
          float zc[CK];
#pragma unroll
          for (int kl = 0; kl < CK; ++kl) zc[kl] = zn[kl];
          if (c + 1 < NCHW) {
#pragma unroll
            for (int kl = 0; kl < CK; ++kl) zn[kl] = om * zt_at((c + 1) * CK + kl);
          }
          const uint32_t td = chunk_begin();
#pragma unroll
          for (int kl = 0; kl < CK; ++kl) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int q = 0; q < CH / 32; ++q) {
              float v[32];
              tc_ld32(td + (uint32_t)(kl * NP + half * CH + q * 32), v);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                acc[q * 32 + e] = fmaf(zc[kl], v[e], acc[q * 32 + e]);
                acc[q * 32 + e + 1] = fmaf(zc[kl], v[e + 1], acc[q * 32 + e + 1]);
                s0 = fmaf(v[e], hm[q * 32 + e], s0);
                s1 = fmaf(v[e + 1], hm[q * 32 + e + 1], s1);
              }
            }
            const int kk = c * CK + kl;
            if (kk < K1) dzs[kk * 128] += om * (s0 + s1);
          }
          chunk_end();
        }
      }

      // ---- dz row: the two halves' accumulators added ----
      named_bar_sync(1, 256);
      if (live) {
        const float* d0 = dzs_all + r;
        const float* d1 = dzs_all + (long long)K1 * 128 + r;
        for (int kk = half; kk < K; kk += 2) a.dz[b * K + kk] = d0[kk * 128] + d1[kk * 128];
      }
      named_bar_sync(1, 256);  // the accumulators are cleared by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 256);
}

static int bfb_pick_stages(size_t fixed_bytes) {
  for (int nst = 4; nst >= 2; --nst)
    if (fixed_bytes + (size_t)nst * BF_STAGE_BYTES <= 227 * 1024) return nst;
  return 0;
}

int nif_bf_bwd_data_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                         const float* save, const float* du, float* da, float* dz, cudaStream_t st) {
  if (!nif_plan_uses_bf(pl) || pl.K < 1) return NIF_E_UNSUPPORTED;
  BfBwdArgs a;
  a.nst = bfb_pick_stages(nif_bfb_smem_fixed(pl));
  if (!a.nst) return NIF_E_UNSUPPORTED;
  const size_t smem = nif_bfb_smem_fixed(pl) + (size_t)a.nst * BF_STAGE_BYTES;
  a.B = B;
  a.total_tiles = (B + 127) / 128;
  a.z = z; a.x = x; a.packed = packed; a.save = save; a.du = du; a.da = da; a.dz = dz;
  auto kern = pl.NP == 128 ? nif_bf_bwd_data_kernel<128> : nif_bf_bwd_data_kernel<64>;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = sms;
  if (grid > a.total_tiles) grid = a.total_tiles;
  if (grid < 1) return NIF_OK;
  { NIF_PROF("nif_bf_bwd_data_kernel", st); kern<<<(unsigned)grid, BFB_THREADS, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient of the hidden matrices:  D[(kappa_l, i)][j] += A[(kappa_l, i)][b] * Bm[j][b],  A = zt (x) h_m, Bm = da_m,
// the batch index b is the MMA K dimension; both operands are generated on the fly by row-owning threads, MN-major
// (for a fixed row b the (kappa_l, i) / j index is contiguous):
//   offset(mn, k) = (mn/8) * 1024 + (k/8) * 128 + (k%8) * 16 + (mn%8) * 2        [MN x 64 (b)] bf16 tiles
// One CTA = (hidden matrix, group of 4 "slabs", batch split); a slab is one M = 128 operand: CK = 128/NP latent
// coordinates x NP rows i, N = NP.  64-row sub-tiles, two operand slots (generation of sub-tile t+1 overlaps the MMAs of
// sub-tile t); accumulators stay in TMEM (4 slabs x NP columns) for the whole batch range; the result is written as a
// partial in the layout nif_unpack_grad_kernel sums.  Row data comes straight from the tiled stash / da slots: the 32
// lanes of a warp read 512 contiguous bytes per access.
// ---------------------------------------------------------------------------------------------------
struct BfWgtArgs {
  long long B, rows_per_split;
  int S;
  const float *z, *save, *da;
  float* part;  // [S][H][K+1][NP][NP]
};

#define BFW_THREADS 288
#define BFW_A_BYTES 16384u  // [128 (kappa_l, i) x 64 (b)] bf16, MN-major

__device__ __forceinline__ uint64_t bfw_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;   // LBO: next group of 8 k
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;  // SBO: next group of 8 mn
  d |= (uint64_t)1 << 46;
  return d;
}

template <int NP>
__global__ void __launch_bounds__(BFW_THREADS, 1) nif_bf_bwd_weight_kernel(const Plan pl, const BfWgtArgs a) {
  constexpr int CK = 128 / NP;              // latent coordinates per slab
  constexpr int KQ = 4 * CK;                // latent coordinates per CTA
  constexpr uint32_t B_BYTES = NP * 128u;   // [NP (j) x 64 (b)] bf16
  constexpr uint32_t SLOT_BYTES = 4 * BFW_A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t slot_full[2], slot_empty[2], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = pl.K, K1 = pl.K + 1, H = pl.H;
  const int NG = (K1 + KQ - 1) / KQ;
  const int h = blockIdx.x / NG, pg = blockIdx.x % NG;
  const int s = blockIdx.y;
  const long long r0 = (long long)s * a.rows_per_split;
  long long r1 = r0 + a.rows_per_split;
  if (r1 > a.B) r1 = a.B;
  const long long nsub = r1 > r0 ? (r1 - r0 + 63) / 64 : 0;

  if (tid == 0) {
    mbar_init(&slot_full[0], 128); mbar_init(&slot_full[1], 128);
    mbar_init(&slot_empty[0], 1); mbar_init(&slot_empty[1], 1);
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 8) tc_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 8) {
    // MMA issuer: warp-uniform loop, one elected lane issues
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = bf_idesc(NP, 1, 1);
    for (long long t = 0; t < nsub; ++t) {
      const int sl = (int)(t & 1);
      mbar_wait(&slot_full[sl], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if (tc_elect_one()) {
        const uint32_t base = smem_u32(smem + sl * SLOT_BYTES);
        const uint64_t db = bfw_make_desc(base + 4 * BFW_A_BYTES);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint64_t dA = bfw_make_desc(base + q * BFW_A_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)  // 16 rows (k) per instruction = 2 k-groups = 256 B
            tc_mma_f16(tmem_u + (uint32_t)(q * NP), dA + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), idesc,
                       (t > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&slot_empty[sl]);
      }
      __syncwarp();
    }
    if (tc_elect_one()) tc_commit(&done_bar);
    __syncwarp();
  } else {
    // ---------------- operand generators: slot sl = warp / 4; thread (q, r): slabs 2q, 2q+1 and half q of da; row r ----------------
    const int sl = warp >> 2;
    const int q = (tid >> 6) & 1, r = tid & 63;
    unsigned char* slot = smem + sl * SLOT_BYTES;
    unsigned char* Bt = slot + 4 * BFW_A_BYTES;
    const uint32_t koff = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;  // this row's position along k
    const long long slot_floats = bf_slot_floats(a.B, NP);
    const float* hsrc = a.save + (long long)h * slot_floats;      // h_m, m = h + 1 -> stash slot h
    const float* dsrc = a.da + (long long)(h + 1) * slot_floats;  // da_m
    const int kk0 = KQ * pg + 2 * CK * q;                         // first latent coordinate of this thread
    // fp32 row staging of h_m: in the tiled stash a 64-row sub-tile is one contiguous 64 * NP * 4 byte block; the slot's 128
    // threads fetch it with cp.async (fully coalesced), and the copy of sub-tile t+2 flies while sub-tile t is converted.
    // Row r of the block reads its column quads at (r >> 5) * 32 NP * 4 + quad * 512 + (r & 31) * 16: consecutive rows hit
    // consecutive banks.  (Loading the rows straight from L2 in batches paid one L2 round trip per batch: 5000 cycles per
    // sub-tile against 1024 of MMA time.)
    constexpr uint32_t STAGE_BYTES = 64u * NP * 4u;
    unsigned char* stg = smem + 2 * SLOT_BYTES + sl * STAGE_BYTES;
    const int ts = tid & 127;  // thread index inside the slot
    auto stage_rows = [&](long long t) {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(hsrc + bf_tiled_row(r0 + t * 64, NP));
#pragma unroll
      for (int i = 0; i < (int)(STAGE_BYTES / 2048); ++i) {
        const uint32_t off = (uint32_t)(ts + i * 128) * 16u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(stg + off)), "l"(src + off) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (sl < nsub) stage_rows(sl);
    // this thread's row data that comes straight from global memory -- its latent coordinates and its half of da_m --
    // is fetched one sub-tile ahead
    float ztn[2 * CK];
    float4 dqn[NP / 8];
    auto fetch_row = [&](long long t) {
      const long long b = r0 + t * 64 + r;
      const bool live = t < nsub && b < r1;
#pragma unroll
      for (int kl = 0; kl < 2 * CK; ++kl) {
        const int kk = kk0 + kl;
        ztn[kl] = 0.f;
        if (live) ztn[kl] = (kk < K) ? __ldg(&a.z[b * K + kk]) : (kk == K ? 1.f : 0.f);
      }
      const float* drow = dsrc + bf_tiled_row(live ? b : r0, NP) + (long long)(q * (NP / 8)) * 128;  // quads of columns q*NP/2 ..
#pragma unroll
      for (int c = 0; c < NP / 8; ++c) dqn[c] = live ? ldg4(drow + c * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    fetch_row(sl);
    long long n_mine = 0;
    for (long long t = sl; t < nsub; t += 2, ++n_mine) {
      const long long b = r0 + t * 64 + r;
      const bool live = b < r1;
      float zt[2 * CK];
      float4 dq[NP / 8];
#pragma unroll
      for (int kl = 0; kl < 2 * CK; ++kl) zt[kl] = ztn[kl];
#pragma unroll
      for (int c = 0; c < NP / 8; ++c) dq[c] = dqn[c];
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      named_bar_sync(1 + sl, 128);  // every thread's part of the block has landed
      // wait until the MMAs that read this slot two sub-tiles ago have completed
      mbar_wait(&slot_empty[sl], (uint32_t)((n_mine & 1) ^ 1));
      // A: groups of 8 consecutive i, all of this thread's latent coordinates
#pragma unroll 4
      for (int ig = 0; ig < NP / 8; ++ig) {
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (live) {
          p0 = *reinterpret_cast<const float4*>(stg + (r >> 5) * (32 * NP * 4) + (2 * ig) * 512 + (r & 31) * 16);
          p1 = *reinterpret_cast<const float4*>(stg + (r >> 5) * (32 * NP * 4) + (2 * ig + 1) * 512 + (r & 31) * 16);
        }
#pragma unroll
        for (int kl = 0; kl < 2 * CK; ++kl) {
          const float zk = zt[kl];
          // slab 2q + kl / CK, rows (kl % CK) * NP + 8 ig ..
          const uint32_t mg = (uint32_t)((kl % CK) * (NP / 8) + ig);
          *reinterpret_cast<uint4*>(slot + (uint32_t)(2 * q + kl / CK) * BFW_A_BYTES + mg * 1024u + koff) =
              make_uint4(bf_pack2(zk * p0.x, zk * p0.y), bf_pack2(zk * p0.z, zk * p0.w), bf_pack2(zk * p1.x, zk * p1.y),
                         bf_pack2(zk * p1.z, zk * p1.w));
        }
      }
      named_bar_sync(1 + sl, 128);  // every thread of the slot is done with the staged block
      if (t + 2 < nsub) stage_rows(t + 2);
      fetch_row(t + 2);
      // B: this thread's half of the columns j
#pragma unroll
      for (int u = 0; u < NP / 16; ++u) {
        const float4 p0 = dq[2 * u], p1 = dq[2 * u + 1];
        const uint32_t mg = (uint32_t)(q * (NP / 16) + u);
        *reinterpret_cast<uint4*>(Bt + mg * 1024u + koff) =
            make_uint4(bf_pack2(p0.x, p0.y), bf_pack2(p0.z, p0.w), bf_pack2(p1.x, p1.y), bf_pack2(p1.z, p1.w));
      }
      fence_async_smem();
      mbar_arrive(&slot_full[sl]);
    }
    // ---------------- final epilogue (warps 0-3): TMEM lane = (kappa_l, i) row of the gradient ----------------
    if (warp < 4) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
      const int row = tid;  // 0..127
      const int kl = row / NP, i = row % NP;
      const float scale = plan_omega(pl, h + 1);
      const uint32_t tm = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int qq = 0; qq < 4; ++qq) {
        const int kk = KQ * pg + qq * CK + kl;
        float* dst = a.part + ((((long long)s * H + h) * K1 + kk) * NP + i) * NP;
#pragma unroll 1
        for (int c0 = 0; c0 < NP; c0 += 32) {
          float v[32];
          tc_ld32(tm + (uint32_t)(qq * NP + c0), v);
          tc_wait_ld();
          if (kk < K1) {
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              *reinterpret_cast<float4*>(dst + c0 + e) = make_float4(nsub ? scale * v[e] : 0.f, nsub ? scale * v[e + 1] : 0.f,
                                                                      nsub ? scale * v[e + 2] : 0.f, nsub ? scale * v[e + 3] : 0.f);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 512);
}

int nif_bf_bwd_weight_impl(const Plan& pl, long long B, const float* z, const float* save, const float* da, int S,
                           long long rows_per_split, float* part, cudaStream_t st) {
  if (!nif_plan_uses_bf(pl) || pl.K < 1) return NIF_E_UNSUPPORTED;
  BfWgtArgs a;
  a.B = B; a.rows_per_split = rows_per_split; a.S = S;
  a.z = z; a.save = save; a.da = da; a.part = part;
  const int CK = 128 / pl.NP, KQ = 4 * CK;
  const int NG = (pl.K + 1 + KQ - 1) / KQ;
  const size_t smem = 2 * (size_t)(4 * BFW_A_BYTES + pl.NP * 128u) + 2 * (size_t)(64 * pl.NP * 4);  // slots + row staging
  dim3 grid((unsigned)(pl.H * NG), (unsigned)S);
  if (pl.NP == 128) {
    NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_bf_bwd_weight_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { NIF_PROF("nif_bf_bwd_weight_kernel", st); nif_bf_bwd_weight_kernel<128><<<grid, BFW_THREADS, smem, st>>>(pl, a); }
  } else {
    NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_bf_bwd_weight_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { NIF_PROF("nif_bf_bwd_weight_kernel", st); nif_bf_bwd_weight_kernel<64><<<grid, BFW_THREADS, smem, st>>>(pl, a); }
  }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// ---------------------------------------------------------------------------------------------------
// Thin weight-gradient terms (bias rows of every layer, first matrix, last matrix) on the tensor cores:
//   out[kappa][q] = sum_b zt[b][kappa] F[b][q],   F = on-the-fly features of row b (column space Q of nif_bwd_edge_kernel:
//   [(H+1) NP: da_m[j]] [so: du[c]] [si NP: omega x[i] da_0[j]] [NP so: h_{H+1}[i] du[c]])
// A batch-reduction GEMM with M = kappa (one 128-row tile: K + 1 <= 128), N = a block of 128 feature columns, K = batch.
//   A = zt^T  [128 kappa x 64 b], MN-major, generated by warps 0-3 (two threads per batch row: vector loads of z, 16-byte stores)
//   B = F^T   [128 q x 64 b], MN-major, generated by warps 4-7 / 8-11 (one slot each; two threads per row)
// One CTA = (feature block, batch split); the accumulator stays in TMEM; partials in the layout nif_unpack_grad_kernel sums.
// Feature blocks: m = 0..H (da_m) | i < si (omega x_i da_0) | c < so (du_c h_{H+1}) | one block holding du.
// ---------------------------------------------------------------------------------------------------
struct BfEdgeArgs {
  long long B, rows_per_split;
  int S, Q;
  const float *z, *x, *save, *da, *du;
  float* part;  // [S][K+1][Q]
};
#define BFE_THREADS 416

template <int NP>
__global__ void __launch_bounds__(BFE_THREADS, 1) nif_bf_bwd_edge_kernel(const Plan pl, const BfEdgeArgs a) {
  constexpr uint32_t A_BYTES = 16384u;      // [128 kappa x 64 b] bf16, MN-major
  constexpr uint32_t B_BYTES = NP * 128u;   // [NP q x 64 b] bf16, MN-major
  constexpr uint32_t SLOT_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t slot_full[2], slot_empty[2], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = pl.K, K1 = pl.K + 1, H = pl.H, si = pl.si, so = pl.so;
  const int blk = blockIdx.x, s = blockIdx.y;
  const long long r0 = (long long)s * a.rows_per_split;
  long long r1 = r0 + a.rows_per_split;
  if (r1 > a.B) r1 = a.B;
  const long long nsub = r1 > r0 ? (r1 - r0 + 63) / 64 : 0;
#ifdef NIF_TRACE
  int trace_n = 0;
#define BFE_TRACE(role, tag) do { if (blockIdx.y == 0) TRACE(role, tag); } while (0)
#else
#define BFE_TRACE(role, tag) do {} while (0)
#endif

  if (tid == 0) {
    mbar_init(&slot_full[0], 256); mbar_init(&slot_full[1], 256);
    mbar_init(&slot_empty[0], 1); mbar_init(&slot_empty[1], 1);
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 12) tc_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const long long slot_floats = bf_slot_floats(a.B, NP);

  if (warp == 12) {
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = bf_idesc(NP, 1, 1);  // A and B MN-major
    for (long long t = 0; t < nsub; ++t) {
      const int sl = (int)(t & 1);
      mbar_wait(&slot_full[sl], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if ((tid & 31) == 0) BFE_TRACE(2, (int)t * 8 + 7);
      if (tc_elect_one()) {
        const uint32_t base = smem_u32(smem + sl * SLOT_BYTES);
        const uint64_t dA = bfw_make_desc(base);
        const uint64_t dB = bfw_make_desc(base + A_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc_mma_f16(tmem_u, dA + (uint64_t)(ks * 16), dB + (uint64_t)(ks * 16), idesc, (t > 0 || ks > 0) ? 1u : 0u);
        tc_commit(&slot_empty[sl]);
      }
      __syncwarp();
    }
    if (tc_elect_one()) tc_commit(&done_bar);
    __syncwarp();
  } else if (warp < 4) {
    // ---------------- A generators: thread (hf, r) = batch row r of the sub-tile, kappa groups g = hf, hf + 2, ...; every
    // sub-tile (both slots).  A = zt^T is MN-major like B, so a thread owns a ROW of z: (K + 1) / 8 vector loads and 16-byte
    // stores per row, instead of 64 scalar loads and a long predicate chain per kappa (the device timeline showed this
    // warp group, not the MMAs, setting the pace: 2600 of 3250 cycles per sub-tile).  kappa groups past K are zeroed once.
    const int hf = tid >> 6, r = tid & 63;
    const int K8 = (K1 + 7) / 8;  // kappa groups that hold values (<= 16)
    const uint32_t koff = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
    const bool vec = (K & 3) == 0;  // rows of z are 16-byte aligned
    for (int e = tid; e < 2 * (int)(A_BYTES / 16); e += 128) {
      const int sl = e / (int)(A_BYTES / 16), o = e % (int)(A_BYTES / 16);
      *reinterpret_cast<uint4*>(smem + sl * SLOT_BYTES + o * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    named_bar_sync(1, 128);
    float zn[8][8];
    auto fetch = [&](long long t) {
      const long long b = r0 + t * 64 + r;
      const bool live = t < nsub && b < r1;
      const float* zrow = a.z + (live ? b : r0) * K;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int g = hf + 2 * u;
#pragma unroll
        for (int e = 0; e < 8; ++e) zn[u][e] = 0.f;
        if (g < K8 && live) {
          if (vec && 8 * g + 8 <= K) {
            const float4 p0 = ldg4(zrow + 8 * g), p1 = ldg4(zrow + 8 * g + 4);
            zn[u][0] = p0.x; zn[u][1] = p0.y; zn[u][2] = p0.z; zn[u][3] = p0.w;
            zn[u][4] = p1.x; zn[u][5] = p1.y; zn[u][6] = p1.z; zn[u][7] = p1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int kk = 8 * g + e;
              zn[u][e] = kk < K ? __ldg(zrow + kk) : (kk == K ? 1.f : 0.f);
            }
          }
        }
      }
    };
    fetch(0);
    for (long long t = 0; t < nsub; ++t) {
      const int sl = (int)(t & 1);
      if (tid == 0) BFE_TRACE(0, (int)t * 8 + 0);
      uint4 w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        w[u] = make_uint4(bf_pack2(zn[u][0], zn[u][1]), bf_pack2(zn[u][2], zn[u][3]), bf_pack2(zn[u][4], zn[u][5]),
                          bf_pack2(zn[u][6], zn[u][7]));
      if (tid == 0) BFE_TRACE(0, (int)t * 8 + 1);
      fetch(t + 1);
      mbar_wait(&slot_empty[sl], (uint32_t)(((t >> 1) & 1) ^ 1));
      if (tid == 0) BFE_TRACE(0, (int)t * 8 + 2);
      unsigned char* At = smem + sl * SLOT_BYTES;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (hf + 2 * u < K8) *reinterpret_cast<uint4*>(At + (uint32_t)(hf + 2 * u) * 1024u + koff) = w[u];
      fence_async_smem();
      mbar_arrive(&slot_full[sl]);
      if (tid == 0) BFE_TRACE(0, (int)t * 8 + 3);
    }
    const int kk = tid;  // final epilogue: TMEM lane = kappa
    // ---------------- final epilogue: TMEM lane = kappa ----------------
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    {
      const uint32_t tm = tmem + ((uint32_t)(warp * 32) << 16);
      float* prow = a.part + ((long long)s * K1 + (kk < K1 ? kk : 0)) * a.Q;
      const long long qA = (long long)(H + 1) * NP, qB = qA + so, qC = qB + (long long)si * NP;
#pragma unroll 1
      for (int c0 = 0; c0 < NP; c0 += 32) {
        float v[32];
        tc_ld32(tm + (uint32_t)c0, v);  // warp-collective: every lane executes it, only rows kappa < K + 1 are stored
        tc_wait_ld();
        if (kk < K1) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int j = c0 + e;
            const float val = nsub ? v[e] : 0.f;
            if (blk <= H) prow[(long long)blk * NP + j] = val;                                   // dC_m[j]
            else if (blk < H + 1 + si) prow[qB + (long long)(blk - H - 1) * NP + j] = val;        // dM_0[i][j]
            else if (blk < H + 1 + si + so) prow[qC + (long long)j * so + (blk - H - 1 - si)] = val;  // dM_last[i = j][c]
            else if (j < so) prow[qA + j] = val;                                                 // dC_last[c]
          }
        }
      }
    }
  } else {
    // ---------------- B generators: slot sl; thread (q2, r): row r of the sub-tile, feature columns 64 q2 .. ----------------
    const int sl = (warp - 4) >> 2;
    const int ts = (tid - 128) & 127;
    const int q2 = ts >> 6, r = ts & 63;
    constexpr int NQ = NP / 8;  // quads (of 4 columns) per thread: NP / 2 columns
    const uint32_t koff = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
    // source of this block's features: a tiled slot (da_m / da_0 / h_{H+1}) times an optional per-row factor
    const float* src = nullptr;
    int kind = 0, col = 0;  // 0: plain  1: * omega x[i]  2: * du[c]  3: the du block
    if (blk <= H) { src = a.da + (long long)blk * slot_floats; }
    else if (blk < H + 1 + si) { src = a.da; kind = 1; col = blk - H - 1; }
    else if (blk < H + 1 + si + so) { src = a.save + (long long)H * slot_floats; kind = 2; col = blk - H - 1 - si; }
    else { kind = 3; }
    const float om0 = plan_omega(pl, 0);
    float4 fn[NQ];
    float mul_n = 1.f;
    auto fetch = [&](long long t) {
      const long long b = r0 + t * 64 + r;
      const bool live = t < nsub && b < r1;
      mul_n = 1.f;
      if (kind == 3) {
#pragma unroll
        for (int c = 0; c < NQ; ++c) fn[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && q2 == 0) {
          float d[NIF_MAX_SO];
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c) d[c] = c < so ? __ldg(&a.du[b * so + c]) : 0.f;
          fn[0] = make_float4(d[0], d[1], d[2], d[3]);
          fn[1] = make_float4(d[4], d[5], d[6], d[7]);
        }
        return;
      }
      const float* row = src + bf_tiled_row(live ? b : r0, NP) + (long long)(q2 * NQ) * 128;
#pragma unroll
      for (int c = 0; c < NQ; ++c) fn[c] = live ? ldg4(row + c * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && kind == 1) mul_n = om0 * __ldg(&a.x[b * si + col]);
      if (live && kind == 2) mul_n = __ldg(&a.du[b * so + col]);
    };
    fetch(sl);
    long long n_mine = 0;
    for (long long t = sl; t < nsub; t += 2, ++n_mine) {
      if (tid == 128) BFE_TRACE(1, (int)t * 8 + 4);
      float4 f[NQ];
      const float mul = mul_n;
#pragma unroll
      for (int c = 0; c < NQ; ++c) f[c] = fn[c];
      fetch(t + 2);
      mbar_wait(&slot_empty[sl], (uint32_t)((n_mine & 1) ^ 1));
      if (tid == 128) BFE_TRACE(1, (int)t * 8 + 5);
      unsigned char* Bt = smem + sl * SLOT_BYTES + A_BYTES;
#pragma unroll
      for (int u = 0; u < NQ / 2; ++u) {
        const float4 p0 = f[2 * u], p1 = f[2 * u + 1];
        const uint32_t mg = (uint32_t)(q2 * (NQ / 2) + u);
        *reinterpret_cast<uint4*>(Bt + mg * 1024u + koff) =
            make_uint4(bf_pack2(mul * p0.x, mul * p0.y), bf_pack2(mul * p0.z, mul * p0.w), bf_pack2(mul * p1.x, mul * p1.y),
                       bf_pack2(mul * p1.z, mul * p1.w));
      }
      fence_async_smem();
      mbar_arrive(&slot_full[sl]);
      if (tid == 128) BFE_TRACE(1, (int)t * 8 + 6);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tc_dealloc(tmem, 128);
}

int nif_bf_bwd_edge_impl(const Plan& pl, long long B, const float* z, const float* x, const float* save, const float* da,
                         const float* du, int S, long long rows_per_split, int Q, float* part, cudaStream_t st) {
  if (!nif_plan_uses_bf(pl) || pl.K < 1 || pl.K + 1 > 128) return NIF_E_UNSUPPORTED;
  BfEdgeArgs a;
  a.B = B; a.rows_per_split = rows_per_split; a.S = S; a.Q = Q;
  a.z = z; a.x = x; a.save = save; a.da = da; a.du = du; a.part = part;
  const size_t smem = 2 * (size_t)(16384 + pl.NP * 128);
  dim3 grid((unsigned)(pl.H + 1 + pl.si + pl.so + 1), (unsigned)S);
  if (pl.NP == 128) {
    NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_bf_bwd_edge_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { NIF_PROF("nif_bf_bwd_edge_kernel", st); nif_bf_bwd_edge_kernel<128><<<grid, BFE_THREADS, smem, st>>>(pl, a); }
  } else {
    NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_bf_bwd_edge_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { NIF_PROF("nif_bf_bwd_edge_kernel", st); nif_bf_bwd_edge_kernel<64><<<grid, BFE_THREADS, smem, st>>>(pl, a); }
  }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
