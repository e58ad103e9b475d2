// Fused forward on the tensor cores with bf16 operands (dtype_compute = 1), padded widths NP = 64 and 128.
//
// Same mathematics and chunk structure as nif_tc_fwd.cu:
//   pre[b][j] = sum_kappa zt[b][kappa] * ( omega * V[b][kappa][j] + C_m[kappa][j] ),   V = h @ M_m[kappa]
// with V on tcgen05 (A = the h tile [128 rows x NP] written by the epilogue threads, B = a chunk of CK = 128/NP latent
// coordinates [128 (kappa_l, j) x NP (i)] streamed from the W image by cp.async.bulk), the contraction over kappa on the
// CUDA cores straight out of TMEM, and layer 0 / the bias sums / the last layer as small tensor-core chunks.
//
// One 128-row tile per CTA, 256 epilogue threads: thread (r, half) owns row r and NP/2 columns (nif_bf.cuh).  The
// latent code is read from global memory as it is needed (one value per chunk and thread, L1-resident sectors), so
// shared memory holds only operands:  A tile | zt tile | NST weight stages | the last layer's half sums.
//   warps 0-3  rows 0..127, column half 0      warps 4-7  rows 0..127, column half 1
//   warp 8     MMA issuer (one elected lane) and TMEM owner (2 accumulator stages x 128 columns)
//   warp 9     weight-stream producer;  warps 10-11 idle (register donors)
// Grouped launches (G > 1: K == 0, one W image per group, optionally one coordinate grid shared by all groups) serve the
// latent-sweep inference of SURVEY 8 a7 (model_x_to_u_given_w, nif/model.py:435-464, 956-986, in factored form).
#include "nif_bf.cuh"

NIF_TRACE_READER(nif_debug_read_trace_bfg)

struct BfFwdArgs {
  long long G, B, tiles_per_group, total_tiles;
  const float *z, *x, *packed;
  int x_shared;
  float *u, *save;
  int nst;
};

#define BFF_THREADS 384

__host__ __device__ inline size_t bff_smem_bytes(int NP, int KZ, int so, int nst) {
  (void)NP;
  return (size_t)128 * KZ * 2 + (size_t)nst * BF_STAGE_BYTES + (size_t)2 * so * 128 * 4 + 256;
}

// Chunk schedule of one tile (identical for producer, MMA issuer and epilogue):
//   Z0(i'), i' = 0..si     A = zt tile, B = X0[i']          N = NP     layer 0:  pre0 += xt[i'] * D
//   for m = 1..H:  ZC(m)   A = zt tile, B = XC[m]           N = NP     bias sum of layer m (initialises acc)
//                  M(m,c)  A = h_m tile, B = WF[m-1][c]     N = 128    c = 0..NCHW-1
//   L(c), c < so           A = h_{H+1} tile, B = XL[c]      N = KZ     last layer
template <int NP, bool SAVE, bool SINE>
__global__ void __launch_bounds__(BFF_THREADS, 1) nif_bf_fwd_kernel(const Plan pl, const BfFwdArgs a) {
  constexpr int CH = NP / 2;    // columns per epilogue thread
  constexpr int CK = 128 / NP;  // latent coordinates per main chunk
  constexpr uint32_t SBO_A = (NP / 8) * 128u;
  constexpr uint32_t MAIN_BYTES = 128u * NP * 2u;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Z_tile = smem;  // (the h operand tile lives in tensor memory)
  const uint32_t zbytes = 128u * (uint32_t)pl.KZ * 2u;
  const uint32_t sbo_z = (uint32_t)(pl.KZ / 8) * 128u;
  unsigned char* Bst = Z_tile + zbytes;
  float* ys = reinterpret_cast<float*>(Bst + a.nst * BF_STAGE_BYTES);  // [2][so][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ys + 2 * pl.so * 128);
  uint64_t* b_full = bars;          // [8]
  uint64_t* b_empty = bars + 8;     // [8]
  uint64_t* t_full = bars + 16;     // [2]
  uint64_t* t_empty = bars + 18;    // [2]
  uint64_t* a_ready = bars + 20;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = pl.K, K1 = pl.K + 1, H = pl.H, n = pl.n, si = pl.si, so = pl.so, KZ = pl.KZ;
  const int NCHW = bf_nchw(pl);
  const uint32_t nst = (uint32_t)a.nst;
  const uint32_t small_bytes = (uint32_t)NP * (uint32_t)KZ * 2u;
#ifdef NIF_TRACE
  int trace_n = 0;  // tools/bff_trace.py: roles 0 = epilogue thread 0, 2 = MMA issuer, 3 = weight-stream producer
#endif

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 256); }
    mbar_init(&a_ready[0], 256);
    mbar_fence_init();
  }
  // columns [0, 256): two accumulator stages; [256, 256 + NP / 2): the h operand tile (A of the main and last-layer chunks)
  if (warp == 8) tc_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  long long my_tiles = 0;
  if ((long long)blockIdx.x < a.total_tiles) my_tiles = (a.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp >= 8) {
    tc_reg_dec<56>();
    if (warp == 9) {
      // ---------------- weight-stream producer ----------------
      if (lane == 0) {
        uint32_t s = 0, ph = 0;
        auto put = [&](const float* src, uint32_t bytes) {
          mbar_wait(&b_empty[s], ph ^ 1u);
          TRACE(3, 0);
          mbar_expect_tx(&b_full[s], bytes);
          bulk_g2s(Bst + s * BF_STAGE_BYTES, src, bytes, &b_full[s]);
          if (++s == nst) { s = 0; ph ^= 1u; }
        };
        for (long long t = 0; t < my_tiles; ++t) {
          const long long tile = blockIdx.x + t * gridDim.x;
          const long long g = tile / a.tiles_per_group;
          const float* pk = a.packed + g * pl.packed_floats;
          const float* wf = pk + pl.off_WF;
          const float* wx = pk + pl.off_WX;
          for (int i = 0; i <= si; ++i) put(wx + (long long)bf_t_x0(pl, i) * bf_small_floats(pl), small_bytes);
          for (int m = 1; m <= H; ++m) {
            put(wx + (long long)bf_t_xc(pl, m) * bf_small_floats(pl), small_bytes);
            for (int c = 0; c < NCHW; ++c) put(wf + ((long long)(m - 1) * NCHW + c) * bf_chunk_floats(pl), MAIN_BYTES);
          }
          for (int c = 0; c < so; ++c) put(wx + (long long)bf_t_xl(pl, c) * bf_small_floats(pl), small_bytes);
        }
      }
    } else if (warp == 8) {
      // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues (see tc_elect_one) ----------
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint64_t dz = bf_make_desc(smem_u32(Z_tile), sbo_z);
      uint32_t g = 0;          // chunk counter (accumulator stage / phase)
      uint32_t s = 0, ph = 0;  // weight-stream stage and phase
      uint32_t ar = 0;         // a_ready phases consumed
      // one chunk: A operand (zt tile or h tile), K extent in 16-element steps, B row-group stride, N
      auto chunk = [&](bool a_main, bool wait_a, int ksteps, uint32_t b_sbo, int N) {
        const uint32_t as = g & 1u;
        if (wait_a) { mbar_wait(&a_ready[0], ar & 1u); ++ar; }
        if (lane == 0) TRACE(2, g * 8 + 0);
        mbar_wait(&t_empty[as], ((g >> 1) & 1u) ^ 1u);
        if (lane == 0) TRACE(2, g * 8 + 1);
        mbar_wait(&b_full[s], ph);
        if (lane == 0) TRACE(2, g * 8 + 2);
        tc_fence_after();
        const uint64_t db = bf_make_desc(smem_u32(Bst + s * BF_STAGE_BYTES), b_sbo);
        const uint32_t d = tmem_u + as * 128u;
        const uint32_t idesc = bf_idesc(N);
        if (tc_elect_one()) {
          if (a_main) {  // A = the h tile in tensor memory: its 32 KB are not re-read from shared memory for every chunk
            for (int ks = 0; ks < ksteps; ++ks)
              tc_mma_f16_ts(d, tmem_u + 256u + (uint32_t)(ks * 8), db + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : 0u);
          } else {
            for (int ks = 0; ks < ksteps; ++ks)
              tc_mma_f16(d, dz + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : 0u);
          }
          tc_commit(&t_full[as]);
          tc_commit(&b_empty[s]);
        }
        __syncwarp();
        ++g;
        if (++s == nst) { s = 0; ph ^= 1u; }
      };
      for (long long t = 0; t < my_tiles; ++t) {
        for (int i = 0; i <= si; ++i) chunk(false, i == 0, KZ / 16, sbo_z, NP);
        for (int m = 1; m <= H; ++m) {
          chunk(false, false, KZ / 16, sbo_z, NP);
          for (int c = 0; c < NCHW; ++c) chunk(true, c == 0, NP / 16, SBO_A, 128);
        }
        for (int c = 0; c < so; ++c) chunk(true, c == 0, NP / 16, SBO_A, KZ);
      }
    }
  } else {
    tc_reg_inc<224>();
    // ---------------- epilogue warps: thread (r, half) <-> row r (TMEM lane r), columns half*CH .. ----------------
    const int half = warp >> 2;
    const int r = tid & 127;
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t row_off = (uint32_t)(r >> 3) * SBO_A + (uint32_t)(r & 7) * 16u;
    const uint32_t zrow_off = (uint32_t)(r >> 3) * sbo_z + (uint32_t)(r & 7) * 16u;
    const long long slot_floats = bf_slot_floats(a.B, NP);
    uint32_t g = 0;
    for (long long t = 0; t < my_tiles; ++t) {
      const long long tile = blockIdx.x + t * gridDim.x;
      const long long grp = tile / a.tiles_per_group;
      const long long b = (tile - grp * a.tiles_per_group) * 128 + r;  // row inside the group
      const bool live = b < a.B;
      const float* zrow = a.z + (grp * a.B + (live ? b : 0)) * K;      // (never dereferenced when K == 0)
      const float* xrow = a.x + ((a.x_shared ? 0 : grp * a.B) + (live ? b : 0)) * si;
      const float* C_all = a.packed + grp * pl.packed_floats + pl.off_C;
      auto zt_at = [&](int kk) -> float { return kk < K ? (live ? __ldg(zrow + kk) : 0.f) : (kk == K ? 1.f : 0.f); };

      // ---- zt operand tile (layer 0 and every bias-sum chunk): this thread writes its half of the k-groups ----
      {
        const int ng = KZ / 8;  // 16-byte groups of 8 latent coordinates
        for (int c = half; c < ng; c += 2) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = bf_pack2(zt_at(8 * c + 2 * e), zt_at(8 * c + 2 * e + 1));
          *reinterpret_cast<uint4*>(Z_tile + zrow_off + c * 128) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_async_smem();
        mbar_arrive(&a_ready[0]);
      }

      float hcur[CH];  // output of the layer being finished (this thread's columns)

      auto chunk_begin = [&]() -> uint32_t {
        if (tid == 0) TRACE(0, g * 8 + 3);
        mbar_wait(&t_full[g & 1u], (g >> 1) & 1u);
        tc_fence_after();
        if (tid == 0) TRACE(0, g * 8 + 4);
        return tm + (g & 1u) * 128u;
      };
      auto chunk_end = [&]() {
        tc_fence_before();
        mbar_arrive(&t_empty[g & 1u]);
        if (tid == 0) TRACE(0, g * 8 + 5);
        ++g;
      };
      // acc[e] (+)= coef * D[col0 + e], e < CH
      auto drain = [&](uint32_t td, uint32_t col0, float (&acc)[CH], float coef, bool init) {
#pragma unroll
        for (int q = 0; q < CH / 32; ++q) {
          float v[32];
          tc_ld32(td + col0 + (uint32_t)(q * 32), v);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[q * 32 + e] = init ? coef * v[e] : fmaf(coef, v[e], acc[q * 32 + e]);
        }
      };
      // the h operand tile of the next tensor-core layer (every MMA that read the previous contents has completed:
      // this thread observed the t_full of the last chunk that used it)
      auto publish_h = [&]() {
        uint32_t w[CH / 2];
#pragma unroll
        for (int e = 0; e < CH / 2; ++e) w[e] = bf_pack2(hcur[2 * e], hcur[2 * e + 1]);
        if constexpr (CH / 2 == 32) tc_st32(tm + 256u + (uint32_t)(half * (CH / 2)), w);
        else tc_st16(tm + 256u + (uint32_t)(half * (CH / 2)), w);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&a_ready[0]);
      };
      // activation / residual / stash of layer m; `pre` holds this thread's pre-activations, result in hcur.
      // hold: the layer's input (residual variants), fp32
      auto finish_layer = [&](int m, float (&pre)[CH]) {
        const int res = plan_res(pl, m);  // 0 or 1 on this path
        float* sh = a.save + (long long)m * slot_floats + bf_tiled_row(b, NP) + (long long)(half * (CH / 4)) * 128;
        float* sd = a.save + (long long)(H + 1 + m) * slot_floats + bf_tiled_row(b, NP) + (long long)(half * (CH / 4)) * 128;
#pragma unroll
        for (int c = 0; c < CH / 8; ++c) {
          float f8[8], d8[8];
          if (SINE) {
            nif_sincos_fold8(&pre[8 * c]);
#pragma unroll
            for (int e = 0; e < 8; ++e) nif_sincosf_core(pre[8 * c + e], f8[e], d8[e]);
          } else {
#pragma unroll
            for (int e4 = 0; e4 < 8; e4 += 4) {
              const float v4[4] = {pre[8 * c + e4], pre[8 * c + e4 + 1], pre[8 * c + e4 + 2], pre[8 * c + e4 + 3]};
              float f4[4], d4[4];
              act_fd4(pl.act, v4, f4, d4);
#pragma unroll
              for (int e = 0; e < 4; ++e) { f8[e4 + e] = f4[e]; d8[e4 + e] = d4[e]; }
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = half * CH + 8 * c + e;
            float o = f8[e];
            if (res == 1) o += hcur[8 * c + e];  // NIF hidden layer: out = in + act(pre)
            if (j >= n) { o = 0.f; d8[e] = 0.f; }
            hcur[8 * c + e] = o;
          }
          if (SAVE && live) {
            *reinterpret_cast<float4*>(sh + (2 * c) * 128) = make_float4(hcur[8 * c], hcur[8 * c + 1], hcur[8 * c + 2], hcur[8 * c + 3]);
            *reinterpret_cast<float4*>(sh + (2 * c + 1) * 128) = make_float4(hcur[8 * c + 4], hcur[8 * c + 5], hcur[8 * c + 6], hcur[8 * c + 7]);
            *reinterpret_cast<float4*>(sd + (2 * c) * 128) = make_float4(d8[0], d8[1], d8[2], d8[3]);
            *reinterpret_cast<float4*>(sd + (2 * c + 1) * 128) = make_float4(d8[4], d8[5], d8[6], d8[7]);
          }
        }
      };

      // ---- layers 0 .. H ----
#pragma unroll 1
      for (int m = 0; m <= H; ++m) {
        float acc[CH];
        if (m == 0) {
          const float om = plan_omega(pl, 0);
          for (int i = 0; i <= si; ++i) {
            const float coef = i < si ? om * (live ? __ldg(xrow + i) : 0.f) : 1.f;
            const uint32_t td = chunk_begin();
            drain(td, (uint32_t)(half * CH), acc, coef, i == 0);
            chunk_end();
          }
        } else {
          publish_h();
          {
            const uint32_t td = chunk_begin();  // bias sum: acc[j] = sum_kappa zt[kappa] C_m[kappa][j]
            drain(td, (uint32_t)(half * CH), acc, 1.f, true);
            chunk_end();
          }
          const float om = plan_omega(pl, m);
          // per-chunk row coefficients, fetched one chunk ahead
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
            for (int kl = 0; kl < CK; ++kl) drain(td, (uint32_t)(kl * NP + half * CH), acc, zc[kl], false);
            chunk_end();
          }
        }
        finish_layer(m, acc);
      }

      // ---- last layer:  y[c] = sum_kappa zt[kappa] * ( (h @ ML[kappa])[c] + CL[kappa][c] ); the two halves split kappa ----
      publish_h();
      {
        const float* CL = C_all + (long long)(H + 1) * K1 * pl.NP;
        const int kh = KZ / 2, k0 = half * kh;  // this thread's latent coordinates (KZ / 2 is a multiple of 8)
        for (int c = 0; c < so; ++c) {
          const uint32_t td = chunk_begin();
          float y = 0.f;
          for (int q = 0; q < kh; q += 8) {
            float v[8];
            tc_ld8(td + (uint32_t)(k0 + q), v);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int kk = k0 + q + e;
              if (kk < K1) y = fmaf(zt_at(kk), v[e] + __ldg(&CL[(long long)kk * pl.NP + c]), y);
            }
          }
          chunk_end();
          ys[(half * so + c) * 128 + r] = y;
        }
        named_bar_sync(1, 256);
        if (half == 0 && live) {
          for (int c = 0; c < so; ++c) a.u[(grp * a.B + b) * so + c] = ys[c * 128 + r] + ys[(so + c) * 128 + r];
        }
        named_bar_sync(1, 256);  // ys is rewritten by the next tile
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 512);
}

static int bf_pick_stages(size_t fixed_bytes) {
  for (int nst = 4; nst >= 2; --nst)
    if (fixed_bytes + (size_t)nst * BF_STAGE_BYTES <= 227 * 1024) return nst;
  return 0;
}

size_t nif_bfb_smem_fixed(const Plan& pl);
// One static predicate for forward and reverse (the stash layout depends on it): shapes the bf16 tensor-core kernels cover.
bool nif_plan_uses_bf(const Plan& pl) {
  if (!pl.bf || (pl.NP != 64 && pl.NP != 128) || pl.H < 1 || pl.variant == NIF_VARIANT_SIREN_RES) return false;
  if ((size_t)pl.NP * pl.KZ * 2 > BF_STAGE_BYTES || pl.KZ > 128) return false;  // small tiles fit one stage; N = KZ <= 128
  if (!bf_pick_stages(bff_smem_bytes(pl.NP, pl.KZ, pl.so, 0))) return false;
  return bf_pick_stages(nif_bfb_smem_fixed(pl)) != 0;
}

template <int NP, bool SAVE, bool SINE>
static int launch_bff(const Plan& pl, BfFwdArgs& a, cudaStream_t st) {
  a.nst = bf_pick_stages(bff_smem_bytes(NP, pl.KZ, pl.so, 0));
  const size_t smem = bff_smem_bytes(NP, pl.KZ, pl.so, a.nst);
  auto kern = nif_bf_fwd_kernel<NP, SAVE, SINE>;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = sms;
  if (grid > a.total_tiles) grid = a.total_tiles;
  if (grid < 1) return NIF_OK;
  { NIF_PROF("nif_bf_fwd_kernel", st); kern<<<(unsigned)grid, BFF_THREADS, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

template <int NP>
static int dispatch_bff(const Plan& pl, BfFwdArgs& a, cudaStream_t st) {
  const bool save = a.save != nullptr;
  if (pl.act == NIF_ACT_SINE) return save ? launch_bff<NP, true, true>(pl, a, st) : launch_bff<NP, false, true>(pl, a, st);
  return save ? launch_bff<NP, true, false>(pl, a, st) : launch_bff<NP, false, false>(pl, a, st);
}

template <int NP>
static int launch_bfg(const Plan& pl, BfFwdArgs& a, cudaStream_t st);

// returns NIF_E_UNSUPPORTED (without setting an error) when the shape does not fit these kernels
int nif_bf_forward_impl(const Plan& pl, long long G, long long B, const float* z, const float* x, int x_shared,
                        const float* packed, float* u, float* save, cudaStream_t st) {
  if (!nif_plan_uses_bf(pl)) return NIF_E_UNSUPPORTED;
  if (G > 1 && pl.K > 0) return NIF_E_UNSUPPORTED;
  if (save && pl.K < 1) return NIF_E_UNSUPPORTED;  // the reverse kernels of this path need a latent code
  BfFwdArgs a;
  a.G = G; a.B = B;
  a.tiles_per_group = (B + 127) / 128;
  a.total_tiles = a.tiles_per_group * G;
  a.z = z; a.x = x; a.packed = packed; a.x_shared = x_shared; a.u = u; a.save = save; a.nst = 0;
  if (pl.K == 0 && !save) {  // explicit weight vectors (grouped sweeps): the two-tile inference kernel
    const int rc = pl.NP == 128 ? launch_bfg<128>(pl, a, st) : launch_bfg<64>(pl, a, st);
    if (rc != NIF_E_UNSUPPORTED) return rc;
  }
  return pl.NP == 128 ? dispatch_bff<128>(pl, a, st) : dispatch_bff<64>(pl, a, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// K == 0 inference (grouped latent sweeps, SURVEY 8 a7 / BASELINE config 5): every group is a plain 128-wide SIREN MLP
// with its own weights, so a layer is ONE tensor-core chunk and the activation epilogue, not the tensor pipe, sets the
// pace.  Two tiles per CTA share every staged weight chunk and ping-pong: while the 256 epilogue threads finish a layer
// of one tile, the MMAs of the other tile run.  Layer 0 (si inputs) and every bias come from the fp32 sections of the
// image on the CUDA cores; sin is the hardware approximation (sin.approx: |error| < 1e-6 for |x| < 100, far inside a
// bf16 operand's rounding), which leaves one MUFU op, one bias add and half a pack per element.
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t bfg_smem_bytes(int NP, int nst) {
  return (size_t)2 * 128 * NP * 2 + (size_t)nst * BF_STAGE_BYTES + 256;
}

#define BFG_THREADS 640  // 16 epilogue warps (four threads per row: 32 lanes x NP/4 columns each), issuer, producer, 2 idle

template <int NP, bool SINE>
__global__ void __launch_bounds__(BFG_THREADS, 1) nif_bf_group_fwd_kernel(const Plan pl, const BfFwdArgs a) {
  constexpr int CQ = NP / 4;  // columns per epilogue thread
  constexpr uint32_t SBO_A = (NP / 8) * 128u;
  constexpr uint32_t TILE_BYTES = 128u * NP * 2u;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A_all = smem;  // tile t at t * TILE_BYTES
  unsigned char* Bst = smem + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bst + a.nst * BF_STAGE_BYTES);
  uint64_t* b_full = bars;         // [8]
  uint64_t* b_empty = bars + 8;    // [8]
  uint64_t* t_full = bars + 16;    // [2]  accumulator of tile t ready
  uint64_t* t_free = bars + 18;    // [2]  accumulator of tile t drained
  uint64_t* a_ready = bars + 20;   // [2]  operand tile of tile t written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = pl.H, n = pl.n, si = pl.si, so = pl.so, KZ = pl.KZ;
  const int NCHW = bf_nchw(pl);
  const uint32_t nst = (uint32_t)a.nst;
  const long long pairs_per_group = (a.tiles_per_group + 1) / 2;
  const long long total_pairs = pairs_per_group * a.G;
#ifdef NIF_TRACE
  int trace_n = 0;
  // roles: 0 = epilogue thread 0 (warp 0, column quarter 0), 1 = epilogue thread 416 (warp 13, quarter 3), 2 = MMA issuer
#define BFG_TRACE(tag) do { if (tid == 0) TRACE(0, tag); else if (tid == 416) TRACE(1, tag); } while (0)
#else
#define BFG_TRACE(tag) do {} while (0)
#endif

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_free[i], 512); mbar_init(&a_ready[i], 512); }
    mbar_fence_init();
  }
  if (warp == 16) tc_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  long long my_pairs = 0;
  if ((long long)blockIdx.x < total_pairs) my_pairs = (total_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp >= 16) {
    tc_reg_dec<40>();
    if (warp == 17) {
      if (lane == 0) {  // weight stream: per pair H main chunks (chunk 0 of every layer) and so last-layer tiles
        uint32_t s = 0, ph = 0;
        auto put = [&](const float* src, uint32_t bytes) {
          mbar_wait(&b_empty[s], ph ^ 1u);
          mbar_expect_tx(&b_full[s], bytes);
          bulk_g2s(Bst + s * BF_STAGE_BYTES, src, bytes, &b_full[s]);
          if (++s == nst) { s = 0; ph ^= 1u; }
        };
        for (long long p = 0; p < my_pairs; ++p) {
          const long long g = (blockIdx.x + p * gridDim.x) / pairs_per_group;
          const float* pk = a.packed + g * pl.packed_floats;
          for (int m = 1; m <= H; ++m) put(pk + pl.off_WF + (long long)(m - 1) * NCHW * bf_chunk_floats(pl), NP * NP * 2u);
          for (int c = 0; c < so; ++c)
            put(pk + pl.off_WX + (long long)bf_t_xl(pl, c) * bf_small_floats(pl), (uint32_t)NP * (uint32_t)KZ * 2u);
        }
      }
    } else if (warp == 16) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      uint32_t s = 0, ph = 0, aph = 0, fph0 = 0, fph1 = 0;
      for (long long p = 0; p < my_pairs; ++p) {
        for (int st = 0; st < H + so; ++st) {
          const bool last = st >= H;
          mbar_wait(&b_full[s], ph);
          for (int t = 0; t < 2; ++t) {
            if (!last || st == H) { mbar_wait(&a_ready[t], aph); }
            uint32_t& fph = t ? fph1 : fph0;
            mbar_wait(&t_free[t], fph ^ 1u);
            fph ^= 1u;
            tc_fence_after();
            const uint64_t dA = bf_make_desc(smem_u32(A_all + t * TILE_BYTES), SBO_A);
            const uint64_t dB = bf_make_desc(smem_u32(Bst + s * BF_STAGE_BYTES), SBO_A);
            if (lane == 0) TRACE(2, (st * 2 + t) * 8 + 0);
            if (tc_elect_one()) {
              const uint32_t idesc = bf_idesc(last ? KZ : NP);
#pragma unroll
              for (int ks = 0; ks < NP / 16; ++ks)
                tc_mma_f16(tmem_u + (uint32_t)t * 128u, dA + (uint64_t)(ks * 16), dB + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : 0u);
              tc_commit(&t_full[t]);
              if (t == 1) tc_commit(&b_empty[s]);
            }
            __syncwarp();
          }
          if (!last || st == H) aph ^= 1u;
          if (++s == nst) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    tc_reg_inc<104>();  // (the pool only holds what the other warp group released: 128 x (96 - 40) >= 512 x (104 - 96))
    const int qt = warp >> 2;  // column quarter
    const int r = tid & 127;
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t row_off = (uint32_t)(r >> 3) * SBO_A + (uint32_t)(r & 7) * 16u;
    uint32_t fph[2] = {0, 0};  // t_full phases
    auto act = [&](float v) -> float {
      if (SINE) return __sinf(v);
      return act_f(pl.act, v);
    };
    // h values (this thread's CQ columns of tile t) -> operand tile, bf16
    auto publish = [&](int t, const float (&h)[CQ]) {
#pragma unroll
      for (int c = 0; c < CQ / 8; ++c) {
        *reinterpret_cast<uint4*>(A_all + t * TILE_BYTES + row_off + (uint32_t)(qt * (CQ / 8) + c) * 128u) =
            make_uint4(bf_pack2(h[8 * c], h[8 * c + 1]), bf_pack2(h[8 * c + 2], h[8 * c + 3]), bf_pack2(h[8 * c + 4], h[8 * c + 5]),
                       bf_pack2(h[8 * c + 6], h[8 * c + 7]));
      }
      fence_async_smem();
      mbar_arrive(&a_ready[t]);
    };
    for (long long p = 0; p < my_pairs; ++p) {
      const long long pair = blockIdx.x + p * gridDim.x;
      const long long grp = pair / pairs_per_group;
      const long long pr = pair - grp * pairs_per_group;
      const float* pk = a.packed + grp * pl.packed_floats;
      const float* C_all = pk + pl.off_C;   // [Lm][1][NP] fp32 biases
      const float* M0 = pk + pl.off_M0;     // [1][si][NP]
      const float om0 = plan_omega(pl, 0);
      long long brow[2];
      bool live[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        brow[t] = (pr * 2 + t) * 128 + r;
        live[t] = brow[t] < a.B;
      }
      // ---- layer 0 on the CUDA cores ----
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float xv[NIF_MAX_SI];
        const float* xrow = a.x + ((a.x_shared ? 0 : grp * a.B) + (live[t] ? brow[t] : 0)) * si;
#pragma unroll
        for (int i = 0; i < NIF_MAX_SI; ++i) xv[i] = (i < si && live[t]) ? om0 * __ldg(xrow + i) : 0.f;
        float h[CQ];
#pragma unroll
        for (int c = 0; c < CQ / 4; ++c) {
          const int j0 = qt * CQ + 4 * c;
          const float4 b4 = ldg4(C_all + j0);
          float pre[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < NIF_MAX_SI; ++i)
            if (i < si) {
              const float4 w4 = ldg4(M0 + (long long)i * pl.NP + j0);
              pre[0] = fmaf(xv[i], w4.x, pre[0]); pre[1] = fmaf(xv[i], w4.y, pre[1]);
              pre[2] = fmaf(xv[i], w4.z, pre[2]); pre[3] = fmaf(xv[i], w4.w, pre[3]);
            }
#pragma unroll
          for (int e = 0; e < 4; ++e) h[4 * c + e] = (SINE || j0 + e < n) ? act(pre[e]) : 0.f;  // (padding: sin(0) = 0)
        }
        publish(t, h);
      }
      // ---- hidden layers ----
#pragma unroll 1
      for (int m = 1; m <= H; ++m) {
        const float* bm = C_all + (long long)m * pl.NP + qt * CQ;
        const int res = plan_res(pl, m);
        const float om = plan_omega(pl, m);
        // this layer's biases (the same for both tiles), fetched before the accumulator wait
        float bb[CQ];
#pragma unroll
        for (int c = 0; c < CQ / 4; ++c) {
          const float4 b4 = ldg4(bm + 4 * c);
          bb[4 * c] = b4.x; bb[4 * c + 1] = b4.y; bb[4 * c + 2] = b4.z; bb[4 * c + 3] = b4.w;
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          BFG_TRACE((m * 2 + t) * 8 + 0);
          mbar_wait(&t_full[t], fph[t]);
          fph[t] ^= 1u;
          tc_fence_after();
          BFG_TRACE((m * 2 + t) * 8 + 1);
          float h[CQ];
          if (CQ == 32) {
            float v[32];
            tc_ld32(tm + (uint32_t)(t * 128 + qt * CQ), v);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) h[e] = fmaf(om, v[e], bb[e]);
          } else {
            float v[16];
            tc_ld16(tm + (uint32_t)(t * 128 + qt * CQ), v);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) h[e] = fmaf(om, v[e], bb[e]);
          }
          tc_fence_before();
          mbar_arrive(&t_free[t]);
          BFG_TRACE((m * 2 + t) * 8 + 2);
          // padded columns have zero weights and biases: sin(0) = 0 without a mask (2.5 of 7 instructions per element)
#pragma unroll
          for (int e = 0; e < CQ; ++e) h[e] = (SINE || qt * CQ + e < n) ? act(h[e]) : 0.f;
          if (res == 1) {  // NIF hidden layer: out = in + act(pre); the input is this tile's operand (bf16)
#pragma unroll
            for (int c = 0; c < CQ / 8; ++c) {
              const uint4 q4 = *reinterpret_cast<const uint4*>(A_all + t * TILE_BYTES + row_off + (uint32_t)(qt * (CQ / 8) + c) * 128u);
              const uint32_t w[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
                h[8 * c + 2 * e] += f2.x;
                h[8 * c + 2 * e + 1] += f2.y;
              }
            }
          }
#ifdef NIF_TRACE
          asm volatile("" ::"f"(h[0]), "f"(h[CQ - 1]), "f"(h[CQ / 2]) : "memory");
#endif
          BFG_TRACE((m * 2 + t) * 8 + 3);
          publish(t, h);
          BFG_TRACE((m * 2 + t) * 8 + 4);
        }
      }
      // ---- last layer: y[c] = h . ML[:, c] + CL[c]  (column kappa = 0 of the N = KZ chunk) ----
      const float* CL = C_all + (long long)(H + 1) * pl.NP;
      for (int c = 0; c < so; ++c) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&t_full[t], fph[t]);
          fph[t] ^= 1u;
          tc_fence_after();
          float v[8];
          tc_ld8(tm + (uint32_t)(t * 128), v);
          tc_wait_ld();
          tc_fence_before();
          mbar_arrive(&t_free[t]);
          if (qt == 0 && live[t]) a.u[(grp * a.B + brow[t]) * so + c] = v[0] + __ldg(CL + c);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tc_dealloc(tmem, 256);
}

template <int NP>
static int launch_bfg(const Plan& pl, BfFwdArgs& a, cudaStream_t st) {
  int nst = 0;
  for (int s = 4; s >= 2 && !nst; --s)
    if (bfg_smem_bytes(NP, s) <= 227 * 1024) nst = s;
  if (!nst) return NIF_E_UNSUPPORTED;
  a.nst = nst;
  const size_t smem = bfg_smem_bytes(NP, nst);
  auto kern = pl.act == NIF_ACT_SINE ? nif_bf_group_fwd_kernel<NP, true> : nif_bf_group_fwd_kernel<NP, false>;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long total_pairs = (a.tiles_per_group + 1) / 2 * a.G;
  long long grid = sms;
  if (grid > total_pairs) grid = total_pairs;
  if (grid < 1) return NIF_OK;
  { NIF_PROF("nif_bf_group_fwd_kernel", st); kern<<<(unsigned)grid, BFG_THREADS, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
