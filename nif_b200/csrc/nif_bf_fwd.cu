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

struct BfFwdArgs {
  long long G, B, tiles_per_group, total_tiles;
  const float *z, *x, *packed;
  int x_shared;
  float *u, *save;
  int nst;
};

#define BFF_THREADS 384

__host__ __device__ inline size_t bff_smem_bytes(int NP, int KZ, int so, int nst) {
  return (size_t)128 * NP * 2 + (size_t)128 * KZ * 2 + (size_t)nst * BF_STAGE_BYTES + (size_t)2 * so * 128 * 4 + 256;
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
  unsigned char* A_tile = smem;
  unsigned char* Z_tile = smem + 128 * NP * 2;
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
      // ---------------- weight-stream producer ----------------
      if (lane == 0) {
        uint32_t s = 0, ph = 0;
        auto put = [&](const float* src, uint32_t bytes) {
          mbar_wait(&b_empty[s], ph ^ 1u);
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
      const uint64_t da = bf_make_desc(smem_u32(A_tile), SBO_A);
      const uint64_t dz = bf_make_desc(smem_u32(Z_tile), sbo_z);
      uint32_t g = 0;          // chunk counter (accumulator stage / phase)
      uint32_t s = 0, ph = 0;  // weight-stream stage and phase
      uint32_t ar = 0;         // a_ready phases consumed
      // one chunk: A operand (zt tile or h tile), K extent in 16-element steps, B row-group stride, N
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
        mbar_wait(&t_full[g & 1u], (g >> 1) & 1u);
        tc_fence_after();
        return tm + (g & 1u) * 128u;
      };
      auto chunk_end = [&]() {
        tc_fence_before();
        mbar_arrive(&t_empty[g & 1u]);
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
#pragma unroll
        for (int c = 0; c < CH / 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = bf_pack2(hcur[8 * c + 2 * e], hcur[8 * c + 2 * e + 1]);
          *reinterpret_cast<uint4*>(A_tile + row_off + (uint32_t)(half * (CH / 8) + c) * 128u) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_async_smem();
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
  if (warp == 8) tc_dealloc(tmem, 256);
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
  return pl.NP == 128 ? dispatch_bff<128>(pl, a, st) : dispatch_bff<64>(pl, a, st);
}
