// Fused forward on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-grade accuracy via FP16x3.
//
// Same mathematics as nif_fwd.cu:
//   pre[b][j] = sum_kappa zt[b][kappa] * ( omega * V[b][kappa][j] + C_m[kappa][j] ),   V = h @ M_m[kappa]
// with the shared-operand product V on the tensor cores:
//   A = h tile [128 rows x 64], per-row power-of-two scale, hi/lo fp16 split, written by the epilogue threads;
//   B = chunk of 2 latent coordinates [128 (kappa_l, j) x 64 (i)], per-slab scale, hi/lo split, streamed from the
//       packed image by cp.async.bulk (32 KB per chunk, 2 stages);
// Layer 0, the bias sums sum_kappa zt[kappa] C_m[kappa][:] and the last layer are tensor-core chunks too (A = the
// zt tile resp. the h tile, small B tiles from the packed image), so the CUDA cores only run the per-row
// latent contraction, the activations and the operand split.
//   D = [128 x 128] fp32 in TMEM, two accumulator stages per tile (the MMAs of chunk c+1 run while chunk c is drained).
// The per-row contraction over kappa (and the un-scaling) runs on the CUDA cores straight out of TMEM: thread r
// owns row r = TMEM lane r.
//
// One CTA per SM, persistent over PAIRS of 128-row tiles.  The two tiles of a pair share every staged weight
// chunk (halves the L2 traffic per row) and ping-pong on the tensor pipe: while the MMAs of one tile run, the
// other tile's epilogue warps drain its accumulators.
//   warps 0-3  epilogue of tile 0        warps 4-7  epilogue of tile 1
//   warp 8     MMA issuer of tile 0 (one elected lane) and TMEM owner (512 columns: 2 tiles x 2 accumulator stages x 128)
//   warp 9     weight-stream producer
//   warp 10    MMA issuer of tile 1;  warp 11 idle (register donor)
#include "nif_tc.cuh"

NIF_TRACE_READER(nif_debug_read_trace)

struct TcFwdArgs {
  long long B, total_pairs;
  const float *z, *x, *packed;
  float *u, *save;
  const float* psave;  // tangent mode: the primal stash (h_{m+1}, d_m) of the same rows
};

#define TCF_THREADS 384  // 8 epilogue warps + MMA warp + producer warp + 2 idle warps (register donors)
#define TCF_STAGES 2
#define TCF_STAGE_BYTES 32768u  // [hi | lo] x 16 KB (small chunks use a prefix)

// Chunk schedule of one tile pair (identical for producer, MMA issuer and epilogue):
//   Z0(i'), i' = 0..si      A = zt tile, B = X0[i']     N = 64      layer 0:  pre0 += xt[i'] * D
//   for m = 1..H:  ZC(m)    A = zt tile, B = XC[m]      N = 64      bias sum of layer m (initialises acc)
//                  M(m,c)   A = h_m tile, B = TCF[m-1][c] N = 128   c = 0..NCH-1
//   L(q), q < NLC           A = h_{H+1} tile, B = XL[q] N = LPC*KZ  last layer
__host__ __device__ inline size_t tcf_smem_bytes(int KP, int KZ, int si) {
  return 4 * (size_t)TC_TILE_BYTES + 2 * 2 * 128 * (size_t)KZ * 2 + TCF_STAGES * (size_t)TCF_STAGE_BYTES +
         2 * (size_t)KP * 128 * 4 + 2 * (size_t)si * 128 * 4 + 256;
}

// SINE: the activation is sine (SIREN variants), inlined; otherwise the out-of-line activation switch is called.
// TAN: forward-mode tangent of a sine network without residual layers along one ShapeNet-input direction (a.x = xdot):
// the same chunk schedule with the constant terms dropped (xt' = [omega xdot, 0], no bias sums), and the epilogue
//   h'_{m+1} = d_m * pre'_m,   e_m = act''(pre_m) * pre'_m = -h_{m+1} * pre'_m
// with d_m = cos(pre_m) and h_{m+1} = sin(pre_m) read from the primal stash; a.save receives h' (slots 0..H) and e
// (slots H+1..2H+1), a.u the tangent of the output (JacobianLayer inside the loss, nif/layers/gradient.py:207-231).
template <bool SAVE, bool SINE, bool TAN = false>
__global__ void __launch_bounds__(TCF_THREADS, 1) nif_tc_fwd_kernel(const Plan pl, const TcFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A_all = smem;                                  // tile t: hi at t*32K, lo at t*32K + 16K
  unsigned char* Z_all = smem + 4 * TC_TILE_BYTES;              // zt operand, tile t: hi at t*2*zbytes, lo at + zbytes
  const uint32_t zbytes = 128u * (uint32_t)pl.KZ * 2u;          // one of hi / lo
  const uint32_t sbo_z = (uint32_t)(pl.KZ / 8) * 128u;          // 8-row group stride of a KZ-wide K-major tile
  unsigned char* Bst = Z_all + 4 * zbytes;                      // [TCF_STAGES][hi | lo]
  float* zs_all = reinterpret_cast<float*>(Bst + TCF_STAGES * TCF_STAGE_BYTES);  // [2][KP][128]
  float* xs_all = zs_all + 2 * pl.KP * 128;                     // [2][si][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs_all + 2 * pl.si * 128);
  uint64_t* b_full = bars;                    // [TCF_STAGES]
  uint64_t* b_empty = bars + TCF_STAGES;      // [TCF_STAGES]
  uint64_t* t_full = bars + 2 * TCF_STAGES;   // [2][2]  accumulator stage s of tile t ready
  uint64_t* t_empty = t_full + 4;             // [2][2]  accumulator stage s of tile t drained
  uint64_t* a_ready = t_empty + 4;            // [2]  operand tile of tile t written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = pl.K, K1 = pl.K + 1, KP = pl.KP, NCH = pl.NCH, H = pl.H, n = pl.n, si = pl.si, so = pl.so;
  const int KZ = pl.KZ, NLC = pl.NLC, LPC = pl.LPC;
#ifdef NIF_TRACE
  int trace_n = 0;
#endif
  const uint32_t small_bytes = 2u * 64u * (uint32_t)KZ * 2u;        // X0 / XC chunk [hi | lo]
  const uint32_t last_bytes = 2u * (uint32_t)(LPC * KZ) * 64u * 2u;  // XL chunk [hi | lo]

  if (tid == 0) {
    for (int i = 0; i < TCF_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 2);  // both MMA issuers
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 128);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&a_ready[i], 128);
    mbar_fence_init();
  }
  if (warp == 8) tc_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  long long my_pairs = 0;
  if ((long long)blockIdx.x < a.total_pairs) my_pairs = (a.total_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp >= 8) {
  tc_reg_dec<56>();  // MMA / producer / idle warps donate registers: 128 x (168 - 56) freed = 256 x (224 - 168) claimed below
  if (warp == 9) {
    // ---------------- weight-stream producer ----------------
    if (lane == 0) {
      long long g = 0;
      const float* tcf = a.packed + pl.off_TCF;
      const float* tcx = a.packed + pl.off_TCX;
      auto put = [&](const float* src, uint32_t bytes) {
        const int s = (int)(g % TCF_STAGES);
        mbar_wait(&b_empty[s], (uint32_t)(((g / TCF_STAGES) & 1) ^ 1));
        TRACE(3, g);
        mbar_expect_tx(&b_full[s], bytes);
        bulk_g2s(Bst + s * TCF_STAGE_BYTES, src, bytes, &b_full[s]);
        ++g;
      };
      for (long long t = 0; t < my_pairs; ++t) {
        for (int i = 0; i <= si; ++i) put(tcx + (long long)i * plan_x0_floats(pl), small_bytes);
        for (int h = 0; h < H; ++h) {
          put(tcx + (long long)(si + 1 + h) * plan_x0_floats(pl), small_bytes);
          for (int c = 0; c < NCH; ++c) put(tcf + ((long long)h * NCH + c) * NIF_TC_CHUNK_FLOATS, TCF_STAGE_BYTES);
        }
        for (int q = 0; q < NLC; ++q)
          put(tcx + (long long)(si + 1 + H) * plan_x0_floats(pl) + (long long)q * plan_xl_floats(pl), last_bytes);
      }
    }
  } else if (warp == 8 || warp == 10) {
    // ---------------- MMA issuers: warp 8 feeds tile 0, warp 10 feeds tile 1 ----------------
    // (one issuer per tile: while one is between chunks -- barrier waits, descriptors, commits -- the other keeps
    // the tensor pipe fed; both read the same staged weight chunk and each commits to its b_empty, count 2)
    {  // the whole warp runs the loop (uniform control flow and operands); one elected lane issues -- see tc_elect_one
      const int t = __shfl_sync(0xffffffffu, warp == 8 ? 0 : 1, 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint64_t da_hi = tc_make_desc(smem_u32(A_all + t * 2 * TC_TILE_BYTES), TC_SBO);
      const uint64_t da_lo = tc_make_desc(smem_u32(A_all + t * 2 * TC_TILE_BYTES + TC_TILE_BYTES), TC_SBO);
      const uint64_t dz_hi = tc_make_desc(smem_u32(Z_all + t * 2 * zbytes), sbo_z);
      const uint64_t dz_lo = tc_make_desc(smem_u32(Z_all + t * 2 * zbytes + zbytes), sbo_z);
      long long g = 0;   // chunk counter (stage / accumulator phases)
      long long ar = 0;  // a_ready phases consumed
      // one chunk: A operand kind (0 = zt tile, 1 = h tile), K extent, B geometry, N
      auto chunk = [&](int a_kind, bool wait_a, int ksteps, uint32_t b_half_bytes, uint32_t b_sbo, int N) {
        const int s = (int)(g % TCF_STAGES);
        const int as = (int)(g & 1);  // accumulator stage
        if (wait_a) { mbar_wait(&a_ready[t], (uint32_t)(ar & 1)); ++ar; }
        mbar_wait(&t_empty[2 * t + as], (uint32_t)(((g >> 1) & 1) ^ 1));
        mbar_wait(&b_full[s], (uint32_t)((g / TCF_STAGES) & 1));
        if (t == 0 && lane == 0) TRACE(2, g * 8 + 1 + t);
        tc_fence_after();
        const uint64_t db_hi = tc_make_desc(smem_u32(Bst + s * TCF_STAGE_BYTES), b_sbo);
        const uint64_t db_lo = tc_make_desc(smem_u32(Bst + s * TCF_STAGE_BYTES + b_half_bytes), b_sbo);
        const uint32_t d = tmem_u + (uint32_t)t * 256u + (uint32_t)as * 128u;
        if (tc_elect_one()) {
          tc_mma_split(d, a_kind ? da_hi : dz_hi, a_kind ? da_lo : dz_lo, db_hi, db_lo, tc_idesc_f16(N), ksteps);
          tc_commit(&t_full[2 * t + as]);
          tc_commit(&b_empty[s]);
        }
        __syncwarp();
        if (t == 0 && lane == 0) TRACE(2, g * 8 + 3 + t);
        ++g;
      };
      for (long long p = 0; p < my_pairs; ++p) {
        for (int i = 0; i <= si; ++i) chunk(0, i == 0, KZ / 16, small_bytes / 2, sbo_z, 64);
        for (int h = 0; h < H; ++h) {
          chunk(0, false, KZ / 16, small_bytes / 2, sbo_z, 64);
          for (int c = 0; c < NCH; ++c) chunk(1, c == 0, 4, TC_TILE_BYTES, TC_SBO, 128);
        }
        for (int q = 0; q < NLC; ++q) chunk(1, q == 0, 4, last_bytes / 2, TC_SBO, LPC * KZ);
      }
    }
  }
  } else {
    tc_reg_inc<224>();
    // ---------------- epilogue warps: thread r <-> row r of tile wg <-> TMEM lane r ----------------
    const int wg = warp >> 2;  // tile of the pair
    const int r = tid & 127;
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)wg * 256u;
    unsigned char* A_hi = A_all + wg * 2 * TC_TILE_BYTES;
    unsigned char* A_lo = A_hi + TC_TILE_BYTES;
    unsigned char* Z_hi = Z_all + wg * 2 * zbytes;
    unsigned char* Z_lo = Z_hi + zbytes;
    float* zs = zs_all + wg * KP * 128;
    float* xs = xs_all + wg * si * 128;
    const float* C_all = a.packed + pl.off_C;
    const float* invB = a.packed + pl.off_TCS;
    const float* invX = a.packed + pl.off_TCS2;
    const long long slot_floats = nif_tiled_rows(a.B) * 64;  // one stash slot
    const uint32_t row_off = (uint32_t)(r >> 3) * TC_SBO + (uint32_t)(r & 7) * 16u;
    long long g = 0;  // chunk counter (accumulator phases), advances exactly like the MMA issuer's
    for (long long p = 0; p < my_pairs; ++p) {
      const long long row0 = ((blockIdx.x + p * gridDim.x) * 2 + wg) * 128;
      const long long b = row0 + r;
      const bool live = b < a.B;
      named_bar_sync(1 + wg, 128);  // this tile's threads are done with the previous zs / xs
      // every thread stages its own row of z (independent vector loads, conflict-free transposed stores)
      if ((K & 3) == 0) {
        for (int k4 = 0; k4 < K; k4 += 4) {
          const float4 q4 = live ? ldg4(a.z + b * K + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
          zs[k4 * 128 + r] = q4.x; zs[(k4 + 1) * 128 + r] = q4.y; zs[(k4 + 2) * 128 + r] = q4.z; zs[(k4 + 3) * 128 + r] = q4.w;
        }
      } else {
        for (int kk = 0; kk < K; ++kk) zs[kk * 128 + r] = live ? __ldg(&a.z[b * K + kk]) : 0.f;
      }
      zs[K * 128 + r] = 1.f;
      for (int kk = K1; kk < KP; ++kk) zs[kk * 128 + r] = 0.f;
      for (int i = 0; i < si; ++i) xs[i * 128 + r] = live ? __ldg(&a.x[b * si + i]) : 0.f;
      named_bar_sync(1 + wg, 128);

      // ---- zt operand tile (used by layer 0 and by every bias-sum chunk) ----
      float inv_z;
      {
        float zmax = 0.f;
        for (int kk = 0; kk < K1; ++kk) zmax = fmaxf(zmax, fabsf(zs[kk * 128 + r]));
        float sc_z;
        tc_row_scale(zmax, sc_z, inv_z);
        // the MMAs that read the previous pair's zt tile completed before this thread finished that pair
        tc_store_row_split_fn(Z_hi, Z_lo, r, sbo_z, KZ / 8, sc_z, [&](int kk) { return kk < K1 ? zs[kk * 128 + r] : 0.f; });
        fence_async_smem();
        mbar_arrive(&a_ready[wg]);
      }

      float hcur[64];     // output of the layer being finished (input of the next one)
      float inv_a = 1.f;  // inverse of the power-of-two scale of this row's h operand tile

      // acc[j] (+)= coef * (D1 + D2)[j] for the 64 columns of a small chunk
      auto drain64 = [&](uint32_t td, float (&acc)[64], float coef, bool init) {
        float v1[32], v2[32];
        tc_ld32(td, v1);
        tc_ld32(td + 32u, v2);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          acc[e] = init ? coef * v1[e] : fmaf(coef, v1[e], acc[e]);
          acc[32 + e] = init ? coef * v2[e] : fmaf(coef, v2[e], acc[32 + e]);
        }
      };
      // accumulator stage g & 1 of this tile: wait until its MMAs have completed; returns its TMEM address
      auto chunk_begin = [&]() -> uint32_t {
        if (r == 0) TRACE(wg, g * 4 + 0);
        mbar_wait(&t_full[2 * wg + (int)(g & 1)], (uint32_t)((g >> 1) & 1));
        if (r == 0) TRACE(wg, g * 4 + 1);
        tc_fence_after();
        return tm + (uint32_t)(g & 1) * 128u;
      };
      auto chunk_end = [&]() {
        tc_fence_before();
        mbar_arrive(&t_empty[2 * wg + (int)(g & 1)]);
        if (r == 0) TRACE(wg, g * 4 + 2);
        ++g;
      };

      // activation / residual / stash for layer m given pre-activations in `pre`; result in hcur.
      auto finish_layer = [&](int m, float (&pre)[64]) {
        const int res = plan_res(pl, m);  // 0 or 1 on this path (res-blocks use the CUDA-core kernel)
        float* sh = a.save + (long long)m * slot_floats + nif_tiled_row(b);            // h_{m+1}, tiled layout
        float* sd = a.save + (long long)(H + 1 + m) * slot_floats + nif_tiled_row(b);  // d_m
        if (TAN) {  // (padded columns: the primal stash holds d = h = 0 there)
          const float* ph = a.psave + (long long)m * slot_floats + nif_tiled_row(b);
          const float* pd = a.psave + (long long)(H + 1 + m) * slot_floats + nif_tiled_row(b);
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 dv = live ? ldg4(pd + c * 128) : z4;
            const float4 hv = live ? ldg4(ph + c * 128) : z4;
            const float4 ev = make_float4(-hv.x * pre[4 * c], -hv.y * pre[4 * c + 1], -hv.z * pre[4 * c + 2], -hv.w * pre[4 * c + 3]);
            hcur[4 * c] = dv.x * pre[4 * c]; hcur[4 * c + 1] = dv.y * pre[4 * c + 1];
            hcur[4 * c + 2] = dv.z * pre[4 * c + 2]; hcur[4 * c + 3] = dv.w * pre[4 * c + 3];
            if (live) {
              *reinterpret_cast<float4*>(sh + c * 128) = make_float4(hcur[4 * c], hcur[4 * c + 1], hcur[4 * c + 2], hcur[4 * c + 3]);
              *reinterpret_cast<float4*>(sd + c * 128) = ev;
            }
          }
          return;
        }
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
          float dch[8], f8[8], d8[8];
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
            const int j = 8 * c + e;
            float f = f8[e], d = d8[e];
            float o = f;
            if (res == 1) o += hold[e];
            if (j >= n) { o = 0.f; d = 0.f; }
            hcur[j] = o;
            dch[e] = d;
          }
          if (SAVE && live) {  // column quads 2c, 2c+1: a warp writes 512 contiguous bytes per store
            *reinterpret_cast<float4*>(sh + (2 * c) * 128) = make_float4(hcur[8 * c], hcur[8 * c + 1], hcur[8 * c + 2], hcur[8 * c + 3]);
            *reinterpret_cast<float4*>(sh + (2 * c + 1) * 128) = make_float4(hcur[8 * c + 4], hcur[8 * c + 5], hcur[8 * c + 6], hcur[8 * c + 7]);
            *reinterpret_cast<float4*>(sd + (2 * c) * 128) = make_float4(dch[0], dch[1], dch[2], dch[3]);
            *reinterpret_cast<float4*>(sd + (2 * c + 1) * 128) = make_float4(dch[4], dch[5], dch[6], dch[7]);
          }
        }
      };
      // h operand tile for the next tensor-core layer.  Every MMA that read the previous contents has
      // completed: this thread observed the t_full of the last chunk that used it.
      auto publish_h = [&]() {
        float amax = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) amax = fmaxf(amax, fabsf(hcur[j]));
        float sc_a;
        tc_row_scale(amax, sc_a, inv_a);
        tc_store_row_split(A_hi, A_lo, r, hcur, sc_a);
        fence_async_smem();
        mbar_arrive(&a_ready[wg]);
      };

      // ---- layers 0 .. H (one loop, so the activation epilogue exists once in the instruction stream) ----
#pragma unroll 1
      for (int m = 0; m <= H; ++m) {
        float acc[64];
        if (m == 0) {
          // layer 0:  pre0[j] = sum_i' xt[i'] * (zt @ X0[i'])[j],  xt = [omega x, 1]
          const float om = plan_omega(pl, 0);
          for (int i = 0; i <= si; ++i) {
            const float coef = inv_z * __ldg(&invX[i]) * (i < si ? om * xs[i * 128 + r] : (TAN ? 0.f : 1.f));
            const uint32_t td = chunk_begin();
            drain64(td, acc, coef, i == 0);
            chunk_end();
          }
        } else {
          publish_h();
          {
            const float coef = TAN ? 0.f : inv_z * __ldg(&invX[si + m]);
            const uint32_t td = chunk_begin();  // bias sum: acc[j] = sum_kappa zt[kappa] C_m[kappa][j]
            drain64(td, acc, coef, true);
            chunk_end();
          }
          const float om_inv = plan_omega(pl, m) * inv_a;
          const float* invBm = invB + (m - 1) * KP;
          // per-chunk row coefficients, fetched one chunk ahead (their latency hides behind the accumulator wait)
          float zo2[2] = {zs[r] * om_inv * __ldg(&invBm[0]), zs[128 + r] * om_inv * __ldg(&invBm[1])};
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const float zc2[2] = {zo2[0], zo2[1]};
            if (c + 1 < NCH) {
              zo2[0] = zs[(2 * c + 2) * 128 + r] * om_inv * __ldg(&invBm[2 * c + 2]);
              zo2[1] = zs[(2 * c + 3) * 128 + r] * om_inv * __ldg(&invBm[2 * c + 3]);
            }
            const uint32_t td = chunk_begin();
#pragma unroll
            for (int kl = 0; kl < 2; ++kl) {  // the 64 columns of one latent coordinate, then its contraction
              const float zo = zc2[kl];
              float v1[32], v2[32];
              tc_ld32(td + (uint32_t)(kl * 64), v1);
              tc_ld32(td + (uint32_t)(kl * 64 + 32), v2);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                acc[e] = fmaf(zo, v1[e], acc[e]);
                acc[32 + e] = fmaf(zo, v2[e], acc[32 + e]);
              }
            }
            chunk_end();
          }
        }
        finish_layer(m, acc);
        if (r == 0) TRACE(wg, g * 4 + 3);
      }

      // ---- last layer:  y[c] = sum_kappa zt[kappa] * ( (h @ ML[kappa])[c] + CL[kappa][c] ) ----
      publish_h();
      {
        const float* CL = C_all + (long long)(H + 1) * K1 * 64;
        for (int q = 0; q < NLC; ++q) {
          const float sL = inv_a * __ldg(&invX[si + 1 + H + q]);
          const uint32_t td = chunk_begin();
          for (int cl = 0; cl < LPC; ++cl) {
            const int c = LPC * q + cl;
            float y = 0.f;
            for (int k0 = 0; k0 < KZ; k0 += 16) {
              float v1[16];
              tc_ld16(td + (uint32_t)(cl * KZ + k0), v1);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int kk = k0 + e;
                if (kk < K1 && c < so)
                  y = fmaf(zs[kk * 128 + r], TAN ? sL * v1[e] : fmaf(sL, v1[e], __ldg(&CL[(long long)kk * 64 + c])), y);
              }
            }
            if (live && c < so) a.u[b * so + c] = y;
          }
          chunk_end();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 512);
}

template <bool SAVE, bool SINE, bool TAN = false>
static int launch_tcf(const Plan& pl, const TcFwdArgs& a, cudaStream_t st) {
  const size_t smem = tcf_smem_bytes(pl.KP, pl.KZ, pl.si);
  auto kern = nif_tc_fwd_kernel<SAVE, SINE, TAN>;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = sms;
  if (grid > a.total_pairs) grid = a.total_pairs;
  if (grid < 1) return NIF_OK;
  { NIF_PROF(TAN ? "nif_tc_fwd_kernel<tangent>" : "nif_tc_fwd_kernel", st); kern<<<(unsigned)grid, TCF_THREADS, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

size_t nif_tcb_smem_bytes(int K);
// One static predicate for forward and reverse (the stash layout depends on it): shapes the tensor-core kernels cover.
bool nif_plan_uses_tc(const Plan& pl) {
  if (!pl.tc || pl.NP != 64 || pl.H < 1 || pl.variant == NIF_VARIANT_SIREN_RES || pl.K < 1) return false;
  if (tcf_smem_bytes(pl.KP, pl.KZ, pl.si) > 227 * 1024 || pl.LPC * pl.KZ > 128) return false;
  return nif_tcb_smem_bytes(pl.K) <= 227 * 1024;
}

// returns NIF_E_UNSUPPORTED (without setting an error) when the shape does not fit this kernel, so that the
// dispatcher can use the CUDA-core kernel instead
int nif_tc_forward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed, float* u,
                        float* save, cudaStream_t st) {
  if (!nif_plan_uses_tc(pl)) return NIF_E_UNSUPPORTED;
  TcFwdArgs a;
  a.B = B;
  a.total_pairs = (B + 255) / 256;
  a.z = z; a.x = x; a.packed = packed; a.u = u; a.save = save; a.psave = nullptr;
  if (pl.act == NIF_ACT_SINE) return save ? launch_tcf<true, true>(pl, a, st) : launch_tcf<false, true>(pl, a, st);
  return save ? launch_tcf<true, false>(pl, a, st) : launch_tcf<false, false>(pl, a, st);
}

// Shapes whose Sobolev step (forward tangents with a stash + reverse-over-forward) runs on the tensor-core kernels: the
// tangent epilogue derives act'' from the stashed sine, and the residual bookkeeping is not carried through it.
bool nif_plan_tc_sobolev(const Plan& pl) {
  if (!nif_plan_uses_tc(pl) || pl.act != NIF_ACT_SINE) return false;
  for (int m = 0; m <= pl.H; ++m)
    if (plan_res(pl, m) != 0 || plan_alpha(pl, m) != 1.0f) return false;
  return true;
}

// Tangent of the network along one ShapeNet-input direction xdot [B][si] (the latent code does not move), given the
// primal stash of nif_tc_forward_impl (tiled): udot [B][so], tsave = (H+1) slots h' + (H+1) slots e (tiled).
int nif_tc_forward_tangent_impl(const Plan& pl, long long B, const float* z, const float* xdot, const float* packed,
                                const float* psave, float* udot, float* tsave, cudaStream_t st) {
  if (!nif_plan_tc_sobolev(pl)) return NIF_E_UNSUPPORTED;
  TcFwdArgs a;
  a.B = B;
  a.total_pairs = (B + 255) / 256;
  a.z = z; a.x = xdot; a.packed = packed; a.u = udot; a.save = tsave; a.psave = psave;
  return launch_tcf<true, true, true>(pl, a, st);
}
