// Reverse data pass on the tensor cores (tcgen05 + TMEM, FP16x3) -- the mirror image of nif_tc_fwd.cu.
//
// Per 128-row tile, hidden matrices m = H .. 1:
//   T[b][kappa][i] = sum_j da_m[b][j] M_m[kappa][i][j]        tcgen05 GEMM: A = da_m tile (row-scaled, hi/lo split),
//                                                             B = TCB chunk [128 (kappa_l, i) x 64 (j)]
//   dh_m[b][i]   = omega * sum_kappa zt[b][kappa] T[b][kappa][i]            epilogue, thread = row
//   dz[b][kappa] += omega * sum_i T[b][kappa][i] h_m[b][i]                  epilogue, thread = row (no shuffles)
//   da_{m-1}     = dh_m * act'(pre_{m-1})                                   -> global stash + next operand tile
// The thin dz terms are tensor-core chunks of the same kernel, over operand tiles that are in shared memory anyway:
//   bias rows     dz[kappa] += (da_m @ BCt[m])[kappa]                       A = da_m tile,     N = KZ     (m = H .. 0)
//   first matrix  dz[kappa] += omega x[i] (da_0 @ B0t[i])[kappa]            A = da_0 tile,     N = KZ
//   last matrix   dz[kappa] += du[c] ((h_{H+1} @ XL[q])[c_l, kappa] + CL[kappa][c])   A = h_{H+1} tile, N = LPC KZ
// (round 1 ran them in a separate CUDA-core kernel, nif_dz_edge_kernel: 111 us of the 1.5 ms step for 1 % of its flops).
// Also records max|da_m| per layer and max|zt|, max|h_m| (atomicMax on the float bits) for the operand scales
// of the weight-gradient GEMM.
#include "nif_tc.cuh"

NIF_TRACE_READER(nif_debug_read_trace_bwd)

struct TcBwdArgs {
  long long B, total_pairs;
  const float *z, *x, *packed, *save, *du;
  float* da;        // [(H+1)][B][64]
  float* dz;        // [B][K]
  unsigned* maxes;  // [0..H] max|da_m| bits, [H+1] max|zt| bits, [H+2 .. 2H+2] max|h_m| bits (m = 1..H+1)
  int nst;          // weight-stream stages (3 where shared memory allows, else 2)
  // Reverse-over-forward passes (EXT instantiation; the same hooks as BwdArgs of the CUDA-core kernel, nif_bwd.cu):
  const float* h_stash;  // the h_m slots this pass pairs with its da_m (save, or the tangent activations h'_m); d_m stays save's
  const float* e_stash;  // [H+1] slots e_m; with ext_out: ext_out[m] = dh_{m+1} * e_m (tangent-adjoint pass)
  float* ext_out;
  const float* ext_add;  // [H+1] slots added to da_m (the primal-adjoint pass consumes what the tangent pass wrote)
  int dz_accumulate;     // dz += instead of =
};

// slots of the maxima buffer beyond [0, 2H + 2): operand-scale bounds for nif_tc_bwd_edge_kernel
#define NIF_MAX_X 192    // + i : max |x[:, i]|
#define NIF_MAX_DU 200   // + c : max |du[:, c]|
#define NIF_MAX_HL 208   //       max |h_{H+1}|

#define TCB_THREADS 384  // 8 epilogue warps + MMA warp + producer warp + 2 idle warps (register donors)
#define TCB_MAX_STAGES 3
#define TCB_STAGE_BYTES 32768u

// operand tiles of both row tiles | weight-stream stages | zt and dz rows of both tiles ([K+1][128] each) | barriers.
// A third stage decouples the two tiles of a pair: with two, the tile that runs ahead holds a stage until the other has
// used it and then waits a full L2 round trip for the chunk after next (traced: 2350 cycles per chunk against 1536 of MMA).
size_t nif_tcb_smem_bytes(int K);
__host__ __device__ inline size_t tcb_smem_bytes(int K, int nst) {
  return 4 * (size_t)TC_TILE_BYTES + (size_t)nst * TCB_STAGE_BYTES + 4 * (size_t)(K + 1) * 128 * 4 + 256;
}

size_t nif_tcb_smem_bytes(int K) { return tcb_smem_bytes(K, 2); }

__device__ __forceinline__ void warp_atomic_max(unsigned* dst, float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) atomicMax(dst, __float_as_uint(v));  // v >= 0: uint order == float order
}

// MMA issuer of tile T (warp 8: tile 0, warp 10: tile 1; see nif_tc_fwd.cu).  The whole warp runs the loop and one
// elected lane issues.  T is a template constant and every counter is a 32-bit value derived from kernel parameters, so
// descriptors, barrier addresses and phases all live in uniform registers: the per-chunk instruction stream of this
// warp is a few dozen instructions.  It has to be: the warp shares its scheduler with two epilogue warps whose dense
// FMA streams keep the issue slot, and every extra instruction here delays the next chunk's MMAs (traced: the issuer
// used to reach the top of its loop only when the epilogue of the previous chunk went to sleep).
template <int T>
__device__ __forceinline__ void tcb_issue(const Plan& pl, const TcBwdArgs& a, unsigned char* smem, uint64_t* bars,
                                          uint32_t tmem, long long my_pairs) {
  const uint32_t nst = (uint32_t)a.nst;
  unsigned char* A_all = smem;
  unsigned char* Bst = smem + 4 * TC_TILE_BYTES;
  uint64_t* b_full = bars;
  uint64_t* b_empty = bars + TCB_MAX_STAGES;
  uint64_t* t_full = bars + 2 * TCB_MAX_STAGES;
  uint64_t* t_empty = t_full + 4;
  uint64_t* a_ready = t_empty + 4;
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0) + (uint32_t)T * 256u;
  const uint32_t idesc_main = tc_idesc_f16(128);
  const uint32_t idesc_kz = tc_idesc_f16(pl.KZ);
  const uint32_t idesc_last = tc_idesc_f16(pl.LPC * pl.KZ);
  const uint32_t lo_kz = (uint32_t)(pl.KZ * 128) >> 4;             // lo half of a [KZ x 64] tile, in descriptor units
  const uint32_t lo_last = (uint32_t)(pl.LPC * pl.KZ * 128) >> 4;  // lo half of an XL tile
  const uint64_t da_hi = tc_make_desc(smem_u32(A_all + T * 2 * TC_TILE_BYTES));
  const uint64_t da_lo = tc_make_desc(smem_u32(A_all + T * 2 * TC_TILE_BYTES + TC_TILE_BYTES));
  const uint64_t db0 = tc_make_desc(smem_u32(Bst));
  const uint32_t H = (uint32_t)pl.H, NCH = (uint32_t)pl.NCH, NLC = (uint32_t)pl.NLC, si = (uint32_t)pl.si;
  uint32_t g = 0, ar = 0;    // chunk counter; operand-tile publications consumed
  uint32_t s = 0, sph = 0;   // weight-stream stage and its phase
  const int lane = threadIdx.x & 31;
  (void)lane;
#ifdef NIF_TRACE
  int trace_n = 0;
#endif
  // one chunk: K = 64 (4 k-steps, three products); lo_off: distance of the B tile's lo half
  auto chunk = [&](bool wait_a, uint32_t idesc, uint32_t lo_off) {
    const uint32_t as = g & 1u;  // accumulator stage: the MMAs of chunk g+1 run while chunk g is drained
    if (wait_a) { mbar_wait(&a_ready[T], ar & 1u); ++ar; }
    mbar_wait(&t_empty[2 * T + as], ((g >> 1) & 1u) ^ 1u);
    if (T == 0 && lane == 0) TRACE(2, g * 8 + 2);
    mbar_wait(&b_full[s], sph);
    if (T == 0 && lane == 0) TRACE(2, g * 8 + 1);
    tc_fence_after();
    const uint64_t boff = (uint64_t)(s * (TCB_STAGE_BYTES >> 4));
    if (tc_elect_one()) {
      tc_mma_split_k64(tmem_u + as * 128u, da_hi, da_lo, db0 + boff, db0 + boff + lo_off, idesc);
      tc_commit(&t_full[2 * T + as]);
      tc_commit(&b_empty[s]);
    }
    __syncwarp();
    if (T == 0 && lane == 0) TRACE(2, g * 8 + 3);
    ++g;
    if (++s == nst) { s = 0; sph ^= 1u; }
  };
  for (long long p = 0; p < my_pairs; ++p) {
    for (uint32_t q = 0; q < NLC; ++q) chunk(q == 0, idesc_last, lo_last);   // last matrix: A = h_{H+1} tile
    for (uint32_t h = 0; h < H; ++h) {
      chunk(true, idesc_kz, lo_kz);                                           // bias rows: A = da_m tile (just published)
      for (uint32_t c = 0; c < NCH; ++c) chunk(false, idesc_main, (uint32_t)(TC_TILE_BYTES >> 4));
    }
    chunk(true, idesc_kz, lo_kz);                                             // bias rows of layer 0: A = da_0 tile
    for (uint32_t i = 0; i < si; ++i) chunk(false, idesc_kz, lo_kz);          // first matrix
  }
}

// MODE: 0 = the plain reverse pass; reverse-over-forward passes (compile-time, so that the extra loads of the layer loop
// are straight-line code the compiler can batch): 1 = tangent adjoint (writes ext_out, bias rows dropped), 2 = the same,
// accumulating into ext_out (second and later directions), 3 = primal adjoint (adds ext_add)
template <int MODE>
__global__ void __launch_bounds__(TCB_THREADS, 1) nif_tc_bwd_data_kernel(const Plan pl, const TcBwdArgs a) {
  constexpr bool EXT = MODE != 0;
  constexpr bool NO_BIAS = MODE == 1 || MODE == 2;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A_all = smem;
  unsigned char* Bst = smem + 4 * TC_TILE_BYTES;
  const int nst = a.nst;
  float* zs_all = reinterpret_cast<float*>(Bst + nst * TCB_STAGE_BYTES);  // [2][K+1][128]
  float* dzs_all = zs_all + 2 * (pl.K + 1) * 128;                         // [2][K+1][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dzs_all + 2 * (pl.K + 1) * 128);
  uint64_t* b_full = bars;
  uint64_t* b_empty = bars + TCB_MAX_STAGES;
  uint64_t* t_full = bars + 2 * TCB_MAX_STAGES;  // [2][2]  accumulator stage s of tile t ready
  uint64_t* t_empty = t_full + 4;            // [2][2]  accumulator stage s of tile t drained
  uint64_t* a_ready = t_empty + 4;           // [2]     operand tile of tile t written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = pl.K, K1 = pl.K + 1, KP = pl.KP, NCH = pl.NCH, H = pl.H, n = pl.n, so = pl.so, si = pl.si;
  const int KZ = pl.KZ, NLC = pl.NLC, LPC = pl.LPC;
#ifdef NIF_TRACE
  int trace_n = 0;
#endif

  if (tid == 0) {
    for (int i = 0; i < nst; ++i) {
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
    if (lane == 0) {  // weight-stream producer: the chunk schedule of tcb_issue
      long long g = 0;
      const float* src0 = a.packed + pl.off_TCB;
      const float* tcx = a.packed + pl.off_TCX;
      const uint32_t small_bytes = (uint32_t)pl.KZ * 256u;                 // [KZ x 64] hi | lo
      const uint32_t last_bytes = (uint32_t)(pl.LPC * pl.KZ) * 256u;       // XL tile hi | lo
      const int NLC = pl.NLC, si = pl.si;
      const long long t_xl = (long long)(si + 1 + H) * plan_x0_floats(pl);                      // first XL tile
      const long long t_bc = t_xl + (long long)NLC * plan_xl_floats(pl);                        // first BCt tile
      const long long t_b0 = t_bc + (long long)(H + 1) * plan_x0_floats(pl);                    // first B0t tile
      auto put = [&](const float* src, uint32_t bytes) {
        const int s = (int)(g % nst);
        mbar_wait(&b_empty[s], (uint32_t)(((g / nst) & 1) ^ 1));
        TRACE(3, g);
        mbar_expect_tx(&b_full[s], bytes);
        bulk_g2s(Bst + s * TCB_STAGE_BYTES, src, bytes, &b_full[s]);
        ++g;
      };
      for (long long t = 0; t < my_pairs; ++t) {
        for (int q = 0; q < NLC; ++q) put(tcx + t_xl + (long long)q * plan_xl_floats(pl), last_bytes);
        for (int h = H - 1; h >= 0; --h) {
          put(tcx + t_bc + (long long)(h + 1) * plan_x0_floats(pl), small_bytes);
          for (int c = 0; c < NCH; ++c) put(src0 + ((long long)h * NCH + c) * NIF_TC_CHUNK_FLOATS, TCB_STAGE_BYTES);
        }
        put(tcx + t_bc, small_bytes);
        for (int i = 0; i < si; ++i) put(tcx + t_b0 + (long long)i * plan_x0_floats(pl), small_bytes);
      }
    }
  } else if (warp == 8) {
    tcb_issue<0>(pl, a, smem, bars, tmem, my_pairs);
  } else if (warp == 10) {
    tcb_issue<1>(pl, a, smem, bars, tmem, my_pairs);
  }
  } else {
    tc_reg_inc<224>();
    // ---------------- epilogue warps: thread r <-> row r of tile wg <-> TMEM lane r ----------------
    const int wg = warp >> 2;
    const int r = tid & 127;
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)wg * 256u;
    unsigned char* A_hi = A_all + wg * 2 * TC_TILE_BYTES;
    unsigned char* A_lo = A_hi + TC_TILE_BYTES;
    float* zs = zs_all + wg * K1 * 128;
    float* dzs = dzs_all + wg * K1 * 128;
    const float* invB = a.packed + pl.off_TCS;
    const long long slot_floats = nif_tiled_rows(a.B) * 64;  // one slot of the (tiled) stash / da buffer
    long long g = 0;
    for (long long p = 0; p < my_pairs; ++p) {
      const long long row0 = ((blockIdx.x + p * gridDim.x) * 2 + wg) * 128;
      const long long b = row0 + r;
      const bool live = b < a.B;
      named_bar_sync(1 + wg, 128);
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
      for (int kk = 0; kk < K1; ++kk) dzs[kk * 128 + r] = 0.f;
      named_bar_sync(1 + wg, 128);
      {
        float zmax = 0.f;
        for (int kk = 0; kk < K1; ++kk) zmax = fmaxf(zmax, fabsf(zs[kk * 128 + r]));
        warp_atomic_max(&a.maxes[H + 1], live ? zmax : 0.f);
      }
      const float* invX = a.packed + pl.off_TCS2;  // inverse scales of the small tiles
      const int T_xl = si + 1 + H, T_bc = T_xl + NLC, T_b0 = T_bc + H + 1;
      // accumulator stage g & 1 of this tile: wait for its MMAs; returns its TMEM address
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
      // an N = KZ chunk: dz[kappa] += coef * D[col0 + kappa]
      auto drain_kz = [&](float coef, uint32_t col0) {
        const uint32_t td = chunk_begin();
        for (int k0 = 0; k0 < KZ; k0 += 16) {
          float v[16];
          tc_ld16(td + col0 + (uint32_t)k0, v);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (k0 + e < K1) dzs[(k0 + e) * 128 + r] = fmaf(coef, v[e], dzs[(k0 + e) * 128 + r]);
        }
      };

      float dy[NIF_MAX_SO];
#pragma unroll
      for (int c = 0; c < NIF_MAX_SO; ++c) dy[c] = (c < so && live) ? __ldg(&a.du[b * so + c]) : 0.f;
      // ---- operand tile h_{H+1} (the last matrix's dz terms) and the bounds the thin weight-gradient kernel scales with ----
      float inv_hl;
      {
        float hl[64];
        const float* hsrc = (EXT ? a.h_stash : a.save) + (long long)H * slot_floats + nif_tiled_row(b);
        float hmax = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) q = ldg4(hsrc + c * 128);
          hl[4 * c] = q.x; hl[4 * c + 1] = q.y; hl[4 * c + 2] = q.z; hl[4 * c + 3] = q.w;
          hmax = fmaxf(fmaxf(hmax, fabsf(q.x)), fmaxf(fabsf(q.y), fmaxf(fabsf(q.z), fabsf(q.w))));
        }
        warp_atomic_max(&a.maxes[NIF_MAX_HL], hmax);
        float sc_h;
        tc_row_scale(hmax, sc_h, inv_hl);
        tc_store_row_split(A_hi, A_lo, r, hl, sc_h);
        fence_async_smem();
        mbar_arrive(&a_ready[wg]);
        for (int i = 0; i < si; ++i) warp_atomic_max(&a.maxes[NIF_MAX_X + i], live ? fabsf(__ldg(&a.x[b * si + i])) : 0.f);
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c)
          if (c < so) warp_atomic_max(&a.maxes[NIF_MAX_DU + c], fabsf(dy[c]));
      }

      float acc[64];
      // ---- last matrix (n -> so): dh_{H+1}[i] = sum_kappa zt[kappa] sum_c ML[kappa][i][c] du[c]  (CUDA cores) ----
      {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        const float* ML = a.packed + pl.off_ML;
        for (int kk = 0; kk < K1; ++kk) {
          const float zk = zs[kk * 128 + r];
          const float* Mk = ML + (long long)kk * 64 * so;
          if (so == 1) {
            const float w = zk * dy[0];
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
              const float4 q = ldg4(Mk + 4 * c4);
              acc[4 * c4] = fmaf(w, q.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(w, q.y, acc[4 * c4 + 1]);
              acc[4 * c4 + 2] = fmaf(w, q.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(w, q.w, acc[4 * c4 + 3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              float s = 0.f;
#pragma unroll
              for (int c = 0; c < NIF_MAX_SO; ++c)
                if (c < so) s = fmaf(dy[c], __ldg(&Mk[i * so + c]), s);
              acc[i] = fmaf(zk, s, acc[i]);
            }
          }
        }
      }

      // ---- last matrix, its dz terms: dz[kappa] += du[c] * ((h_{H+1} @ ML[kappa])[c] + CL[kappa][c]) ----
      {
        const float* CL = a.packed + pl.off_C + (long long)(H + 1) * K1 * 64;
        for (int q = 0; q < NLC; ++q) {
          const float sL = inv_hl * __ldg(&invX[T_xl + q]);
          const uint32_t td = chunk_begin();
          for (int cl = 0; cl < LPC; ++cl) {
            const int c = LPC * q + cl;
            float dyc = 0.f;
#pragma unroll
            for (int cc = 0; cc < NIF_MAX_SO; ++cc) if (cc == c) dyc = dy[cc];
            for (int k0 = 0; k0 < KZ; k0 += 16) {
              float v[16];
              tc_ld16(td + (uint32_t)(cl * KZ + k0), v);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int kk = k0 + e;
                if (kk < K1 && c < so) {
                  const float cl = NO_BIAS ? 0.f : __ldg(&CL[(long long)kk * 64 + c]);
                  dzs[kk * 128 + r] = fmaf(dyc, fmaf(sL, v[e], cl), dzs[kk * 128 + r]);
                }
              }
            }
          }
          chunk_end();
        }
      }

      // ---- layers H .. 0: da_m = dh_{m+1} * d_m; hidden matrices on the tensor cores ----
      for (int m = H; m >= 0; --m) {
        const float* dsv = a.save + (long long)(H + 1 + m) * slot_floats + nif_tiled_row(b);  // d_m row (tiled)
        float* dag = a.da + (long long)m * slot_floats + nif_tiled_row(b);
        float dav[64];
        float amax = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {  // column quad c: a warp moves 512 contiguous bytes per access
          float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) dv = ldg4(dsv + c * 128);
          dav[4 * c] = acc[4 * c] * dv.x; dav[4 * c + 1] = acc[4 * c + 1] * dv.y;
          dav[4 * c + 2] = acc[4 * c + 2] * dv.z; dav[4 * c + 3] = acc[4 * c + 3] * dv.w;
          if (EXT && live) {
            const long long eo = (long long)m * slot_floats + nif_tiled_row(b) + c * 128;
            if (MODE == 3) {  // + dh'_{m+1} * e_m, written by the tangent-adjoint pass
              const float4 xv = ldg4(a.ext_add + eo);
              dav[4 * c] += xv.x; dav[4 * c + 1] += xv.y; dav[4 * c + 2] += xv.z; dav[4 * c + 3] += xv.w;
            } else {
              const float4 ev = ldg4(a.e_stash + eo);
              float4 xv = make_float4(acc[4 * c] * ev.x, acc[4 * c + 1] * ev.y, acc[4 * c + 2] * ev.z, acc[4 * c + 3] * ev.w);
              float4* xo = reinterpret_cast<float4*>(a.ext_out + eo);
              if (MODE == 2) { const float4 o = *xo; xv.x += o.x; xv.y += o.y; xv.z += o.z; xv.w += o.w; }
              *xo = xv;
            }
          }
          if (live) *reinterpret_cast<float4*>(dag + c * 128) = make_float4(dav[4 * c], dav[4 * c + 1], dav[4 * c + 2], dav[4 * c + 3]);
          amax = fmaxf(fmaxf(amax, fabsf(dav[4 * c])), fmaxf(fabsf(dav[4 * c + 1]), fmaxf(fabsf(dav[4 * c + 2]), fabsf(dav[4 * c + 3]))));
        }
        warp_atomic_max(&a.maxes[m], amax);
        if (r == 0) TRACE(wg, g * 4 + 3);  // da_m formed and stored

        float sc_a, inv_a;
        tc_row_scale(amax, sc_a, inv_a);
        tc_store_row_split(A_hi, A_lo, r, dav, sc_a);
        fence_async_smem();
        mbar_arrive(&a_ready[wg]);
        // bias rows of layer m: dz[kappa] += sum_j C_m[kappa][j] da_m[j]
        drain_kz(NO_BIAS ? 0.f : inv_a * __ldg(&invX[T_bc + m]), 0u);
        chunk_end();
        if (m == 0) {  // first matrix: dz[kappa] += omega x[i] sum_j M0[kappa][i][j] da_0[j]
          const float om0 = plan_omega(pl, 0);
          for (int i = 0; i < si; ++i) {
            drain_kz(om0 * (live ? __ldg(&a.x[b * si + i]) : 0.f) * inv_a * __ldg(&invX[T_b0 + i]), 0u);
            chunk_end();
          }
          break;
        }

        // this layer's input row h_m (for the dz dot products)
        float hm[64];
        {
          const float* hsrc = (EXT ? a.h_stash : a.save) + (long long)(m - 1) * slot_floats + nif_tiled_row(b);
          float hmax = 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) q = ldg4(hsrc + c * 128);
            hm[4 * c] = q.x; hm[4 * c + 1] = q.y; hm[4 * c + 2] = q.z; hm[4 * c + 3] = q.w;
            hmax = fmaxf(fmaxf(hmax, fabsf(q.x)), fmaxf(fabsf(q.y), fmaxf(fabsf(q.z), fabsf(q.w))));
          }
          warp_atomic_max(&a.maxes[H + 2 + (m - 1)], hmax);
        }
        const int res = plan_res(pl, m);
        if (res != 1) {  // res == 1 (NIF hidden layer): dh_m = T + dh_{m+1}, keep acc
#pragma unroll
          for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        }
        const float om_inv = plan_omega(pl, m) * inv_a;
        const float* invBm = invB + (m - 1) * KP;
        // per-chunk row coefficients, fetched one chunk ahead (their latency hides behind the accumulator wait)
        float sBn[2] = {om_inv * __ldg(&invBm[0]), om_inv * __ldg(&invBm[1])};
        float zkn[2] = {zs[r], zs[128 + r]};
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const float sBc[2] = {sBn[0], sBn[1]}, zkc[2] = {zkn[0], zkn[1]};
          if (c + 1 < NCH) {
            sBn[0] = om_inv * __ldg(&invBm[2 * c + 2]); sBn[1] = om_inv * __ldg(&invBm[2 * c + 3]);
            zkn[0] = zs[(2 * c + 2) * 128 + r];
            zkn[1] = 2 * c + 3 < K1 ? zs[(2 * c + 3) * 128 + r] : 0.f;  // the padding coordinate of an odd K + 1
          }
          const uint32_t td = chunk_begin();
#pragma unroll
          for (int kl = 0; kl < 2; ++kl) {
            const int kk = 2 * c + kl;
            const float sB = sBc[kl];
            const float zo = zkc[kl] * sB;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {  // 32 columns (values of i) at a time
              float v1[16], v2[16];
              const uint32_t col = (uint32_t)(kl * 64 + q * 32);
              tc_ld16(td + col, v1);
              tc_ld16(td + col + 16u, v2);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                acc[q * 32 + e] = fmaf(zo, v1[e], acc[q * 32 + e]);
                acc[q * 32 + 16 + e] = fmaf(zo, v2[e], acc[q * 32 + 16 + e]);
                s0 = fmaf(v1[e], hm[q * 32 + e], s0);
                s1 = fmaf(v2[e], hm[q * 32 + 16 + e], s1);
              }
            }
            if (kk < K1) dzs[kk * 128 + r] += sB * (s0 + s1);
          }
          chunk_end();
        }
      }

      if (live) {
        if (EXT && a.dz_accumulate) {
          for (int kk = 0; kk < K; ++kk) a.dz[b * K + kk] += dzs[kk * 128 + r];
        } else {
          for (int kk = 0; kk < K; ++kk) a.dz[b * K + kk] = dzs[kk * 128 + r];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------
// Hooks of the reverse-over-forward passes (null = the plain reverse pass); see TcBwdArgs.
struct TcBwdExt {
  const float *h_stash, *e_stash, *ext_add;
  float* ext_out;
  int no_bias, dz_accumulate, ext_accumulate;
};

int nif_tc_bwd_data_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                         const float* save, const float* du, float* da, float* dz, unsigned* maxes,
                         cudaStream_t st, const TcBwdExt* ext = nullptr) {
  if (!nif_plan_uses_tc(pl)) return NIF_E_UNSUPPORTED;
  TcBwdArgs a;
  a.h_stash = save; a.e_stash = nullptr; a.ext_out = nullptr; a.ext_add = nullptr; a.dz_accumulate = 0;
  if (ext) {
    a.h_stash = ext->h_stash ? ext->h_stash : save; a.e_stash = ext->e_stash; a.ext_out = ext->ext_out; a.ext_add = ext->ext_add;
    a.dz_accumulate = ext->dz_accumulate;
  }
  a.nst = tcb_smem_bytes(pl.K, 3) <= 227 * 1024 ? 3 : 2;
  const size_t smem = tcb_smem_bytes(pl.K, a.nst);
  a.B = B;
  a.total_pairs = (B + 255) / 256;
  a.z = z; a.x = x; a.packed = packed; a.save = save; a.du = du; a.da = da; a.dz = dz; a.maxes = maxes;
  NIF_CUDA_CHECK(cudaMemsetAsync(maxes, 0, sizeof(unsigned) * 256, st));
  auto kern = nif_tc_bwd_data_kernel<0>;
  if (ext) {  // the two passes of nif_sobolev_backward_impl
    if (ext->ext_out && ext->no_bias && ext->e_stash && !ext->ext_add) kern = ext->ext_accumulate ? nif_tc_bwd_data_kernel<2> : nif_tc_bwd_data_kernel<1>;
    else if (ext->ext_add && !ext->ext_out && !ext->no_bias) kern = nif_tc_bwd_data_kernel<3>;
    else { nif_set_error("nif_tc_bwd_data: unsupported combination of reverse-over-forward hooks"); return NIF_E_UNSUPPORTED; }
  }
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = sms;
  if (grid > a.total_pairs) grid = a.total_pairs;
  { NIF_PROF(ext ? "nif_tc_bwd_data_kernel<ext>" : "nif_tc_bwd_data_kernel", st); kern<<<(unsigned)grid, TCB_THREADS, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient of the hidden matrices on the tensor cores:
//   dM_m[kappa][i][j] = omega * sum_b zt[b][kappa] h_m[b][i] da_m[b][j]
// A batch-reduction GEMM: D[(kappa_l, i)][j] += A[(kappa_l, i)][b] * B[j][b] with the batch index b as the
// MMA K dimension.  Both operands are generated on the fly by the row-owning threads (A = zt (x) h, B = da),
// which is naturally "MN-major" (for a fixed row b the (kappa_l, i) / j index is contiguous).  Because b is the
// reduction index the power-of-two scales must be constant over b: one global scale per layer from the maxima
// the data pass recorded (max|zt| * max|h_m| and max|da_m|).
// One CTA = (hidden matrix, group of 4 latent coordinates, batch split); 64-row sub-tiles, two operand slots
// (generation of sub-tile t+1 overlaps the MMAs of sub-tile t); accumulators stay in TMEM for the whole batch
// range; the result is written as a partial in the layout nif_unpack_grad_kernel sums.
// ---------------------------------------------------------------------------------------------------
struct TcWgtArgs {
  long long B, rows_per_split;
  int S;
  const float *z, *save, *da;
  const unsigned* maxes;
  float* part;  // [S][H][K+1][64][64]
};

#define TCW_THREADS 288
#define TCW_A_BYTES 16384u   // [128 (kappa_l, i) x 64 (b)] fp16, MN-major
#define TCW_B_BYTES 8192u    // [64 (j) x 64 (b)] fp16, MN-major
#define TCW_SLOT_BYTES (4 * TCW_A_BYTES + 2 * TCW_B_BYTES)  // A[q][hi|lo], B[hi|lo]
#define TCW_STAGE_BYTES 16384u                     // the h rows of one 64-row sub-tile: two 32-row groups of the tiled stash

// MN-major core-matrix layout of a [MN x 64] tile: 8 (k) x 16 B (8 consecutive mn) core matrices,
// offset(mn, k) = (mn/8) * 1024 + (k/8) * 128 + (k%8) * 16 + (mn%8) * 2
__device__ __forceinline__ uint64_t tcw_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;   // LBO: next group of 8 k
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;  // SBO: next group of 8 mn
  d |= (uint64_t)1 << 46;
  return d;
}

// 8 values -> scaled hi / lo fp16 chunks (16 bytes each)
__device__ __forceinline__ void tcw_split8(const float (&v)[8], float sc, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a0 = v[2 * e] * sc, a1 = v[2 * e + 1] * sc;
    const __half2 hh = __floats2half2_rn(a0, a1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
    h[e] = *reinterpret_cast<const uint32_t*>(&hh);
    l[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(TCW_THREADS, 1) nif_tc_bwd_weight_kernel(const Plan pl, const TcWgtArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t slot_full[2], slot_empty[2], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = pl.K, K1 = pl.K + 1, H = pl.H;
  const int NPG = (pl.KP + 3) / 4;  // groups of 4 latent coordinates
  const int h = blockIdx.x / NPG, pg = blockIdx.x % NPG;
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
  if (warp == 8) tc_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // global power-of-two operand scales of this layer (m = h + 1)
  float scA, invA, scB, invB;
  {
    const float maxz = __uint_as_float(a.maxes[H + 1]);
    const float maxh = __uint_as_float(a.maxes[H + 2 + h]);
    const float maxd = __uint_as_float(a.maxes[h + 1]);
    tc_row_scale(maxz * maxh, scA, invA);
    tc_row_scale(maxd, scB, invB);
  }

  if (warp == 8) {
    {  // MMA issuer: the whole warp runs the loop (uniform operands), one elected lane issues -- see tc_elect_one
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      for (long long t = 0; t < nsub; ++t) {
        const int sl = (int)(t & 1);
        mbar_wait(&slot_full[sl], (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        if (tc_elect_one()) {
        const uint32_t base = smem_u32(smem + sl * TCW_SLOT_BYTES);
        const uint64_t b_hi = tcw_make_desc(base + 4 * TCW_A_BYTES), b_lo = tcw_make_desc(base + 4 * TCW_A_BYTES + TCW_B_BYTES);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint64_t a_hi = tcw_make_desc(base + (2 * q) * TCW_A_BYTES), a_lo = tcw_make_desc(base + (2 * q + 1) * TCW_A_BYTES);
          const uint32_t d1 = tmem_u + (uint32_t)q * 128u, d2 = d1 + 64u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 16 rows (k) per instruction = 2 k-groups = 256 B
            const uint64_t adv = (uint64_t)(ks * 16);
            const uint32_t accf = (t > 0 || ks > 0) ? 1u : 0u;
            tc_mma_f16(d2, a_lo + adv, b_hi + adv, idesc, accf);
            tc_mma_f16(d2, a_hi + adv, b_lo + adv, idesc, 1u);
            tc_mma_f16(d1, a_hi + adv, b_hi + adv, idesc, accf);
          }
        }
        tc_commit(&slot_empty[sl]);
        }
        __syncwarp();
      }
      if (tc_elect_one()) tc_commit(&done_bar);
      __syncwarp();
    }
  } else {
    // ---------------- operand generators: slot sl = warp / 4, thread (q, r) = pair q, row r of the sub-tile ----------------
    const int sl = warp >> 2;
    const int q = (tid >> 6) & 1, r = tid & 63;
    unsigned char* slot = smem + sl * TCW_SLOT_BYTES;
    unsigned char* Aq_hi = slot + (2 * q) * TCW_A_BYTES;
    unsigned char* Aq_lo = Aq_hi + TCW_A_BYTES;
    unsigned char* B_hi = slot + 4 * TCW_A_BYTES;
    unsigned char* B_lo = B_hi + TCW_B_BYTES;
    const uint32_t koff = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;  // this row's position along k
    const long long slot_floats = nif_tiled_rows(a.B) * 64;
    const float* hsrc = a.save + (long long)h * slot_floats;        // h_m, m = h + 1 -> stash slot h (tiled layout)
    const float* dsrc = a.da + (long long)(h + 1) * slot_floats;    // da_m (tiled layout)
    const int kk0 = 4 * pg + 2 * q;
    // fp32 row staging: in the tiled stash a 64-row sub-tile is one contiguous 16 KB block; the slot's 128 threads
    // fetch it with cp.async (8 x 16 B each, fully coalesced), and the copy of sub-tile t+2 flies while sub-tile t is
    // converted.  Row r of the block reads its column quads at (r >> 5) * 8 KB + quad * 512 + (r & 31) * 16:
    // consecutive rows hit consecutive banks.
    unsigned char* stg = smem + 2 * TCW_SLOT_BYTES + sl * TCW_STAGE_BYTES;
    const int ts = tid & 127;  // thread index inside the slot
    auto stage_rows = [&](long long t) {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(hsrc + nif_tiled_row(r0 + t * 64));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t off = (uint32_t)(ts + i * 128) * 16u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(stg + off)), "l"(src + off) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (sl < nsub) stage_rows(sl);
    // this thread's row data that comes straight from global memory -- its latent coordinates (pre-multiplied by the
    // operand scale) and its half of da_m (32 values) -- is fetched one sub-tile ahead, so that its latency hides
    // behind the conversion of the current sub-tile
    float ztn[2];
    float4 dqn[8];
    auto fetch_row = [&](long long t) {
      const long long b = r0 + t * 64 + r;
      const bool live = t < nsub && b < r1;
#pragma unroll
      for (int kl = 0; kl < 2; ++kl) {
        const int kk = kk0 + kl;
        ztn[kl] = 0.f;
        if (live) ztn[kl] = (kk < K) ? __ldg(&a.z[b * K + kk]) : (kk == K ? 1.f : 0.f);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) dqn[c] = live ? ldg4(dsrc + nif_tiled_row(b) + (8 * q + c) * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    fetch_row(sl);
    long long n_mine = 0;
    for (long long t = sl; t < nsub; t += 2, ++n_mine) {
      const long long b = r0 + t * 64 + r;
      const bool live = b < r1;
      const float zt[2] = {ztn[0] * scA, ztn[1] * scA};
      float4 hq[16], dq[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) dq[c] = dqn[c];
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      named_bar_sync(1 + sl, 128);  // every thread's part of the block has landed
#pragma unroll
      for (int c = 0; c < 16; ++c)
        hq[c] = live ? *reinterpret_cast<const float4*>(stg + (r >> 5) * 8192 + c * 512 + (r & 31) * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
      named_bar_sync(1 + sl, 128);  // every thread of the slot has its row in registers
      if (t + 2 < nsub) stage_rows(t + 2);
      fetch_row(t + 2);
      // wait until the MMAs that read this slot two sub-tiles ago have completed
      mbar_wait(&slot_empty[sl], (uint32_t)((n_mine & 1) ^ 1));
#pragma unroll
      for (int ig = 0; ig < 8; ++ig) {
        const float4 p0 = hq[2 * ig], p1 = hq[2 * ig + 1];
        const float hv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int kl = 0; kl < 2; ++kl) {
          float pv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) pv[e] = zt[kl] * hv[e];  // zt carries the operand scale (a power of two: exact)
          uint4 hi, lo;
          tcw_split8(pv, 1.f, hi, lo);
          const uint32_t off = (uint32_t)(kl * 8 + ig) * 1024u + koff;
          *reinterpret_cast<uint4*>(Aq_hi + off) = hi;
          *reinterpret_cast<uint4*>(Aq_lo + off) = lo;
        }
      }
#pragma unroll
      for (int jg = 0; jg < 4; ++jg) {  // pair q writes columns j = 32 q .. 32 q + 31 of the B operand
        const int j0 = 32 * q + 8 * jg;
        const float4 p0 = dq[2 * jg], p1 = dq[2 * jg + 1];
        const float dv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        uint4 hi, lo;
        tcw_split8(dv, scB, hi, lo);
        const uint32_t off = (uint32_t)(j0 >> 3) * 1024u + koff;
        *reinterpret_cast<uint4*>(B_hi + off) = hi;
        *reinterpret_cast<uint4*>(B_lo + off) = lo;
      }
      fence_async_smem();
      mbar_arrive(&slot_full[sl]);
    }
    // ---------------- final epilogue (warps 0-3): TMEM lane = (kappa_l, i) row of the gradient ----------------
    if (warp < 4) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
      const int row = tid;  // 0..127
      const int kl = row >> 6, i = row & 63;
      const float scale = plan_omega(pl, h + 1) * invA * invB;
      const uint32_t tm = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int kk = 4 * pg + 2 * qq + kl;
        float* dst = a.part + ((((long long)s * H + h) * K1 + kk) * 64 + i) * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          float v1[16], v2[16];
          tc_ld16(tm + (uint32_t)(qq * 128 + c0), v1);
          tc_ld16(tm + (uint32_t)(qq * 128 + 64 + c0), v2);
          tc_wait_ld();
          if (kk < K1) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              *reinterpret_cast<float4*>(dst + c0 + e) =
                  make_float4(nsub ? scale * (v1[e] + v2[e]) : 0.f, nsub ? scale * (v1[e + 1] + v2[e + 1]) : 0.f,
                              nsub ? scale * (v1[e + 2] + v2[e + 2]) : 0.f, nsub ? scale * (v1[e + 3] + v2[e + 3]) : 0.f);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 256);
}

int nif_tc_bwd_weight_impl(const Plan& pl, long long B, const float* z, const float* save, const float* da,
                           const unsigned* maxes, int S, long long rows_per_split, float* part, cudaStream_t st) {
  if (!nif_plan_uses_tc(pl)) return NIF_E_UNSUPPORTED;
  TcWgtArgs a;
  a.B = B; a.rows_per_split = rows_per_split; a.S = S;
  a.z = z; a.save = save; a.da = da; a.maxes = maxes; a.part = part;
  const size_t smem = 2 * (size_t)TCW_SLOT_BYTES + 2 * (size_t)TCW_STAGE_BYTES;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_tc_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int NPG = (pl.KP + 3) / 4;
  dim3 grid((unsigned)(pl.H * NPG), (unsigned)S);
  { NIF_PROF("nif_tc_bwd_weight_kernel", st); nif_tc_bwd_weight_kernel<<<grid, TCW_THREADS, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
