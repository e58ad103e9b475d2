// ParameterNet trunk on the tensor cores (tcgen05 + TMEM), fp32-grade: bf16x3.
//
// Reference: _call_parameter_net up to the bottleneck (nif/model.py:326-343, 176-216, 668-720) over
//   Dense(act) -> l x MLP_SimpleShortCut (nif/layers/mlp.py:148-160:  h + act(h W + b)) -> Dense(latent, linear).
// A 64-wide shared-weight MLP: per 128-row tile every layer is one [128 x 64] @ [64 x 64] product.  Each fp32 operand
// value is split into three bfloat16 parts  x = x0 + x1 + x2  (24 significant bits, and bf16 has the exponent range of
// fp32, so no operand scales are needed -- which is what lets the batch-reduced weight gradients share this scheme: a
// scale would have to be constant over the batch) and a product keeps the six terms down to 2^-16:
//   x2*y0, x0*y2, x1*y1, x1*y0, x0*y1, x0*y0      (smallest first; fp32 accumulation in TMEM)
// Twice the tensor work of the FP16x3 split of nif_tc.cuh, on a part of the step that has 3 % of its flops.
//
//   nif_trunk_tc_pack_kernel  theta -> bf16x3 operand tiles (K-major core-matrix layout), once per step
//   nif_trunk_tc_fwd_kernel   persistent, pairs of 128-row tiles ping-pong on the tensor pipe; weights resident in
//                             shared memory; thread = row; writes z and the stash (h_m, act'(pre_m); tiled layout)
//   nif_trunk_tc_bwd_kernel   persistent, one tile at a time: reverse data pass AND every parameter gradient.  Per layer
//                             the tile da_m [128 b x 64 j] is written once (3 parts) and read twice: as the K-major A
//                             operand of  dh += da_m W_m^T  and as the MN-major B operand of the batch reduction
//                               D_m[(i | 1 | p)][j] += [h_{m-1} | 1 | p_in]^T da_m
//                             whose rows are dW_m, db_m and (layer 0) dW_first: the thin terms ride in the unused half of
//                             the M = 128 operand.  Accumulators stay in TMEM across the CTA's tiles (at most 4096 rows
//                             per chain), then go to a per-CTA partial that nif_trunk_tc_reduce_kernel sums.
#include "nif_bf.cuh"

struct TrunkGeo {
  int pi, K, n, l, act;   // inputs, latent, units, hidden layers, activation
  int KB;                 // latent rounded up to 16
  long long P;            // floats of theta
  long long off_Wh, off_Wb, off_b0, off_bh, off_bb;  // theta offsets (W_first at 0)
  long long f_bytes, r_bytes;                        // forward / reverse operand images
};

__host__ __device__ inline TrunkGeo trunk_geo(int pi, int K, int n, int l, int act) {
  TrunkGeo g;
  g.pi = pi; g.K = K; g.n = n; g.l = l; g.act = act;
  g.KB = (K + 15) / 16 * 16;
  g.off_Wh = (long long)pi * n;
  g.off_Wb = g.off_Wh + (long long)l * n * n;
  g.off_b0 = g.off_Wb + (long long)n * K;
  g.off_bh = g.off_b0 + n;
  g.off_bb = g.off_bh + (long long)l * n;
  g.P = g.off_bb + K;
  g.f_bytes = (long long)l * 3 * 8192 + 3LL * g.KB * 128;
  g.r_bytes = g.f_bytes;
  return g;
}

// ---- bf16x3 split -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bf3_split2(float a, float b, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(a, b);
  const float2 f0 = __bfloat1622float2(h0);
  const float ra = a - f0.x, rb = b - f0.y;  // exact
  const __nv_bfloat162 h1 = __floats2bfloat162_rn(ra, rb);
  const float2 f1 = __bfloat1622float2(h1);
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(ra - f1.x, rb - f1.y);
  p0 = *reinterpret_cast<const uint32_t*>(&h0);
  p1 = *reinterpret_cast<const uint32_t*>(&h1);
  p2 = *reinterpret_cast<const uint32_t*>(&h2);
}
// 8 values -> one 16-byte chunk per part
__device__ __forceinline__ void bf3_split8(const float (&v)[8], uint4& q0, uint4& q1, uint4& q2) {
  uint32_t a[4], b[4], c[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) bf3_split2(v[2 * e], v[2 * e + 1], a[e], b[e], c[e]);
  q0 = make_uint4(a[0], a[1], a[2], a[3]);
  q1 = make_uint4(b[0], b[1], b[2], b[3]);
  q2 = make_uint4(c[0], c[1], c[2], c[3]);
}

// generic no-swizzle descriptor: lbo = stride between core matrices along K, sbo = along M / N (bytes)
__device__ __forceinline__ uint64_t tk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// the six products of one bf16x3 GEMM: parts of A `pa` bytes apart, parts of B `pb` bytes apart (descriptor units of
// 16 B), K extent 16 * ksteps with `kadv_a` / `kadv_b` descriptor units per step.  `fresh`: the accumulator starts here.
__device__ __forceinline__ void bf3_mma(uint32_t d, uint64_t da, uint32_t pa, uint32_t kadv_a, uint64_t db, uint32_t pb,
                                        uint32_t kadv_b, uint32_t idesc, int ksteps, bool fresh) {
  const int ia[6] = {2, 0, 1, 1, 0, 0}, ib[6] = {0, 2, 1, 0, 1, 0};
  uint32_t acc = fresh ? 0u : 1u;
#pragma unroll
  for (int t = 0; t < 6; ++t)
    for (int ks = 0; ks < ksteps; ++ks) {
      tc_mma_f16(d, da + (uint64_t)(ia[t] * pa + ks * kadv_a), db + (uint64_t)(ib[t] * pb + ks * kadv_b), idesc, acc);
      acc = 1u;
    }
}

// ---- operand images --------------------------------------------------------------------------------------------------
// forward  F: k = 1..l: 3 parts x [64 rows j x 64 (i)]  W_k[i][j];   then 3 parts x [KB rows kk x 64 (i)]  W_b[i][kk]
// reverse  R: 3 parts x [64 rows i x KB (kk)]  W_b[i][kk];   then k = l..1: 3 parts x [64 rows i x 64 (j)]  W_k[i][j]
// K-major core-matrix layout with K extent KD:  offset(row, k) = (row/8) * (KD/8*128) + (k/8) * 128 + (row%8) * 16 + (k%8) * 2
__global__ void __launch_bounds__(256) nif_trunk_tc_pack_kernel(const TrunkGeo g, const float* __restrict__ theta,
                                                                float* __restrict__ packed) {
  const long long slots = (g.f_bytes + g.r_bytes) / 4;  // one slot = two bf16 along K
  const int n = g.n, l = g.l, K = g.K, KB = g.KB;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < slots; e += 256LL * gridDim.x) {
    long long r = e;
    const bool fwd = r < g.f_bytes / 4;
    if (!fwd) r -= g.f_bytes / 4;
    // locate (matrix, part, slot in tile)
    const long long hid = (long long)l * 3 * 2048, bot = 3LL * KB * 32;
    int mat, part, t, rows_k;  // mat: 1..l hidden, 0 bottleneck; rows_k: K extent of the tile
    bool is_bot;
    if (fwd) {
      is_bot = r >= hid;
      if (!is_bot) { mat = 1 + (int)(r / (3 * 2048)); r %= 3 * 2048; part = (int)(r / 2048); t = (int)(r % 2048); }
      else { r -= hid; mat = 0; part = (int)(r / (KB * 32)); t = (int)(r % (KB * 32)); }
      rows_k = 64;
    } else {
      is_bot = r < bot;
      if (is_bot) { mat = 0; part = (int)(r / (KB * 32)); t = (int)(r % (KB * 32)); rows_k = KB; }
      else { r -= bot; mat = l - (int)(r / (3 * 2048)); r %= 3 * 2048; part = (int)(r / 2048); t = (int)(r % 2048); rows_k = 64; }
    }
    const int per_rg = (rows_k / 8) * 32;
    const int rg = t / per_rg; t %= per_rg;
    const int kc = t / 32; t %= 32;
    const int row = rg * 8 + t / 4, k = kc * 8 + (t % 4) * 2;
    float w[2] = {0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int i, j;  // W[i][j] (in, out)
      if (fwd) { j = row; i = k + q; } else { i = row; j = k + q; }
      if (is_bot) { if (i < n && j < K) w[q] = theta[g.off_Wb + (long long)i * K + j]; }
      else if (i < n && j < n) w[q] = theta[g.off_Wh + (long long)(mat - 1) * n * n + (long long)i * n + j];
    }
    uint32_t p0, p1, p2;
    bf3_split2(w[0], w[1], p0, p1, p2);
    packed[e] = __uint_as_float(part == 0 ? p0 : (part == 1 ? p1 : p2));
  }
}

// value and first derivative of 8 activations at once, branch-free and inline: with two epilogue warps per scheduler the
// kernel is bound by instruction latency, so the eight independent chains have to interleave (the out-of-line act_fd4 of
// the CUDA-core kernels serialises four).  exp via ex2.approx (relative error 2^-22), 1/(1+e) via the approximate
// reciprocal (2 ulp): far inside the 1e-5 parity gate.
template <int ACT>
__device__ __forceinline__ void trunk_act8(const float (&v)[8], float (&f)[8], float (&d)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float x = v[e];
    if (ACT == NIF_ACT_SWISH) {
      const float sg = __fdividef(1.f, 1.f + __expf(-x));
      f[e] = x * sg;
      d[e] = sg * fmaf(x, 1.f - sg, 1.f);
    } else if (ACT == NIF_ACT_SIGMOID) {
      const float sg = __fdividef(1.f, 1.f + __expf(-x));
      f[e] = sg;
      d[e] = sg * (1.f - sg);
    } else if (ACT == NIF_ACT_TANH) {
      const float t = tanhf(x);
      f[e] = t;
      d[e] = fmaf(-t, t, 1.f);
    } else if (ACT == NIF_ACT_RELU) {
      f[e] = x > 0.f ? x : 0.f;
      d[e] = x > 0.f ? 1.f : 0.f;
    } else {
      f[e] = x;
      d[e] = 1.f;
    }
  }
}

// ---- forward -----------------------------------------------------------------------------------------------------------
struct TrunkFwdArgs {
  long long B, total_pairs;
  const float *p_in, *theta, *packed;
  float *z, *save;
};
#define TKF_THREADS 288

__host__ __device__ inline size_t tkf_smem_bytes(const TrunkGeo& g) {
  return (size_t)g.f_bytes + 2 * 3 * 16384 + (size_t)(g.pi * 64 + (g.l + 1) * 64 + g.KB) * 4 + 128;
}

template <int ACT>
__global__ void __launch_bounds__(TKF_THREADS, 1) nif_trunk_tc_fwd_kernel(const TrunkGeo g, const TrunkFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Wimg = smem;                       // forward image, resident
  unsigned char* A_all = smem + g.f_bytes;          // tile t: parts at t*48K + p*16K
  float* W0s = reinterpret_cast<float*>(A_all + 2 * 3 * 16384);  // [pi][64]
  float* bs = W0s + g.pi * 64;                      // [l+1][64]
  float* bbs = bs + (g.l + 1) * 64;                 // [KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bbs + g.KB);
  uint64_t* w_full = bars;       // [1]
  uint64_t* a_ready = bars + 1;  // [2]
  uint64_t* t_full = bars + 3;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = g.n, l = g.l, K = g.K, KB = g.KB, pi = g.pi;

  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_ready[i], 128); mbar_init(&t_full[i], 1); }
    mbar_fence_init();
  }
  if (warp == 8) tc_alloc(tmem_slot, 128);
  // small fp32 parameters (zero padded to 64 / KB columns)
  for (int e = tid; e < pi * 64; e += TKF_THREADS) { const int i = e / 64, j = e % 64; W0s[e] = j < n ? a.theta[(long long)i * n + j] : 0.f; }
  for (int e = tid; e < (l + 1) * 64; e += TKF_THREADS) {
    const int m = e / 64, j = e % 64;
    bs[e] = j < n ? a.theta[(m == 0 ? g.off_b0 : g.off_bh + (long long)(m - 1) * n) + j] : 0.f;
  }
  for (int e = tid; e < KB; e += TKF_THREADS) bbs[e] = e < K ? a.theta[g.off_bb + e] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  long long my_pairs = 0;
  if ((long long)blockIdx.x < a.total_pairs) my_pairs = (a.total_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == 8) {
    if (lane == 0 && my_pairs > 0) {  // the whole forward image, once (bulk copies of at most 32 KB)
      mbar_expect_tx(w_full, (uint32_t)g.f_bytes);
      for (long long o = 0; o < g.f_bytes; o += 32768) {
        const uint32_t nb = (uint32_t)(g.f_bytes - o < 32768 ? g.f_bytes - o : 32768);
        bulk_g2s(Wimg + o, reinterpret_cast<const unsigned char*>(a.packed) + o, nb, w_full);
      }
    }
    __syncwarp();
    if (my_pairs > 0) mbar_wait(w_full, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t ph = 0;  // a_ready phase (the same for both tiles: they advance in lockstep)
    for (long long p = 0; p < my_pairs; ++p)
      for (int s = 0; s <= l; ++s, ph ^= 1u)
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&a_ready[t], ph);
          tc_fence_after();
          const uint64_t dA = tk_desc(smem_u32(A_all + t * 3 * 16384), 128, 1024);
          const bool bot = s == l;
          const uint64_t dB = tk_desc(smem_u32(Wimg + (bot ? (long long)l * 3 * 8192 : (long long)s * 3 * 8192)), 128, 1024);
          const uint32_t pb = bot ? (uint32_t)(KB * 128 / 16) : 512u;  // B parts: KB*128 or 8192 bytes apart
          if (tc_elect_one()) {
            bf3_mma(tmem_u + (uint32_t)t * 64u, dA, 1024u, 16u, dB, pb, 16u, bf_idesc(bot ? KB : 64), 4, true);
            tc_commit(&t_full[t]);
          }
          __syncwarp();
        }
  } else {
    // ---------------- epilogue warps: thread = row r of tile t ----------------
    const int t = warp >> 2, r = tid & 127;
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)t * 64u;
    unsigned char* A0 = A_all + t * 3 * 16384;
    const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 16u;
    const long long slot_floats = nif_tiled_rows(a.B) * 64;
    uint32_t ph = 0;
    for (long long p = 0; p < my_pairs; ++p) {
      const long long b = ((blockIdx.x + p * gridDim.x) * 2 + t) * 128 + r;
      const bool live = b < a.B;
      float h[64];
      auto stash = [&](int m, const float (&dv)[64]) {
        if (!a.save || !live) return;
        float* sh = a.save + (long long)m * slot_floats + nif_tiled_row(b);
        float* sd = a.save + (long long)(l + 1 + m) * slot_floats + nif_tiled_row(b);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          *reinterpret_cast<float4*>(sh + c * 128) = make_float4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
          *reinterpret_cast<float4*>(sd + c * 128) = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
        }
      };
      auto publish = [&]() {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float v[8] = {h[8 * c], h[8 * c + 1], h[8 * c + 2], h[8 * c + 3], h[8 * c + 4], h[8 * c + 5], h[8 * c + 6], h[8 * c + 7]};
          uint4 q0, q1, q2;
          bf3_split8(v, q0, q1, q2);
          *reinterpret_cast<uint4*>(A0 + row_off + c * 128) = q0;
          *reinterpret_cast<uint4*>(A0 + 16384 + row_off + c * 128) = q1;
          *reinterpret_cast<uint4*>(A0 + 32768 + row_off + c * 128) = q2;
        }
        fence_async_smem();
        mbar_arrive(&a_ready[t]);
      };
      // ---- layer 0 (thin, CUDA cores): h = act(p W_first + b_first) ----
      {
        float pv[NIF_MAX_SI];
#pragma unroll
        for (int i = 0; i < NIF_MAX_SI; ++i) pv[i] = (i < pi && live) ? __ldg(&a.p_in[b * pi + i]) : 0.f;
        float dv[64];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float pre[8], f8[8], d8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float s = bs[8 * c + e];
#pragma unroll
            for (int i = 0; i < NIF_MAX_SI; ++i) if (i < pi) s = fmaf(pv[i], W0s[i * 64 + 8 * c + e], s);
            pre[e] = s;
          }
          trunk_act8<ACT>(pre, f8, d8);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const bool pad = 8 * c + e >= n;
            h[8 * c + e] = pad ? 0.f : f8[e];
            dv[8 * c + e] = pad ? 0.f : d8[e];
          }
        }
        stash(0, dv);
      }
      // ---- hidden layers: h += act(h W_k + b_k) ----
#pragma unroll 1
      for (int k = 1; k <= l; ++k, ph ^= 1u) {
        publish();
        mbar_wait(&t_full[t], ph);
        tc_fence_after();
        float dv[64];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float v[32];
          tc_ld32(tm + (uint32_t)(q * 32), v);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float pre[8], f8[8], d8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) pre[e] = v[8 * c + e] + bs[k * 64 + q * 32 + 8 * c + e];
            trunk_act8<ACT>(pre, f8, d8);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = q * 32 + 8 * c + e;
              const bool pad = j >= n;
              h[j] = pad ? 0.f : h[j] + f8[e];
              dv[j] = pad ? 0.f : d8[e];
            }
          }
        }
        tc_fence_before();
        stash(k, dv);
      }
      // ---- bottleneck: z = h W_b + b_b ----
      publish();
      mbar_wait(&t_full[t], ph);
      ph ^= 1u;
      tc_fence_after();
      for (int c0 = 0; c0 < KB; c0 += 16) {
        float v[16];
        tc_ld16(tm + (uint32_t)c0, v);
        tc_wait_ld();
        if (live) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c0 + e < K) a.z[b * K + c0 + e] = v[e] + bbs[c0 + e];
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 128);
}

// ---- reverse ------------------------------------------------------------------------------------------------------------
struct TrunkBwdArgs {
  long long B, total_tiles;
  const float *p_in, *packed, *save, *dz;
  float* part;     // [gridDim.x * nflush][P]
  int nflush;      // partial sets per CTA
  int flush_tiles; // tiles per accumulation chain
};
#define TKB_THREADS 320
#define TKB_DA_PART 16384u   // [64 mn (j) x 128 k (b)] bf16
#define TKB_A_PART 20480u    // [80 mn x 128 k (b)] bf16: the rows the batch reduction needs (see below)
#define TKB_W_STAGE 24576u   // 3 parts x 8 KB

__host__ __device__ inline size_t tkb_smem_bytes() { return 3 * TKB_DA_PART + 2 * 3 * TKB_A_PART + 2 * TKB_W_STAGE + 256; }

// Two threads per row: thread (r, half) owns row r (TMEM lane r) and the columns [32 half, 32 half + 32) of everything it
// drains or writes.  Warps 0-3: half 0, warps 4-7: half 1, warp 8: MMA issuer, warp 9: weight stream.
__global__ void __launch_bounds__(TKB_THREADS, 1) nif_trunk_tc_bwd_kernel(const TrunkGeo g, const TrunkBwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // da tile (3 parts):  offset(j, b) = (j/8) * 2048 + (b/8) * 128 + (b%8) * 16 + (j%8) * 2
  //   as B, MN-major (N = j, K = b):  LBO = 128, SBO = 2048;   as A, K-major (M = b, K = j):  LBO = 2048, SBO = 128
  unsigned char* DA = smem;
  // A tile, two buffers x 3 parts, MN-major [mn x 128 k (b)], same strides: mn < 64: h_{m-1}[b][mn]; mn = 64: 1;
  // mn = 65 + i: p_in[b][i].  The M = 128 instruction also reads rows 80..127, i.e. the 12 KB behind each 20 KB part: whatever
  // bytes lie there only reach accumulator rows 80..127, which nobody reads (TMEM lanes are independent).
  unsigned char* AT = smem + 3 * TKB_DA_PART;
  unsigned char* Wst = AT + 2 * 3 * TKB_A_PART;  // [2] stages of natural-layout weights (3 parts)
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wst + 2 * TKB_W_STAGE);
  uint64_t* w_full = bars;       // [2]
  uint64_t* w_empty = bars + 2;  // [2]
  uint64_t* a_ready = bars + 4;  // tiles written
  uint64_t* d_full = bars + 5;   // data-path accumulator ready
  uint64_t* w_done = bars + 6;   // every MMA of the step has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = g.n, l = g.l, K = g.K, KB = g.KB, pi = g.pi;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(a_ready, 256); mbar_init(d_full, 1); mbar_init(w_done, 1);
    mbar_fence_init();
  }
  if (warp == 8) tc_alloc(tmem_slot, 512);
  // everything the tensor core may read as operand rows must be finite: clear the A buffers and the weight stages once
  for (int e = tid; e < (int)((2 * 3 * TKB_A_PART + 2 * TKB_W_STAGE) / 16); e += TKB_THREADS)
    reinterpret_cast<uint4*>(AT)[e] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: [0,64) data path | [64, 64+KB) bottleneck gradient | then 64 per layer m = l .. 0
  const uint32_t col_b = 64u, col_m0 = 64u + (uint32_t)KB;  // layer m at col_m0 + (l - m) * 64

  long long my_tiles = 0;
  if ((long long)blockIdx.x < a.total_tiles) my_tiles = (a.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int steps = l + 2;  // per tile: bottleneck, layers l..1, layer 0

  if (warp == 9) {
    if (lane == 0) {  // natural-layout weights, one stage per step that has a data product (bottleneck, layers l..1)
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.packed) + g.f_bytes;
      uint32_t s = 0, ph = 0;
      for (long long t = 0; t < my_tiles; ++t)
        for (int st = 0; st <= l; ++st) {
          mbar_wait(&w_empty[s], ph ^ 1u);
          const uint32_t nb = st == 0 ? (uint32_t)(3 * KB * 128) : 24576u;
          const long long off = st == 0 ? 0 : 3LL * KB * 128 + (long long)(st - 1) * 24576;
          mbar_expect_tx(&w_full[s], nb);
          bulk_g2s(Wst + s * TKB_W_STAGE, src + off, nb, &w_full[s]);
          if (++s == 2) { s = 0; ph ^= 1u; }
        }
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint64_t dA_k = tk_desc(smem_u32(DA), 2048, 128);    // da tile as K-major A (M = b, K = j)
    const uint64_t dB_mn = tk_desc(smem_u32(DA), 128, 2048);   // da tile as MN-major B (N = j, K = b)
    uint32_t s = 0, wph = 0, aph = 0, gs = 0;
    for (long long t = 0; t < my_tiles; ++t) {
      const bool fresh = (t % a.flush_tiles) == 0;
      for (int st = 0; st < steps; ++st, aph ^= 1u, ++gs) {
        mbar_wait(a_ready, aph);
        const bool has_data = st <= l;
        if (has_data) mbar_wait(&w_full[s], wph);
        tc_fence_after();
        const uint64_t dAT = tk_desc(smem_u32(AT + (gs & 1u) * 3 * TKB_A_PART), 128, 2048);  // MN-major A (M = mn, K = b)
        if (tc_elect_one()) {
          if (has_data) {
            // dh (+)= da @ W^T : A = da tile (K-major view), B = natural weights [64 rows i x K extent]
            const bool bot = st == 0;
            const uint64_t dW = tk_desc(smem_u32(Wst + s * TKB_W_STAGE), 128, bot ? (uint32_t)(KB / 8) * 128u : 1024u);
            bf3_mma(tmem_u, dA_k, TKB_DA_PART / 16, 256u, dW, bot ? (uint32_t)(KB * 128 / 16) : 512u, 16u, bf_idesc(64),
                    bot ? KB / 16 : 4, true);
            tc_commit(d_full);
            tc_commit(&w_empty[s]);
          }
          // batch reduction: D[(mn)][j] += AT[mn][b] * da[j][b], K = 128 rows
          const uint32_t dcol = st == 0 ? col_b : col_m0 + (uint32_t)(st - 1) * 64u;
          bf3_mma(tmem_u + dcol, dAT, TKB_A_PART / 16, 16u, dB_mn, TKB_DA_PART / 16, 16u, bf_idesc(st == 0 ? KB : 64, 1, 1), 8,
                  fresh);
          tc_commit(w_done);
        }
        __syncwarp();
        if (has_data && ++s == 2) { s = 0; wph ^= 1u; }
      }
    }
  } else {
    // ---------------- epilogue warps ----------------
    const int r = tid & 127, half = tid >> 7;
    const uint32_t tm = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t koff = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;  // this row's position along k = b
    const long long slot_floats = nif_tiled_rows(a.B) * 64;
    uint32_t dph = 0, wdph = 0, gs = 0;  // gs: global step counter (selects the A buffer)
    bool first_step = true;  // no MMA is in flight: nothing to wait for before writing
    auto wait_mmas = [&]() {  // every MMA issued so far has completed
      if (!first_step) { mbar_wait(w_done, wdph); wdph ^= 1u; }
      first_step = false;
    };
    // 8 values -> group `grp` (8 consecutive mn) of a 3-part tile at `base`, parts `pstride` bytes apart
    auto put8 = [&](unsigned char* base, uint32_t pstride, int grp, const float (&w8)[8]) {
      uint4 q0, q1, q2;
      bf3_split8(w8, q0, q1, q2);
      *reinterpret_cast<uint4*>(base + grp * 2048 + koff) = q0;
      *reinterpret_cast<uint4*>(base + pstride + grp * 2048 + koff) = q1;
      *reinterpret_cast<uint4*>(base + 2 * pstride + grp * 2048 + koff) = q2;
    };
    // this thread's half of the stashed row of slot `slot`: 8 quads -> 32 floats
    auto load_half = [&](int slot, long long b, bool live, float (&v)[32]) {
      const float* src = a.save + (long long)slot * slot_floats + nif_tiled_row(b) + (long long)(half * 8) * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) q = ldg4(src + c * 128);
        v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
      }
    };
    // rows mn = 32 half .. of A buffer `buf` := v (this thread's 32 columns of h)
    auto put_h = [&](uint32_t buf, const float (&v)[32]) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float w8[8] = {v[8 * c], v[8 * c + 1], v[8 * c + 2], v[8 * c + 3], v[8 * c + 4], v[8 * c + 5], v[8 * c + 6], v[8 * c + 7]};
        put8(AT + buf * 3 * TKB_A_PART, TKB_A_PART, half * 4 + c, w8);
      }
    };
    long long flush_idx = 0;
    for (long long t = 0; t < my_tiles; ++t) {
      const long long b = (blockIdx.x + t * gridDim.x) * 128 + r;
      const bool live = b < a.B;
      // ---- step 0: da tile := dz (KB columns), A buffer := [h_l | 1 | p_in]; the [1 | p_in] rows go into both buffers ----
      float hv[32];
      load_half(l, b, live, hv);
      float ev[8];  // half 0: rows 64..71 = [1, p_0 .. p_6];  half 1: rows 72..79 = [p_7, 0 ..]
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = half * 8 + e - 1;  // p_in column
        ev[e] = !live ? 0.f : (half == 0 && e == 0 ? 1.f : ((i >= 0 && i < pi) ? __ldg(&a.p_in[b * pi + i]) : 0.f));
      }
      const int ngz = KB / 16;  // groups of 8 dz columns per half
      float dzv[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int kk = half * 8 * ngz + e;
        dzv[e] = (live && e < 8 * ngz && kk < K) ? __ldg(&a.dz[b * K + kk]) : 0.f;
      }
      wait_mmas();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= ngz) break;
        const float w8[8] = {dzv[8 * c], dzv[8 * c + 1], dzv[8 * c + 2], dzv[8 * c + 3], dzv[8 * c + 4], dzv[8 * c + 5], dzv[8 * c + 6], dzv[8 * c + 7]};
        put8(DA, TKB_DA_PART, half * ngz + c, w8);
      }
      put_h(gs & 1u, hv);
      put8(AT, TKB_A_PART, 8 + half, ev);
      put8(AT + 3 * TKB_A_PART, TKB_A_PART, 8 + half, ev);
      fence_async_smem();
      mbar_arrive(a_ready);
      ++gs;

      float dh[32];
      // ---- steps 1 .. l+1: layer m = l .. 0 ----
#pragma unroll 1
      for (int m = l; m >= 0; --m, ++gs) {
        // this step's stash rows, fetched before any wait: act'(pre_m) and (m >= 1) the layer input h_{m-1}
        float dv[32];
        load_half(l + 1 + m, b, live, dv);
        if (m >= 1) {
          load_half(m - 1, b, live, hv);
          put_h(gs & 1u, hv);  // the other A buffer: its last reader (two steps back) finished before the previous da write
        }
        // dh_m: the data-path accumulator of the previous step (shortcut: dh_m = dh_{m+1} + da_{m+1} W_{m+1}^T)
        mbar_wait(d_full, dph);
        dph ^= 1u;
        tc_fence_after();
        {
          float v[32];
          tc_ld32(tm + (uint32_t)(half * 32), v);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) dh[e] = (m == l) ? v[e] : dh[e] + v[e];
        }
        tc_fence_before();
#pragma unroll
        for (int e = 0; e < 32; ++e) dv[e] *= dh[e];  // da_m
        wait_mmas();  // the da tile is free
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float w8[8] = {dv[8 * c], dv[8 * c + 1], dv[8 * c + 2], dv[8 * c + 3], dv[8 * c + 4], dv[8 * c + 5], dv[8 * c + 6], dv[8 * c + 7]};
          put8(DA, TKB_DA_PART, half * 4 + c, w8);
        }
        fence_async_smem();
        mbar_arrive(a_ready);
      }

      // ---- end of an accumulation chain: accumulators -> this CTA's partial (each half drains its columns) ----
      if ((t + 1) % a.flush_tiles == 0 || t + 1 == my_tiles) {
        mbar_wait(w_done, wdph);
        wdph ^= 1u;
        first_step = true;
        tc_fence_after();
        float* part = a.part + ((long long)blockIdx.x * a.nflush + flush_idx) * g.P;
        ++flush_idx;
        // bottleneck: rows i < n -> W_b[i][kk], row 64 -> b_b[kk]
        for (int c0 = half * (KB / 2); c0 < (half + 1) * (KB / 2); c0 += 8) {
          float v[8];
          tc_ld8(tm + col_b + (uint32_t)c0, v);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int kk = c0 + e;
            if (kk < K) {
              if (r < n) part[g.off_Wb + (long long)r * K + kk] = v[e];
              else if (r == 64) part[g.off_bb + kk] = v[e];
            }
          }
        }
        for (int m = l; m >= 0; --m) {
          const uint32_t col = col_m0 + (uint32_t)(l - m) * 64u + (uint32_t)(half * 32);
          float v[32];
          tc_ld32(tm + col, v);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int j = half * 32 + e;
            if (j < n) {
              if (r < n && m >= 1) part[g.off_Wh + (long long)(m - 1) * n * n + (long long)r * n + j] = v[e];
              else if (r == 64) part[(m == 0 ? g.off_b0 : g.off_bh + (long long)(m - 1) * n) + j] = v[e];
              else if (m == 0 && r > 64 && r <= 64 + pi) part[(long long)(r - 65) * n + j] = v[e];
            }
          }
        }
        tc_fence_before();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 512);
}

// g_theta[e] = beta * g_theta[e] + sum over partials
__global__ void __launch_bounds__(256) nif_trunk_tc_reduce_kernel(long long P, int nparts, const float* __restrict__ part,
                                                                  float* __restrict__ g_theta, float beta) {
  const long long e = blockIdx.x * 256LL + threadIdx.x;
  if (e >= P) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(long long)p * P + e];
  g_theta[e] = beta != 0.f ? fmaf(beta, g_theta[e], s) : s;
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static int trunk_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
static const int kFlushTiles = 32;  // 4096 rows per TMEM accumulation chain (see nif_desc_t.acc_rows)

// does this trunk run on the tensor-core kernels?
bool nif_trunk_uses_tc(int pi, int K, int n, int l, int act) {
  if (pi < 1 || pi > NIF_MAX_SI || K < 1 || K > 64 || n < 1 || n > 64 || l < 1 || act == NIF_ACT_SINE) return false;
  const TrunkGeo g = trunk_geo(pi, K, n, l, act);
  if (tkf_smem_bytes(g) > 227 * 1024) return false;
  return 64 + g.KB + (l + 1) * 64 <= 512;  // TMEM columns of the reverse kernel
}
long long nif_trunk_tc_packed_floats(int pi, int K, int n, int l, int act) {
  const TrunkGeo g = trunk_geo(pi, K, n, l, act);
  return (g.f_bytes + g.r_bytes) / 4;
}
long long nif_trunk_tc_ws_floats(int pi, int K, int n, int l, int act, long long B) {
  const TrunkGeo g = trunk_geo(pi, K, n, l, act);
  const long long tiles = (B + 127) / 128;
  long long grid = trunk_sms();
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  const long long per_cta = (tiles + grid - 1) / grid;
  const long long nflush = (per_cta + kFlushTiles - 1) / kFlushTiles;
  return grid * (nflush < 1 ? 1 : nflush) * g.P;
}

int nif_trunk_tc_forward_impl(int pi, int K, int n, int l, int act, long long B, const float* p_in, const float* theta,
                              float* z, float* save, float* packed, cudaStream_t st) {
  const TrunkGeo g = trunk_geo(pi, K, n, l, act);
  {
    const long long slots = (g.f_bytes + g.r_bytes) / 4;
    NIF_PROF("nif_trunk_tc_pack_kernel", st);
    nif_trunk_tc_pack_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, st>>>(g, theta, packed);
  }
  NIF_CUDA_CHECK(cudaGetLastError());
  TrunkFwdArgs a;
  a.B = B; a.total_pairs = (B + 255) / 256;
  a.p_in = p_in; a.theta = theta; a.packed = packed; a.z = z; a.save = save;
  const size_t smem = tkf_smem_bytes(g);
  void (*kern)(const TrunkGeo, const TrunkFwdArgs) = nullptr;
  switch (act) {
    case NIF_ACT_SWISH: kern = nif_trunk_tc_fwd_kernel<NIF_ACT_SWISH>; break;
    case NIF_ACT_SIGMOID: kern = nif_trunk_tc_fwd_kernel<NIF_ACT_SIGMOID>; break;
    case NIF_ACT_TANH: kern = nif_trunk_tc_fwd_kernel<NIF_ACT_TANH>; break;
    case NIF_ACT_RELU: kern = nif_trunk_tc_fwd_kernel<NIF_ACT_RELU>; break;
    default: kern = nif_trunk_tc_fwd_kernel<NIF_ACT_LINEAR>; break;
  }
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = trunk_sms();
  if (grid > a.total_pairs) grid = a.total_pairs;
  if (grid < 1) return NIF_OK;
  { NIF_PROF("nif_trunk_tc_fwd_kernel", st); kern<<<(unsigned)grid, TKF_THREADS, smem, st>>>(g, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

int nif_trunk_tc_backward_impl(int pi, int K, int n, int l, int act, long long B, const float* p_in, const float* save,
                               const float* dz, float* g_theta, float beta, const float* packed, float* ws,
                               cudaStream_t st) {
  const TrunkGeo g = trunk_geo(pi, K, n, l, act);
  TrunkBwdArgs a;
  a.B = B; a.total_tiles = (B + 127) / 128;
  a.p_in = p_in; a.packed = packed; a.save = save; a.dz = dz; a.part = ws;
  long long grid = trunk_sms();
  if (grid > a.total_tiles) grid = a.total_tiles;
  if (grid < 1) return NIF_OK;
  const long long per_cta = (a.total_tiles + grid - 1) / grid;
  a.flush_tiles = kFlushTiles;
  a.nflush = (int)((per_cta + kFlushTiles - 1) / kFlushTiles);
  // CTAs with fewer tiles leave some partial sets unwritten: clear the workspace they would have used
  NIF_CUDA_CHECK(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)(grid * a.nflush * g.P), st));
  const size_t smem = tkb_smem_bytes();
  NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_trunk_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { NIF_PROF("nif_trunk_tc_bwd_kernel", st); nif_trunk_tc_bwd_kernel<<<(unsigned)grid, TKB_THREADS, smem, st>>>(g, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  {
    NIF_PROF("nif_trunk_tc_reduce_kernel", st);
    nif_trunk_tc_reduce_kernel<<<(unsigned)((g.P + 255) / 256), 256, 0, st>>>(g.P, (int)(grid * a.nflush), ws, g_theta, beta);
  }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
