// Thin reverse-pass terms on the tensor cores (tcgen05 + TMEM, FP16x3): the bias rows of every layer and the first and
// last matrix of the ShapeNet,
//   out[kappa][q] = sum_b zt[b][kappa] * F[b][q],
// a batch-reduction GEMM with a very short M' = K+1 side.  It is the tensor-core twin of nif_bwd_edge_kernel and of
// nif_tc_bwd_weight_kernel: the feature columns sit on the M = 128 side (two 64-column blocks per CTA), the latent
// coordinates on the N side (KZ <= 64), the batch index is the MMA K dimension; both operands are generated on the
// fly by the row-owning threads (MN-major), with global power-of-two scales from the maxima the earlier kernels of
// the reverse pass recorded.
//   blocks (64 columns each):  m = 0..H     F = da_m[b][j]                       -> dC_m[kappa][j]
//                              i < si       F = omega x[b][i] da_0[b][j]         -> dM_0[kappa][i][j]
//                              c < so       F = du[b][c] h_{H+1}[b][i]           -> dM_last[kappa][i][c]
//                              one block    F = du[b][c], c < so                 -> dC_last[kappa][c]
// Sources are in the tiled activation layout (nif_common.cuh), so a warp of generators reads 512 contiguous bytes
// per load.  Partials go to part[S][K+1][Q] in the column order nif_unpack_grad_kernel sums.
#include "nif_tc.cuh"

struct TcEdgeArgs {
  long long B, rows_per_split;
  int S, Q;
  const float *z, *x, *save, *da, *du;
  const unsigned* maxes;  // see nif_tc_bwd_data_kernel / nif_dz_edge_kernel
  float* part;
  int no_bias;  // tangent-adjoint pass (save = the tangent activations, x = xdot): the bias-row columns are zero
};

#define TCE_THREADS 288
#define TCE_A_BYTES 16384u  // [128 (2 blocks x 64 columns) x 64 (b)] fp16, MN-major
#define TCE_B_BYTES 8192u   // [KZ <= 64 (kappa) x 64 (b)] fp16, MN-major
#define TCE_SLOT_BYTES (2 * TCE_A_BYTES + 2 * TCE_B_BYTES)

// slots of the maxima buffer written by nif_dz_edge_kernel (nif_tc_bwd.cu)
#define NIF_MAX_X 192    // + i : max |x[:, i]|
#define NIF_MAX_DU 200   // + c : max |du[:, c]|
#define NIF_MAX_HL 208   //       max |h_{H+1}|

__device__ __forceinline__ uint64_t tce_make_desc(uint32_t saddr) {  // MN-major, 8 (k) x 16 B core matrices
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;   // LBO: next group of 8 k
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;  // SBO: next group of 8 mn
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tce_split8(const float (&v)[8], float sc, uint4& hi, uint4& lo) {
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

__global__ void __launch_bounds__(TCE_THREADS, 1) nif_tc_bwd_edge_kernel(const Plan pl, const TcEdgeArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t slot_full[2], slot_empty[2], done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = pl.K, K1 = pl.K + 1, H = pl.H, si = pl.si, so = pl.so, KZ = pl.KZ;
  const int nblk = H + 1 + si + so + 1;
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
  if (warp == 8) tc_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const long long slot_floats = nif_tiled_rows(a.B) * 64;

  // description of column block `blk`: source slot, per-row multiplier, scale bound, destination columns
  struct Blk { const float* src; const float* mul; int mul_stride; float cst; float bound; int q0, qs, ncol; bool on; };
  auto describe = [&](int blk) {
    Blk d;
    d.on = blk < nblk;
    d.src = a.da; d.mul = nullptr; d.mul_stride = 0; d.cst = 1.f; d.bound = 0.f; d.q0 = 0; d.qs = 1; d.ncol = pl.n;
    if (!d.on) return d;
    if (blk <= H) {
      d.src = a.da + (long long)blk * slot_floats;
      d.bound = __uint_as_float(a.maxes[blk]);
      d.q0 = blk * 64;
      if (a.no_bias) d.cst = 0.f;
    } else if (blk < H + 1 + si) {
      const int i = blk - H - 1;
      d.src = a.da; d.mul = a.x + i; d.mul_stride = si; d.cst = plan_omega(pl, 0);
      d.bound = d.cst * __uint_as_float(a.maxes[NIF_MAX_X + i]) * __uint_as_float(a.maxes[0]);
      d.q0 = (H + 1) * 64 + so + i * 64;
    } else if (blk < H + 1 + si + so) {
      const int c = blk - H - 1 - si;
      d.src = a.save + (long long)H * slot_floats; d.mul = a.du + c; d.mul_stride = so;
      d.bound = __uint_as_float(a.maxes[NIF_MAX_DU + c]) * __uint_as_float(a.maxes[NIF_MAX_HL]);
      d.q0 = (H + 1) * 64 + so + si * 64 + c; d.qs = so;
    } else {  // last bias row: the seed itself, `so` columns straight from du (row-major)
      d.src = nullptr;
      for (int c = 0; c < so; ++c) d.bound = fmaxf(d.bound, __uint_as_float(a.maxes[NIF_MAX_DU + c]));
      d.q0 = (H + 1) * 64; d.ncol = so;
      if (a.no_bias) d.cst = 0.f;
    }
    return d;
  };
  float scB, invB;
  tc_row_scale(__uint_as_float(a.maxes[H + 1]), scB, invB);

  if (warp == 8) {
    {  // MMA issuer: the whole warp runs the loop (uniform operands), one elected lane issues -- see tc_elect_one
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(KZ >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      for (long long t = 0; t < nsub; ++t) {
        const int sl = (int)(t & 1);
        mbar_wait(&slot_full[sl], (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        if (tc_elect_one()) {
        const uint32_t base = smem_u32(smem + sl * TCE_SLOT_BYTES);
        const uint64_t a_hi = tce_make_desc(base), a_lo = tce_make_desc(base + TCE_A_BYTES);
        const uint64_t b_hi = tce_make_desc(base + 2 * TCE_A_BYTES), b_lo = tce_make_desc(base + 2 * TCE_A_BYTES + TCE_B_BYTES);
        const uint32_t d1 = tmem_u, d2 = tmem_u + 64u;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // 16 rows (k) per instruction
          const uint64_t adv = (uint64_t)(ks * 16);
          const uint32_t accf = (t > 0 || ks > 0) ? 1u : 0u;
          // one N = 128 instruction over the adjacent [B_hi | B_lo] tiles: [d1 | d2] += a_hi x [b_hi | b_lo] (columns
          // KZ..63 of either half are never read)
          tc_mma_f16(d1, a_hi + adv, b_hi + adv, idesc2, accf);
          tc_mma_f16(d2, a_lo + adv, b_hi + adv, idesc, 1u);
        }
        tc_commit(&slot_empty[sl]);
        }
        __syncwarp();
      }
      if (tc_elect_one()) tc_commit(&done_bar);
      __syncwarp();
    }
  } else {
    // ---------------- operand generators: slot sl = warp / 4, thread (q, r) = block q of the pair, row r ----------------
    const int sl = warp >> 2;
    const int q = (tid >> 6) & 1, r = tid & 63;
    unsigned char* slot = smem + sl * TCE_SLOT_BYTES;
    unsigned char* A_hi = slot;
    unsigned char* A_lo = slot + TCE_A_BYTES;
    unsigned char* B_hi = slot + 2 * TCE_A_BYTES;
    unsigned char* B_lo = B_hi + TCE_B_BYTES;
    const uint32_t koff = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;  // this row's position along k
    const Blk d = describe(2 * blockIdx.x + q);
    float scA, invA;
    tc_row_scale(d.bound, scA, invA);
    // Row data of the NEXT sub-tile of this slot is requested as soon as the registers of the current one are free (features
    // after the A stores, latent coordinates after the B stores), so no L2 round trip sits between two sub-tiles: the loop
    // used to pay four of them in series (features, then each group of latent coordinates) per sub-tile.
    float4 fq[16];
    float mulv = 0.f;
    auto fetch_f = [&](long long t) {
      const long long b = r0 + t * 64 + r;
      const bool live = t < nsub && b < r1 && d.on;
      mulv = d.cst;
      if (live && d.mul) mulv *= __ldg(&d.mul[b * d.mul_stride]);
      if (d.src) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          fq[c] = live ? ldg4(d.src + nif_tiled_row(b) + c * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {  // the seed block: du[b][0..so)
#pragma unroll
        for (int c = 0; c < 16; ++c) fq[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        float dv[NIF_MAX_SO];
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c) dv[c] = (live && c < so) ? __ldg(&a.du[b * so + c]) : 0.f;
        fq[0] = make_float4(dv[0], dv[1], dv[2], dv[3]);
        fq[1] = make_float4(dv[4], dv[5], dv[6], dv[7]);
      }
    };
    // the zt operand: groups of 8 latent coordinates, even groups by q = 0, odd groups by q = 1 (KZ <= 64: at most 4 each)
    float zv[4][8];
    const bool zvec = (K & 3) == 0;  // rows of z are 16-byte aligned
    auto fetch_z = [&](long long t) {
      const long long b = r0 + t * 64 + r;
      const bool rowlive = t < nsub && b < r1;
      const float* zrow = a.z + (rowlive ? b : r0) * K;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int g = q + 2 * u;
#pragma unroll
        for (int e = 0; e < 8; ++e) zv[u][e] = 0.f;
        if (g < KZ / 8 && rowlive) {
          if (zvec && 8 * g + 8 <= K) {
            const float4 p0 = ldg4(zrow + 8 * g), p1 = ldg4(zrow + 8 * g + 4);
            zv[u][0] = p0.x; zv[u][1] = p0.y; zv[u][2] = p0.z; zv[u][3] = p0.w;
            zv[u][4] = p1.x; zv[u][5] = p1.y; zv[u][6] = p1.z; zv[u][7] = p1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int kk = 8 * g + e;
              zv[u][e] = (kk < K) ? __ldg(zrow + kk) : (kk == K ? 1.f : 0.f);
            }
          }
        }
      }
    };
    fetch_f(sl);
    fetch_z(sl);
    long long n_mine = 0;
    for (long long t = sl; t < nsub; t += 2, ++n_mine) {
      // wait until the MMAs that read this slot two sub-tiles ago have completed
      mbar_wait(&slot_empty[sl], (uint32_t)((n_mine & 1) ^ 1));
#pragma unroll
      for (int ig = 0; ig < 8; ++ig) {
        const float4 p0 = fq[2 * ig], p1 = fq[2 * ig + 1];
        const float fv[8] = {p0.x * mulv, p0.y * mulv, p0.z * mulv, p0.w * mulv, p1.x * mulv, p1.y * mulv, p1.z * mulv, p1.w * mulv};
        uint4 hi, lo;
        tce_split8(fv, scA, hi, lo);
        const uint32_t off = (uint32_t)(q * 8 + ig) * 1024u + koff;
        *reinterpret_cast<uint4*>(A_hi + off) = hi;
        *reinterpret_cast<uint4*>(A_lo + off) = lo;
      }
      fetch_f(t + 2);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int g = q + 2 * u;
        if (g < KZ / 8) {
          uint4 hi, lo;
          tce_split8(zv[u], scB, hi, lo);
          const uint32_t off = (uint32_t)g * 1024u + koff;
          *reinterpret_cast<uint4*>(B_hi + off) = hi;
          *reinterpret_cast<uint4*>(B_lo + off) = lo;
        }
      }
      fetch_z(t + 2);
      fence_async_smem();
      mbar_arrive(&slot_full[sl]);
    }
    // ---------------- final epilogue (warps 0-3): TMEM lane = column (block, j) ----------------
    if (warp < 4) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
      const int row = tid;  // 0..127
      const int blk = row >> 6, j = row & 63;
      const Blk de = describe(2 * blockIdx.x + blk);
      float scE, invE;
      tc_row_scale(de.bound, scE, invE);
      const float scale = invE * invB;
      const uint32_t tm = tmem + ((uint32_t)(warp * 32) << 16);
      const long long qd = de.q0 + (long long)j * de.qs;
      const bool wr = de.on && j < de.ncol;
      for (int c0 = 0; c0 < KZ; c0 += 16) {
        float v1[16], v2[16];
        tc_ld16(tm + (uint32_t)c0, v1);
        tc_ld16(tm + (uint32_t)(64 + c0), v2);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int kk = c0 + e;
          if (wr && kk < K1) a.part[((long long)s * K1 + kk) * a.Q + qd] = nsub ? scale * (v1[e] + v2[e]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tc_dealloc(tmem, 128);
}

// Writes every column of part_e for S batch splits of rows_per_split rows (a multiple of 64).
int nif_tc_bwd_edge_impl(const Plan& pl, long long B, const float* z, const float* x, const float* save, const float* da,
                         const float* du, const unsigned* maxes, int S, long long rows_per_split, int Q, float* part,
                         cudaStream_t st, int no_bias = 0) {
  if (!nif_plan_uses_tc(pl) || pl.KZ > 64) return NIF_E_UNSUPPORTED;
  TcEdgeArgs a;
  a.no_bias = no_bias;
  a.B = B; a.rows_per_split = rows_per_split; a.S = S; a.Q = Q;
  a.z = z; a.x = x; a.save = save; a.da = da; a.du = du; a.maxes = maxes; a.part = part;
  const size_t smem = 2 * (size_t)TCE_SLOT_BYTES;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_tc_bwd_edge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblk = pl.H + 1 + pl.si + pl.so + 1;
  dim3 grid((unsigned)((nblk + 1) / 2), (unsigned)S);
  { NIF_PROF("nif_tc_bwd_edge_kernel", st); nif_tc_bwd_edge_kernel<<<grid, TCE_THREADS, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
