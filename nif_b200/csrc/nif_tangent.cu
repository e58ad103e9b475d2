// Forward-mode tangents through the fused path: the B200 realisation of JacobianLayer
// (nif/layers/gradient.py:36-49, 207-231), which in the reference costs one reverse pass per output.
//
// For a direction d with latent tangent zd (from the ParameterNet trunk, zero for a pure coordinate
// direction) and coordinate tangent xd:
//   pre_m    = sum_k zt[k] (omega h M_m[k] + C_m[k])
//   pre_m'   = sum_k zt[k] omega (h' M_m[k])  +  sum_k zd[k] (omega h M_m[k] + C_m[k])
//   h_{m+1}  = alpha act(pre_m) (+ residual),   h_{m+1}' = alpha act'(pre_m) pre_m' (+ residual')
// The weight stream is shared: every staged chunk M_m[k] is used by the primal GEMM and by one GEMM
// per direction, and the primal product is reused for the zd term.
//
// Second order (HessianLayer, nif/layers/gradient.py:130-180, 234-261), template flag SEC: four streams
// h, h_a', h_b', h_ab'' through the same weight stream,
//   pre_m'' = sum_k zt[k] omega (h_ab'' M_m[k]) + zd_a[k] omega (h_b' M_m[k]) + zd_b[k] omega (h_a' M_m[k])
//             + zdd_ab[k] (omega h M_m[k] + C_m[k])
//   h_{m+1}'' = alpha ( act''(pre_m) pre_a' pre_b' + act'(pre_m) pre_m'' ) (+ residual'')
// one launch per pair of directions (a, b); a == b gives the diagonal.
#include "nif_tile.cuh"

struct TanArgs {
  long long B, total_tiles;
  const float *z, *x, *packed, *zdot, *xdot;  // zdot [ND][B][K], xdot [ND][B][si] (either may be null)
  float *u, *udot;                            // udot [ND][B][so]
  // second-order mode (SEC, HessianLayer): streams 1, 2 are the directions a, b given by zdot / xdot [2][..], stream 3
  // carries the mixed second derivative; zddot [B][K] is the second derivative of the latent code along (a, b)
  // (null = 0; coordinates have none), uddot [B][so] receives d2u / da db
  const float* zddot;
  float* uddot;
  // optional stash for the reverse-over-forward pass (Sobolev training), slots of [B][NP]:
  //   [0, H]            h_{m+1}                       [H+1, 2H+1]     d_m = alpha act'(pre_m)
  //   [2H+2, 3H+2]      h'_{m+1} of direction 0       [3H+3, 4H+3]    e_m = alpha act''(pre_m) pre_m' of direction 0
  float* save;    // slots [0, 2(H+1)): h_{m+1}, d_m
  float* save_t;  // tangent slots of this launch's first direction: per direction (H+1) h'_{m+1}, (H+1) e_m
};

using TCfg32 = TileCfg<32, 128, 1, 1>;
using TCfg64 = TileCfg<64, 64, 1, 1>;
using TCfg128 = TileCfg<128, 64, 1, 2>;

template <class C, int ND>
__host__ __device__ inline size_t tan_smem_bytes(int K, int si, int so) {
  size_t f = 2 * (size_t)C::STAGE_FLOATS + (size_t)(1 + ND) * C::NP * C::TB + (size_t)(1 + ND) * (K + 1) * C::TB +
             (size_t)(1 + ND) * si * C::TB + (size_t)(1 + ND) * so * C::NT;
  return f * 4 + 64;
}

template <class C, int ND, bool RES, bool SEC>
__global__ void __launch_bounds__(C::NT, 1) nif_tangent_kernel(const Plan pl, const TanArgs a) {
  constexpr int NP = C::NP, TB = C::TB, MP = C::MP, MJ = C::MJ, NT = C::NT, NS = 1 + ND;
  static_assert(!SEC || ND == 3, "second-order mode carries the streams h, h_a', h_b', h_ab''");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* act = stage + 2 * C::STAGE_FLOATS;          // [NS][NP][TB]  stream 0 = primal
  float* zs = act + NS * NP * TB;                    // [NS][K+1][TB]
  float* xs = zs + NS * (pl.K + 1) * TB;             // [NS][si][TB]
  float* ys = xs + NS * pl.si * TB;                  // [NS][so][NT] scratch of the last layer
  uint64_t* bar = reinterpret_cast<uint64_t*>(ys + NS * pl.so * NT);
  bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bar) + 7) & ~uintptr_t(7));

  const int tid = threadIdx.x;
  const int tj = tid % C::TY, tp = tid / C::TY;
  const int K = pl.K, K1 = pl.K + 1, H = pl.H, n = pl.n, si = pl.si, so = pl.so;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  long long my_tiles = 0;
  if ((long long)blockIdx.x < a.total_tiles) my_tiles = (a.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  WeightStream<C> ws;
  ws.stage = stage;
  ws.bar = bar;
  ws.chunks_per_tile = H * K1 * C::NH;
  ws.total = my_tiles * ws.chunks_per_tile;
  ws.issued = 0;
  ws.consumed = 0;
  ws.H = H;
  ws.K1 = K1;
  ws.packed = a.packed;
  ws.packed_floats = pl.packed_floats;
  ws.sec_off = pl.off_MH;
  ws.tiles_per_group = a.total_tiles;
  ws.reverse = false;
  if (tid == 0) {
    while (ws.issued < 2 && ws.issued < ws.total) ws.issue_one();
  }
  const float* C_all = a.packed + pl.off_C;

  auto load_rows4 = [&](const float* base, float (&v)[MP]) {  // MP consecutive-group rows from a [..][TB] row
#pragma unroll
    for (int gp = 0; gp < C::GP; ++gp) {
      const float4 q = *reinterpret_cast<const float4*>(&base[gp * C::PSTR + tp * 4]);
      v[gp * 4] = q.x; v[gp * 4 + 1] = q.y; v[gp * 4 + 2] = q.z; v[gp * 4 + 3] = q.w;
    }
  };

  for (long long t = 0; t < my_tiles; ++t) {
    const long long tile = blockIdx.x + t * gridDim.x;
    const long long row0 = tile * TB;
    // ---- stage z, zdot, x, xdot --------------------------------------------------------------------
    for (int s = 0; s < NS; ++s) {
      const float* zsrc = (s == 0) ? a.z : (a.zdot ? a.zdot + (long long)(s - 1) * a.B * K : nullptr);
      if (SEC && s == 3) zsrc = a.zddot;
      for (int idx = tid; idx < TB * K; idx += NT) {
        const int p = idx / K, kk = idx - p * K;
        const long long b = row0 + p;
        zs[(s * K1 + kk) * TB + p] = (zsrc && b < a.B) ? __ldg(&zsrc[b * K + kk]) : 0.f;
      }
      for (int p = tid; p < TB; p += NT) zs[(s * K1 + K) * TB + p] = (s == 0) ? 1.f : 0.f;
      const float* xsrc = (s == 0) ? a.x : (a.xdot ? a.xdot + (long long)(s - 1) * a.B * si : nullptr);
      if (SEC && s == 3) xsrc = nullptr;  // coordinates have no second derivative
      for (int idx = tid; idx < TB * si; idx += NT) {
        const int p = idx / si, i = idx - p * si;
        const long long b = row0 + p;
        xs[(s * si + i) * TB + p] = (xsrc && b < a.B) ? __ldg(&xsrc[b * si + i]) : 0.f;
      }
    }
    __syncthreads();

    float acc[NS][MP][MJ];
    float carry[RES ? NS : 1][RES ? MP : 1][RES ? MJ : 1];

    auto zero_acc = [&]() {
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int r = 0; r < MP; ++r)
#pragma unroll
          for (int c = 0; c < MJ; ++c) acc[s][r][c] = 0.f;
    };

    // fold one partial product into the accumulators.  stream 0 product (from h): goes to acc[0] scaled
    // by zt and to acc[d] scaled by zd_d; stream d product (from h'_d): goes to acc[d] scaled by zt.
    auto fold = [&](int s, const float (&tmp)[MP][MJ], int m, int kk, float om, bool with_bias) {
      float cv[MJ];
#pragma unroll
      for (int gj = 0; gj < C::GJ; ++gj) {
        float4 cb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (with_bias && s == 0) cb = ldg4(&C_all[((long long)m * K1 + kk) * NP + gj * C::JSTR + tj * 4]);
        cv[gj * 4] = cb.x; cv[gj * 4 + 1] = cb.y; cv[gj * 4 + 2] = cb.z; cv[gj * 4 + 3] = cb.w;
      }
      float zv[MP];
      load_rows4(&zs[(0 * K1 + kk) * TB], zv);
#pragma unroll
      for (int r = 0; r < MP; ++r)
#pragma unroll
        for (int c = 0; c < MJ; ++c) acc[s][r][c] = fmaf(zv[r], fmaf(om, tmp[r][c], cv[c]), acc[s][r][c]);
      if (s == 0) {
#pragma unroll
        for (int d = 1; d < NS; ++d) {
          float zd[MP];
          load_rows4(&zs[(d * K1 + kk) * TB], zd);
#pragma unroll
          for (int r = 0; r < MP; ++r)
#pragma unroll
            for (int c = 0; c < MJ; ++c) acc[d][r][c] = fmaf(zd[r], fmaf(om, tmp[r][c], cv[c]), acc[d][r][c]);
        }
      }
      if (SEC && (s == 1 || s == 2)) {  // cross terms: the product of h_a' meets zd_b and vice versa
        float zo[MP];
        load_rows4(&zs[((3 - s) * K1 + kk) * TB], zo);
#pragma unroll
        for (int r = 0; r < MP; ++r) {
          const float w = zo[r] * om;
#pragma unroll
          for (int c = 0; c < MJ; ++c) acc[SEC ? 3 : 0][r][c] = fmaf(w, tmp[r][c], acc[SEC ? 3 : 0][r][c]);
        }
      }
    };

    auto epilogue = [&](int m) {
      const float alpha = plan_alpha(pl, m);
      const int res = plan_res(pl, m);
      float outv[NS][MP][MJ];
#pragma unroll
      for (int r = 0; r < MP; ++r)
#pragma unroll
        for (int c4 = 0; c4 < MJ; c4 += 4) {  // outv[0] / outv[1] hold f / act' until the loop below consumes them
          const float v4[4] = {acc[0][r][c4], acc[0][r][c4 + 1], acc[0][r][c4 + 2], acc[0][r][c4 + 3]};
          float f4[4], d4[4];
          act_fd4(pl.act, v4, f4, d4);
#pragma unroll
          for (int e = 0; e < 4; ++e) { outv[0][r][c4 + e] = f4[e]; outv[1][r][c4 + e] = d4[e]; }
        }
      if (a.save) {  // d_m and e_m of every direction (before the residual bookkeeping below overwrites outv)
        float* sd = a.save + (long long)(H + 1 + m) * a.B * NP;
#pragma unroll
        for (int r = 0; r < MP; ++r) {
          const long long b = row0 + row_of<C>(tp, r);
          if (b < a.B) {
#pragma unroll
            for (int gj = 0; gj < C::GJ; ++gj) {
              float dq[4], dd[4];
#pragma unroll
              for (int f = 0; f < 4; ++f) {
                const int c = gj * 4 + f;
                const bool livec = col_of<C>(tj, c) < n;
                dq[f] = livec ? alpha * outv[1][r][c] : 0.f;
                dd[f] = livec ? alpha * act_dd(pl.act, acc[0][r][c], outv[0][r][c], outv[1][r][c]) : 0.f;
              }
              const int j0 = gj * C::JSTR + tj * 4;
              *reinterpret_cast<float4*>(&sd[b * NP + j0]) = make_float4(dq[0], dq[1], dq[2], dq[3]);
#pragma unroll
              for (int s = 1; s < (SEC ? 2 : NS); ++s) {
                float* se = a.save_t + (long long)((2 * s - 1) * (H + 1) + m) * a.B * NP;
                *reinterpret_cast<float4*>(&se[b * NP + j0]) = make_float4(dd[0] * acc[s][r][gj * 4], dd[1] * acc[s][r][gj * 4 + 1],
                                                                            dd[2] * acc[s][r][gj * 4 + 2], dd[3] * acc[s][r][gj * 4 + 3]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < MP; ++r)
#pragma unroll
        for (int c = 0; c < MJ; ++c) {
          const float f = outv[0][r][c], d = outv[1][r][c];
          const int j = col_of<C>(tj, c);
          const int ai = act_idx<C>(j, row_of<C>(tp, r));
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            float o = (s == 0) ? alpha * f : alpha * d * acc[s][r][c];
            if (SEC && s == 3)
              o = alpha * fmaf(act_dd(pl.act, acc[0][r][c], f, d), acc[1][r][c] * acc[SEC ? 2 : 0][r][c], d * acc[s][r][c]);
            if (res == 1) o += act[s * NP * TB + ai];
            if (RES) {
              if (res == 2) carry[RES ? s : 0][RES ? r : 0][RES ? c : 0] = act[s * NP * TB + ai];
              if (res == 3) o += 0.5f * carry[RES ? s : 0][RES ? r : 0][RES ? c : 0];
            }
            if (j >= n) o = 0.f;
            outv[s][r][c] = o;
          }
        }
      if (a.save) {  // h_{m+1} and h'_{m+1} of every direction
#pragma unroll
        for (int s = 0; s < (SEC ? 2 : NS); ++s) {
          float* sh = (s == 0 ? a.save : a.save_t + (long long)(2 * (s - 1) * (H + 1)) * a.B * NP) + (long long)m * a.B * NP;
#pragma unroll
          for (int r = 0; r < MP; ++r) {
            const long long b = row0 + row_of<C>(tp, r);
            if (b < a.B) {
#pragma unroll
              for (int gj = 0; gj < C::GJ; ++gj)
                *reinterpret_cast<float4*>(&sh[b * NP + gj * C::JSTR + tj * 4]) =
                    make_float4(outv[s][r][gj * 4], outv[s][r][gj * 4 + 1], outv[s][r][gj * 4 + 2], outv[s][r][gj * 4 + 3]);
            }
          }
        }
      }
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int c = 0; c < MJ; ++c) {
          const int j = col_of<C>(tj, c);
#pragma unroll
          for (int gp = 0; gp < C::GP; ++gp) {
            const int p0 = gp * C::PSTR + tp * 4;
            *reinterpret_cast<float4*>(&act[s * NP * TB + act_idx<C>(j, p0)]) = make_float4(
                outv[s][gp * 4][c], outv[s][gp * 4 + 1][c], outv[s][gp * 4 + 2][c], outv[s][gp * 4 + 3][c]);
          }
        }
      __syncthreads();
    };

    // ---- layer 0 ----------------------------------------------------------------------------------------
    {
      const float om = plan_omega(pl, 0);
      const float* M0 = a.packed + pl.off_M0;
      zero_acc();
      for (int kk = 0; kk < K1; ++kk) {
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
          float tmp[MP][MJ];
#pragma unroll
          for (int r = 0; r < MP; ++r)
#pragma unroll
            for (int c = 0; c < MJ; ++c) tmp[r][c] = 0.f;
          for (int i = 0; i < si; ++i) {
            float xv[MP];
            load_rows4(&xs[(s * si + i) * TB], xv);
#pragma unroll
            for (int gj = 0; gj < C::GJ; ++gj) {
              const float4 w = ldg4(&M0[((long long)kk * si + i) * NP + gj * C::JSTR + tj * 4]);
              const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int r = 0; r < MP; ++r)
#pragma unroll
                for (int f = 0; f < 4; ++f) tmp[r][gj * 4 + f] = fmaf(xv[r], wv[f], tmp[r][gj * 4 + f]);
            }
          }
          if (s == 0) fold(0, tmp, 0, kk, om, true);
          else if (s == 1) fold(1, tmp, 0, kk, om, false);
          else if (ND >= 2 && s == 2) fold(ND >= 2 ? 2 : 0, tmp, 0, kk, om, false);
          else if (ND >= 3 && s == 3) fold(ND >= 3 ? 3 : 0, tmp, 0, kk, om, false);
        }
      }
      epilogue(0);
    }

    // ---- hidden layers --------------------------------------------------------------------------------------
    for (int m = 1; m <= H; ++m) {
      const float om = plan_omega(pl, m);
      zero_acc();
      for (int kk = 0; kk < K1; ++kk) {
#pragma unroll 1
        for (int hf = 0; hf < C::NH; ++hf) {
          const float* st = ws.acquire();
#pragma unroll 1
          for (int s = 0; s < NS; ++s) {
            float tmp[MP][MJ];
#pragma unroll
            for (int r = 0; r < MP; ++r)
#pragma unroll
              for (int c = 0; c < MJ; ++c) tmp[r][c] = 0.f;
            mk_gemm<C>(act + s * NP * TB, st, hf * C::NIS, tmp, tp, tj);
            if (s == 0) fold(0, tmp, m, kk, om, hf == 0);
            else if (s == 1) fold(1, tmp, m, kk, om, false);
            else if (ND >= 2 && s == 2) fold(ND >= 2 ? 2 : 0, tmp, m, kk, om, false);
            else if (ND >= 3 && s == 3) fold(ND >= 3 ? 3 : 0, tmp, m, kk, om, false);
          }
          ws.release();
        }
      }
      epilogue(m);
    }

    // ---- last layer: one thread per (row, kappa slice), all streams -------------------------------------------
    {
      const float* ML = a.packed + pl.off_ML;
      const float* CL = C_all + (long long)(H + 1) * K1 * NP;
      const int nsl = NT / TB;
      const int p = tid % TB, q = tid / TB;
      float y[NS][NIF_MAX_SO];
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c) y[s][c] = 0.f;
      for (int kk = q; kk < K1; kk += nsl) {
        float sacc[NS][NIF_MAX_SO];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c) sacc[s][c] = (s == 0 && c < so) ? __ldg(&CL[(long long)kk * NP + c]) : 0.f;
        const float* Mk = ML + (long long)kk * NP * so;
        for (int i = 0; i < n; ++i) {
          float hv[NS];
#pragma unroll
          for (int s = 0; s < NS; ++s) hv[s] = act[s * NP * TB + act_idx<C>(i, p)];
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c)
            if (c < so) {
              const float w = __ldg(&Mk[i * so + c]);
#pragma unroll
              for (int s = 0; s < NS; ++s) sacc[s][c] = fmaf(hv[s], w, sacc[s][c]);
            }
        }
        const float zk = zs[(0 * K1 + kk) * TB + p];
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c) {
          y[0][c] = fmaf(zk, sacc[0][c], y[0][c]);
#pragma unroll
          for (int d = 1; d < NS; ++d) {
            y[d][c] = fmaf(zk, sacc[d][c], y[d][c]);
            y[d][c] = fmaf(zs[(d * K1 + kk) * TB + p], sacc[0][c], y[d][c]);
          }
          if (SEC) {
            y[SEC ? 3 : 0][c] = fmaf(zs[(1 * K1 + kk) * TB + p], sacc[SEC ? 2 : 0][c], y[SEC ? 3 : 0][c]);
            y[SEC ? 3 : 0][c] = fmaf(zs[((SEC ? 2 : 0) * K1 + kk) * TB + p], sacc[1][c], y[SEC ? 3 : 0][c]);
          }
        }
      }
      // combine the kappa slices through shared memory (slot q of thread p)
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c)
          if (c < so) ys[(s * so + c) * NT + q * TB + p] = y[s][c];
      __syncthreads();
      if (q == 0) {
        const long long b = row0 + p;
        if (b < a.B) {
#pragma unroll
          for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int c = 0; c < NIF_MAX_SO; ++c)
              if (c < so) {
                float v = 0.f;
                for (int qq = 0; qq < nsl; ++qq) v += ys[(s * so + c) * NT + qq * TB + p];
                if (s == 0) a.u[b * so + c] = v;
                else if (SEC && s == 3) a.uddot[b * so + c] = v;
                else a.udot[((long long)(s - 1) * a.B + b) * so + c] = v;
              }
        }
      }
      __syncthreads();
    }
  }
}

template <class C, int ND, bool SEC = false>
static int launch_tan(const Plan& pl, TanArgs a, cudaStream_t st) {
  const size_t smem = tan_smem_bytes<C, ND>(pl.K, pl.si, pl.so);
  if (smem > 227 * 1024) {
    nif_set_error("tangent tile needs %zu B of shared memory (latent_dim too large for this build)", smem);
    return NIF_E_UNSUPPORTED;
  }
  const bool res = pl.variant == NIF_VARIANT_SIREN_RES;
  auto kern = res ? nif_tangent_kernel<C, ND, true, SEC> : nif_tangent_kernel<C, ND, false, SEC>;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, occ = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  NIF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, smem));
  if (occ < 1) occ = 1;
  a.total_tiles = (a.B + C::TB - 1) / C::TB;
  long long grid = (long long)sms * occ;
  if (grid > a.total_tiles) grid = a.total_tiles;
  if (grid < 1) return NIF_OK;
  { NIF_PROF("nif_tangent_kernel", st); kern<<<(unsigned)grid, C::NT, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

template <int ND>
static int dispatch_tan(const Plan& pl, const TanArgs& a, cudaStream_t st) {
  switch (pl.NP) {
    case 32: return launch_tan<TCfg32, ND>(pl, a, st);
    case 64: return launch_tan<TCfg64, ND>(pl, a, st);
    case 128: return launch_tan<TCfg128, ND>(pl, a, st);
  }
  nif_set_error("unsupported padded width %d", pl.NP);
  return NIF_E_UNSUPPORTED;
}

int nif_tc_forward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed, float* u,
                        float* save, cudaStream_t st);
int nif_tc_forward_tangent_impl(const Plan& pl, long long B, const float* z, const float* xdot, const float* packed,
                                const float* psave, float* udot, float* tsave, cudaStream_t st);

int nif_tangent_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed, int n_dir,
                     const float* zdot, const float* xdot, float* u, float* udot, float* save, cudaStream_t st) {
  if (save && !zdot && xdot && nif_plan_tc_sobolev(pl)) {
    // Sobolev stash on the tensor cores (tiled slots of nif_tiled_rows(B) x 64 floats): the primal forward with its
    // stash, then one tangent launch per direction that reads d_m / h_{m+1} from it.  nif_sobolev_backward_impl takes
    // the same branch (same predicate, same zdot), so the two agree on the layout.
    int rc = nif_tc_forward_impl(pl, B, z, x, packed, u, save, st);
    if (rc != NIF_OK) return rc;
    const long long slot = nif_tiled_rows(B) * 64;
    for (int d = 0; d < n_dir; ++d) {
      rc = nif_tc_forward_tangent_impl(pl, B, z, xdot + (long long)d * B * pl.si, packed, save,
                                       udot + (long long)d * B * pl.so, save + (2LL + 2 * d) * (pl.H + 1) * slot, st);
      if (rc != NIF_OK) return rc;
    }
    return NIF_OK;
  }
  // directions are processed two at a time (the primal is recomputed per pair)
  for (int d0 = 0; d0 < n_dir; d0 += 2) {
    TanArgs a;
    a.B = B;
    a.total_tiles = 0;
    a.z = z; a.x = x; a.packed = packed;
    a.zdot = zdot ? zdot + (long long)d0 * B * pl.K : nullptr;
    a.xdot = xdot ? xdot + (long long)d0 * B * pl.si : nullptr;
    a.u = u;
    a.udot = udot + (long long)d0 * B * pl.so;
    a.save = save;  // (the primal slots are rewritten with the same values by every pair)
    a.save_t = save ? save + (2LL + 2 * d0) * (pl.H + 1) * B * pl.NP : nullptr;
    a.zddot = nullptr; a.uddot = nullptr;
    const int nd = (n_dir - d0) >= 2 ? 2 : 1;
    const int rc = nd == 2 ? dispatch_tan<2>(pl, a, st) : dispatch_tan<1>(pl, a, st);
    if (rc != NIF_OK) return rc;
  }
  return NIF_OK;
}

// second order: one pair of directions (a, b) per call; udot [2][B][so] = (du/da, du/db), uddot [B][so] = d2u / da db
int nif_tangent2_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                      const float* zdot, const float* xdot, const float* zddot, float* u, float* udot, float* uddot,
                      cudaStream_t st) {
  TanArgs a;
  a.B = B;
  a.total_tiles = 0;
  a.z = z; a.x = x; a.packed = packed; a.zdot = zdot; a.xdot = xdot; a.zddot = zddot;
  a.u = u; a.udot = udot; a.uddot = uddot; a.save = nullptr; a.save_t = nullptr;
  switch (pl.NP) {
    case 32: return launch_tan<TCfg32, 3, true>(pl, a, st);
    case 64: return launch_tan<TCfg64, 3, true>(pl, a, st);
    case 128: return launch_tan<TCfg128, 3, true>(pl, a, st);
  }
  nif_set_error("unsupported padded width %d", pl.NP);
  return NIF_E_UNSUPPORTED;
}
