// Fused forward of the NIF hot path (fp32 CUDA-core parity path).
//
// Replaces, in one persistent kernel and without ever forming the (B, po_dim) tensor:
//   HyperLinearForSIREN.call            nif/layers/siren.py:514-522   (pnet_output = z @ w + b)
//   NIF._call_shape_net                 nif/model.py:233-324
//   NIFMultiScale._call_shape_net_mres  nif/model.py:738-954
//   EinsumLayer("ai,aij->aj")           nif/layers/mlp.py:209-219
//
// Math (re-associated):  for layer m with matrix block M_m[kappa][i][j] of [w_h; b_h]
//   pre[b][j] = sum_kappa zt[b][kappa] * ( omega * sum_i h[b][i] M_m[kappa][i][j] + C_m[kappa][j] ),
// zt = [z, 1].  The inner sum over i is a plain GEMM with a weight operand shared by all rows
// (streamed L2 -> SMEM by cp.async.bulk), the outer sum over kappa is a per-row scaled
// accumulation done in registers.  Per-sample weights never exist.
#include "nif_tile.cuh"

struct FwdArgs {
  long long G, B, tiles_per_group, total_tiles;
  const float* z;
  const float* x;
  int x_shared;
  const float* packed;
  float* u;
  float* save;  // [2*(H+1)][B][NP] or null
};

template <class C>
__host__ __device__ inline size_t fwd_smem_bytes(int K, int si, int so) {
  size_t f = 2 * (size_t)C::STAGE_FLOATS + (size_t)C::NP * C::TB + (size_t)(K + 1) * C::TB + (size_t)si * C::TB +
             (size_t)so * C::NT;
  return f * 4 + 64;
}

template <class C, bool RES, bool SAVE>
__global__ void __launch_bounds__(C::NT, (RES || C::NP == 32) ? 1 : 2) nif_fwd_kernel(const Plan pl, const FwdArgs a) {
  constexpr int NP = C::NP, TB = C::TB, MP = C::MP, MJ = C::MJ, NT = C::NT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* act = stage + 2 * C::STAGE_FLOATS;
  float* zs = act + NP * TB;
  float* xs = zs + (pl.K + 1) * TB;
  float* ys = xs + pl.si * TB;
  uint64_t* bar = reinterpret_cast<uint64_t*>(ys + pl.so * NT);
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
  const int HT = H + pl.wide_last;  // matrices streamed through the GEMM path (wide_last: the last matrix too)
  ws.chunks_per_tile = HT * K1 * C::NH;
  ws.total = my_tiles * ws.chunks_per_tile;
  ws.issued = 0;
  ws.consumed = 0;
  ws.H = HT;
  ws.K1 = K1;
  ws.packed = a.packed;
  ws.packed_floats = pl.packed_floats;
  ws.sec_off = pl.off_MH;
  ws.tiles_per_group = a.tiles_per_group;
  ws.reverse = false;
  if (tid == 0) {
    while (ws.issued < 2 && ws.issued < ws.total) ws.issue_one();
  }

  for (long long t = 0; t < my_tiles; ++t) {
    const long long tile = blockIdx.x + t * gridDim.x;
    const long long g = tile / a.tiles_per_group;
    const long long row0 = (tile - g * a.tiles_per_group) * TB;
    const float* pk = a.packed + g * pl.packed_floats;
    const float* C_all = pk + pl.off_C;

    // ---- stage latent codes and coordinates (zero rows beyond B) --------------------------------
    for (int idx = tid; idx < TB * K; idx += NT) {
      const int p = idx / K, kk = idx - p * K;
      const long long b = row0 + p;
      zs[kk * TB + p] = (b < a.B) ? __ldg(&a.z[(g * a.B + b) * K + kk]) : 0.f;
    }
    for (int p = tid; p < TB; p += NT) zs[K * TB + p] = 1.f;
    for (int idx = tid; idx < TB * si; idx += NT) {
      const int p = idx / si, i = idx - p * si;
      const long long b = row0 + p;
      const long long xr = a.x_shared ? b : (g * a.B + b);
      xs[i * TB + p] = (b < a.B) ? __ldg(&a.x[xr * si + i]) : 0.f;
    }
    __syncthreads();

    float acc[MP][MJ];
    float carry[RES ? MP : 1][RES ? MJ : 1];

    // epilogue of layer m (0..H): activation, residual, stash for the reverse pass, write-back
    auto epilogue = [&](int m) {
      if (m == H + 1) {  // wide last matrix (trunk plans): linear, straight to the output rows
#pragma unroll
        for (int r = 0; r < MP; ++r) {
          const long long b = row0 + row_of<C>(tp, r);
          if (b < a.B) {
#pragma unroll
            for (int c = 0; c < MJ; ++c) {
              const int j = col_of<C>(tj, c);
              if (j < so) a.u[(g * a.B + b) * so + j] = acc[r][c];
            }
          }
        }
        __syncthreads();
        return;
      }
      const float alpha = plan_alpha(pl, m);
      const int res = plan_res(pl, m);
      float outv[MP][MJ], dv[MP][MJ], fv[MP][MJ];
#pragma unroll
      for (int r = 0; r < MP; ++r)
#pragma unroll
        for (int c4 = 0; c4 < MJ; c4 += 4) {
          const float v4[4] = {acc[r][c4], acc[r][c4 + 1], acc[r][c4 + 2], acc[r][c4 + 3]};
          float f4[4], d4[4];
          act_fd4(pl.act, v4, f4, d4);
#pragma unroll
          for (int e = 0; e < 4; ++e) { fv[r][c4 + e] = f4[e]; dv[r][c4 + e] = d4[e]; }
        }
#pragma unroll
      for (int r = 0; r < MP; ++r)
#pragma unroll
        for (int c = 0; c < MJ; ++c) {
          float f = fv[r][c], d = alpha * dv[r][c];
          float o = alpha * f;
          const int j = col_of<C>(tj, c);
          if (res == 1) o += act[act_idx<C>(j, row_of<C>(tp, r))];
          if (RES) {
            if (res == 2) carry[RES ? r : 0][RES ? c : 0] = act[act_idx<C>(j, row_of<C>(tp, r))];
            if (res == 3) o += 0.5f * carry[RES ? r : 0][RES ? c : 0];
          }
          if (j >= n) { o = 0.f; d = 0.f; }
          outv[r][c] = o;
          dv[r][c] = d;
        }
      if (SAVE) {
        float* sh = a.save + (long long)m * a.B * NP;                // h_{m+1}
        float* sd = a.save + (long long)(H + 1 + m) * a.B * NP;      // d_m
#pragma unroll
        for (int r = 0; r < MP; ++r) {
          const long long b = row0 + row_of<C>(tp, r);
          if (b < a.B) {
#pragma unroll
            for (int gj = 0; gj < C::GJ; ++gj) {
              const int j0 = gj * C::JSTR + tj * 4;
              *reinterpret_cast<float4*>(&sh[b * NP + j0]) =
                  make_float4(outv[r][gj * 4], outv[r][gj * 4 + 1], outv[r][gj * 4 + 2], outv[r][gj * 4 + 3]);
              *reinterpret_cast<float4*>(&sd[b * NP + j0]) =
                  make_float4(dv[r][gj * 4], dv[r][gj * 4 + 1], dv[r][gj * 4 + 2], dv[r][gj * 4 + 3]);
            }
          }
        }
      }
      // every read of `act` by this layer's GEMM is behind the barrier in ws.release()
#pragma unroll
      for (int c = 0; c < MJ; ++c) {
        const int j = col_of<C>(tj, c);
#pragma unroll
        for (int gp = 0; gp < C::GP; ++gp) {
          const int p0 = gp * C::PSTR + tp * 4;
          *reinterpret_cast<float4*>(&act[act_idx<C>(j, p0)]) =
              make_float4(outv[gp * 4][c], outv[gp * 4 + 1][c], outv[gp * 4 + 2][c], outv[gp * 4 + 3][c]);
        }
      }
      __syncthreads();
    };

    // ---- layer 0: si -> n (thin; weights straight from L2 through the read-only path) -------------
    {
      const float om = plan_omega(pl, 0);
      const float* M0 = pk + pl.off_M0;
#pragma unroll
      for (int r = 0; r < MP; ++r)
#pragma unroll
        for (int c = 0; c < MJ; ++c) acc[r][c] = 0.f;
      for (int kk = 0; kk < K1; ++kk) {
        float tmp[MP][MJ];
#pragma unroll
        for (int r = 0; r < MP; ++r)
#pragma unroll
          for (int c = 0; c < MJ; ++c) tmp[r][c] = 0.f;
        for (int i = 0; i < si; ++i) {
          float xv[MP];
#pragma unroll
          for (int gp = 0; gp < C::GP; ++gp) {
            const float4 q = *reinterpret_cast<const float4*>(&xs[i * TB + gp * C::PSTR + tp * 4]);
            xv[gp * 4] = q.x; xv[gp * 4 + 1] = q.y; xv[gp * 4 + 2] = q.z; xv[gp * 4 + 3] = q.w;
          }
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
        float zv[MP];
#pragma unroll
        for (int gp = 0; gp < C::GP; ++gp) {
          const float4 q = *reinterpret_cast<const float4*>(&zs[kk * TB + gp * C::PSTR + tp * 4]);
          zv[gp * 4] = q.x; zv[gp * 4 + 1] = q.y; zv[gp * 4 + 2] = q.z; zv[gp * 4 + 3] = q.w;
        }
#pragma unroll
        for (int gj = 0; gj < C::GJ; ++gj) {
          const float4 cb = ldg4(&C_all[((long long)0 * K1 + kk) * NP + gj * C::JSTR + tj * 4]);
          const float cv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
          for (int r = 0; r < MP; ++r)
#pragma unroll
            for (int f = 0; f < 4; ++f)
              acc[r][gj * 4 + f] = fmaf(zv[r], fmaf(om, tmp[r][gj * 4 + f], cv[f]), acc[r][gj * 4 + f]);
        }
      }
      epilogue(0);
    }

    // ---- hidden layers: n -> n ------------------------------------------------------------------------
    for (int m = 1; m <= HT; ++m) {
      const float om = plan_omega(pl, m);
#pragma unroll
      for (int r = 0; r < MP; ++r)
#pragma unroll
        for (int c = 0; c < MJ; ++c) acc[r][c] = 0.f;
      for (int kk = 0; kk < K1; ++kk) {
        float tmp[MP][MJ];
#pragma unroll
        for (int r = 0; r < MP; ++r)
#pragma unroll
          for (int c = 0; c < MJ; ++c) tmp[r][c] = 0.f;
#pragma unroll 1
        for (int hf = 0; hf < C::NH; ++hf) {
          const float* st = ws.acquire();
          mk_gemm<C>(act, st, hf * C::NIS, tmp, tp, tj);
          ws.release();
        }
        float zv[MP];
#pragma unroll
        for (int gp = 0; gp < C::GP; ++gp) {
          const float4 q = *reinterpret_cast<const float4*>(&zs[kk * TB + gp * C::PSTR + tp * 4]);
          zv[gp * 4] = q.x; zv[gp * 4 + 1] = q.y; zv[gp * 4 + 2] = q.z; zv[gp * 4 + 3] = q.w;
        }
#pragma unroll
        for (int gj = 0; gj < C::GJ; ++gj) {
          const float4 cb = ldg4(&C_all[((long long)m * K1 + kk) * NP + gj * C::JSTR + tj * 4]);
          const float cv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
          for (int r = 0; r < MP; ++r)
#pragma unroll
            for (int f = 0; f < 4; ++f)
              acc[r][gj * 4 + f] = fmaf(zv[r], fmaf(om, tmp[r][gj * 4 + f], cv[f]), acc[r][gj * 4 + f]);
        }
      }
      epilogue(m);
    }

    // ---- last layer: n -> so (thin).  Two threads per row split the kappa range -------------------------
    if (!pl.wide_last) {
      const float* ML = pk + pl.off_ML;
      const float* CL = C_all + (long long)(H + 1) * K1 * NP;
      const int nsl = NT / TB;  // kappa slices
      const int p = tid % TB, q = tid / TB;
      float y[NIF_MAX_SO];
#pragma unroll
      for (int c = 0; c < NIF_MAX_SO; ++c) y[c] = 0.f;
      for (int kk = q; kk < K1; kk += nsl) {
        float s[NIF_MAX_SO];
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c) s[c] = (c < so) ? __ldg(&CL[(long long)kk * NP + c]) : 0.f;
        const float* Mk = ML + (long long)kk * NP * so;
        for (int i = 0; i < n; ++i) {
          const float hv = act[act_idx<C>(i, p)];
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c)
            if (c < so) s[c] = fmaf(hv, __ldg(&Mk[i * so + c]), s[c]);
        }
        const float zk = zs[kk * TB + p];
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c) y[c] = fmaf(zk, s[c], y[c]);
      }
      if (q > 0) {
#pragma unroll
        for (int c = 0; c < NIF_MAX_SO; ++c)
          if (c < so) ys[((q - 1) * so + c) * TB + p] = y[c];
      }
      __syncthreads();
      if (q == 0) {
        const long long b = row0 + p;
        for (int qq = 1; qq < nsl; ++qq)
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c)
            if (c < so) y[c] += ys[((qq - 1) * so + c) * TB + p];
        if (b < a.B) {
#pragma unroll
          for (int c = 0; c < NIF_MAX_SO; ++c)
            if (c < so) a.u[(g * a.B + b) * so + c] = y[c];
        }
      }
      __syncthreads();  // act / zs / xs / ys are rewritten by the next tile
    }
  }
}

// ---------------------------------------------------------------------------------------------------
template <class C, bool RES, bool SAVE>
static int launch_fwd(const Plan& pl, const FwdArgs& a, cudaStream_t st) {
  const size_t smem = fwd_smem_bytes<C>(pl.K, pl.si, pl.so);
  if (smem > 227 * 1024) {
    nif_set_error("forward tile needs %zu B of shared memory (latent_dim too large for this build)", smem);
    return NIF_E_UNSUPPORTED;
  }
  auto kern = nif_fwd_kernel<C, RES, SAVE>;
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, occ = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  NIF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, smem));
  if (occ < 1) occ = 1;
  long long grid = (long long)sms * occ;
  if (grid > a.total_tiles) grid = a.total_tiles;
  if (grid < 1) return NIF_OK;
  { NIF_PROF("nif_fwd_kernel", st); kern<<<(unsigned)grid, C::NT, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

template <class C>
static int dispatch_fwd(const Plan& pl, const FwdArgs& a, cudaStream_t st) {
  const bool res = pl.variant == NIF_VARIANT_SIREN_RES;
  const bool save = a.save != nullptr;
  if (res) return save ? launch_fwd<C, true, true>(pl, a, st) : launch_fwd<C, true, false>(pl, a, st);
  return save ? launch_fwd<C, false, true>(pl, a, st) : launch_fwd<C, false, false>(pl, a, st);
}

int nif_tc_forward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed, float* u,
                        float* save, cudaStream_t st);

int nif_bf_forward_impl(const Plan& pl, long long G, long long B, const float* z, const float* x, int x_shared,
                        const float* packed, float* u, float* save, cudaStream_t st);

int nif_forward_impl(const Plan& pl, long long G, long long B, const float* z, const float* x, int x_shared,
                     const float* packed, float* u, float* save, cudaStream_t st) {
  if (pl.bf) {  // bf16 tensor-core path (grouped launches included); other shapes fall through
    const int rc = nif_bf_forward_impl(pl, G, B, z, x, x_shared, packed, u, save, st);
    if (rc != NIF_E_UNSUPPORTED) return rc;
  }
  if (pl.tc && G == 1) {  // tensor-core path; shapes it does not cover fall through to the CUDA-core kernel
    const int rc = nif_tc_forward_impl(pl, B, z, x, packed, u, save, st);
    if (rc != NIF_E_UNSUPPORTED) return rc;
  }
  FwdArgs a;
  a.G = G;
  a.B = B;
  a.z = z;
  a.x = x;
  a.x_shared = x_shared;
  a.packed = packed;
  a.u = u;
  a.save = save;
  const int TB = pl.NP == 128 ? Cfg128::TB : (pl.NP == 64 ? Cfg64::TB : Cfg32::TB);
  a.tiles_per_group = (B + TB - 1) / TB;
  a.total_tiles = a.tiles_per_group * G;
  switch (pl.NP) {
    case 32: return dispatch_fwd<Cfg32>(pl, a, st);
    case 64: return dispatch_fwd<Cfg64>(pl, a, st);
    case 128: return dispatch_fwd<Cfg128>(pl, a, st);
  }
  nif_set_error("unsupported padded width %d", pl.NP);
  return NIF_E_UNSUPPORTED;
}
