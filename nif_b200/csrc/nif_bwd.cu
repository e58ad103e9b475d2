// Reverse pass of the fused NIF hot path (fp32 CUDA-core parity path).
//
// What Keras' GradientTape does for the graph of nif/model.py:130-154 / 510-539, done
// without the (B, po_dim) tensor and without the zero-padded (B, po_dim) slice gradients:
//
//   1. nif_bwd_data_kernel   per row tile, layers last -> first:
//        da_m = dh_{m+1} * act'(pre_m)                       (stashed for step 2)
//        T[b][kappa][i] = sum_j da_m[b][j] M_m[kappa][i][j]  (GEMM, shared transposed weights)
//        dh_m[b][i]  = omega * sum_kappa zt[b][kappa] T[b][kappa][i]
//        dz[b][kappa] += omega * sum_i T[b][kappa][i] h_m[b][i] + sum_j C_m[kappa][j] da_m[b][j]
//   2. nif_bwd_weight_kernel  dM_m[kappa][i][j] = omega * sum_b zt[b][kappa] h_m[b][i] da_m[b][j]
//        a batch-reduction GEMM; the rank-1 update h (x) da is formed once per row in registers and
//        scaled by 4 latent coordinates.  Batch splits write partials (deterministic, no atomics).
//   3. nif_bwd_edge_kernel    the thin terms (all bias rows, first and last matrix) as
//        out[kappa][q] = sum_b zt[b][kappa] F[b][q]  with on-the-fly features F.
//   4. nif_unpack_grad (nif_pack.cu) sums the partials into dw_h [K,P], db_h [P] (reference layout).
#include "nif_tile.cuh"

struct BwdArgs {
  long long B, total_tiles;
  const float *z, *x, *packed, *save, *du;
  float* da;  // [(H+1)][B][NP]
  float* dz;  // [B][K]
  // reverse-over-forward (Sobolev training); all null / 0 for the plain reverse pass:
  const float* h_stash;  // the h_m slots this pass pairs with its da_m (save, or the tangent activations h'_m)
  const float* e_stash;  // [H+1] slots e_m; with ext_out: ext_out[m] = dh_{m+1} * e_m (tangent-adjoint pass)
  float* ext_out;
  const float* ext_add;  // [H+1] slots added to da_m (primal-adjoint pass consumes what the tangent pass wrote)
  int no_bias;           // tangent-adjoint pass: bias rows do not enter pre_m', drop their dz terms
  int dz_accumulate;     // dz += instead of =
  int ext_accumulate;    // ext_out += instead of = (second and later directions)
  // directions that move the latent code (d/d ParameterNet input): pre_m' also holds F_m = sum_k zdot_k (w h_m M_k + C_k).
  // "Local" pass (da_in != null): the latent operand is zdot (bias coordinate zt_last = 0), the activations are the
  // primal h_m, and each layer takes its adjoint da_m' from da_in[m] (left by the tangent-adjoint pass) instead of the
  // chain: dz = dL/dzdot, and dh_out[m] += w sum_k zdot_k M_{m+1,k} da'_{m+1}, the source the primal pass adds to
  // dh_{m+1} through dh_add[m].
  const float* da_in;
  float* dh_out;
  const float* dh_add;
  float zt_last;         // value of the bias coordinate of zt: 1, or 0 in the local pass
};

template <class C>
__host__ __device__ inline size_t bwd_smem_bytes(int K, int si, int so) {
  size_t f = 2 * (size_t)C::STAGE_FLOATS + (size_t)C::NP * C::TB + 2 * (size_t)(K + 1) * C::TB + (size_t)si * C::TB +
             (size_t)so * C::TB;
  return f * 4 + 64;
}

// EXT: the reverse-over-forward hooks of BwdArgs are live (Sobolev training); the plain pass compiles them away
template <class C, bool RES, bool EXT>
__global__ void __launch_bounds__(C::NT, 1) nif_bwd_data_kernel(const Plan pl, const BwdArgs a_in) {
  BwdArgs a = a_in;
  if (!EXT) {
    a.h_stash = a.save; a.e_stash = nullptr; a.ext_out = nullptr; a.ext_add = nullptr; a.no_bias = 0; a.dz_accumulate = 0;
    a.ext_accumulate = 0; a.da_in = nullptr; a.dh_out = nullptr; a.dh_add = nullptr; a.zt_last = 1.f;
  }
  const bool local = EXT && a.da_in != nullptr;
  constexpr int NP = C::NP, TB = C::TB, MP = C::MP, MJ = C::MJ, NT = C::NT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* dact = stage + 2 * C::STAGE_FLOATS;  // da_m, k-major swizzled (GEMM A operand)
  float* zs = dact + NP * TB;                 // [K+1][TB]
  float* dzs = zs + (pl.K + 1) * TB;          // [K+1][TB]
  float* xs = dzs + (pl.K + 1) * TB;          // [si][TB]
  float* dys = xs + pl.si * TB;               // [so][TB]
  uint64_t* bar = reinterpret_cast<uint64_t*>(dys + pl.so * TB);
  bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bar) + 7) & ~uintptr_t(7));

  const int tid = threadIdx.x;
  const int tj = tid % C::TY, tp = tid / C::TY;
  const int K = pl.K, K1 = pl.K + 1, H = pl.H, si = pl.si, so = pl.so;

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
  ws.sec_off = pl.off_MHT;
  ws.tiles_per_group = a.total_tiles;  // single group
  ws.reverse = true;
  if (tid == 0) {
    while (ws.issued < 2 && ws.issued < ws.total) ws.issue_one();
  }

  const float* C_all = a.packed + pl.off_C;

  // sum s[r] over the TY lanes that share a row group, then lane tj == 0 accumulates into dzs[kk]
  auto dz_commit = [&](float (&s)[MP], int kk) {
#pragma unroll
    for (int off = C::TY / 2; off >= 1; off >>= 1)
#pragma unroll
      for (int r = 0; r < MP; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], off);
    if (tj == 0) {
#pragma unroll
      for (int r = 0; r < MP; ++r) dzs[kk * TB + row_of<C>(tp, r)] += s[r];
    }
  };

  for (long long t = 0; t < my_tiles; ++t) {
    const long long tile = blockIdx.x + t * gridDim.x;
    const long long row0 = tile * TB;

    for (int idx = tid; idx < TB * K; idx += NT) {
      const int p = idx / K, kk = idx - p * K;
      const long long b = row0 + p;
      zs[kk * TB + p] = (b < a.B) ? __ldg(&a.z[b * K + kk]) : 0.f;
    }
    for (int p = tid; p < TB; p += NT) zs[K * TB + p] = a.zt_last;
    for (int idx = tid; idx < TB * K1; idx += NT) dzs[idx] = 0.f;
    for (int idx = tid; idx < TB * si; idx += NT) {
      const int p = idx / si, i = idx - p * si;
      const long long b = row0 + p;
      xs[i * TB + p] = (b < a.B) ? __ldg(&a.x[b * si + i]) : 0.f;
    }
    for (int idx = tid; idx < TB * so; idx += NT) {
      const int p = idx / so, c = idx - p * so;
      const long long b = row0 + p;
      dys[c * TB + p] = (b < a.B) ? __ldg(&a.du[b * so + c]) : 0.f;
    }
    __syncthreads();

    long long brow[MP];
#pragma unroll
    for (int r = 0; r < MP; ++r) brow[r] = row0 + row_of<C>(tp, r);

    float acc[MP][MJ];
    float carry[RES ? MP : 1][RES ? MJ : 1];

    // ---- last layer (matrix H+1): n -> so --------------------------------------------------------------
    if (pl.wide_last) {
      // wide output (trunk plans): dh_{H+1} = sum_kappa zt[kappa] (du @ ML[kappa]^T) through the tile GEMM, with the
      // seed du (zero padded to NP columns) as the A operand and the transposed last matrix streamed as matrix H
      for (int idx = tid; idx < TB * NP; idx += NT) {
        const int p = idx / NP, c = idx - p * NP;
        const long long b = row0 + p;
        const float v = (b < a.B && c < so) ? __ldg(&a.du[b * so + c]) : 0.f;
        dact[act_idx<C>(c, p)] = v;
        if (b < a.B) a.da[(long long)(H + 1) * a.B * NP + b * NP + c] = v;  // operand of the last matrix's gradient
      }
      __syncthreads();
      const float* hL = a.h_stash + (long long)H * a.B * NP;
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
          mk_gemm<C>(dact, st, hf * C::NIS, tmp, tp, tj);
          if (!(kk == K1 - 1 && hf == C::NH - 1)) ws.release();
        }
        float s[MP];
#pragma unroll
        for (int r = 0; r < MP; ++r) {
          const float zk = zs[kk * TB + row_of<C>(tp, r)];
          float sr = 0.f;
#pragma unroll
          for (int gj = 0; gj < C::GJ; ++gj) {
            float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (brow[r] < a.B) hv = ldg4(&hL[brow[r] * NP + gj * C::JSTR + tj * 4]);
            const float hvv[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int f = 0; f < 4; ++f) {
              acc[r][gj * 4 + f] = fmaf(zk, tmp[r][gj * 4 + f], acc[r][gj * 4 + f]);
              sr = fmaf(tmp[r][gj * 4 + f], hvv[f], sr);
            }
          }
          s[r] = sr;
        }
        dz_commit(s, kk);  // (bias-row part of dz is irrelevant for the K = 0 trunk plans that use this path)
        if (kk == K1 - 1) ws.release();
      }
    } else {
      const float* ML = a.packed + pl.off_ML;
      const float* CL = C_all + (long long)(H + 1) * K1 * NP;
      const float* hL = a.h_stash + (long long)H * a.B * NP;  // h_{H+1}
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
        float s[MP];
#pragma unroll
        for (int r = 0; r < MP; ++r) s[r] = 0.f;
        for (int cc = 0; cc < so; ++cc) {
          float dyv[MP];
#pragma unroll
          for (int r = 0; r < MP; ++r) dyv[r] = dys[cc * TB + row_of<C>(tp, r)];
#pragma unroll
          for (int c = 0; c < MJ; ++c) {
            const float w = __ldg(&ML[((long long)kk * NP + col_of<C>(tj, c)) * so + cc]);
#pragma unroll
            for (int r = 0; r < MP; ++r) tmp[r][c] = fmaf(dyv[r], w, tmp[r][c]);
          }
          if (tj == 0 && !(EXT && a.no_bias)) {
            const float cb = __ldg(&CL[(long long)kk * NP + cc]);
#pragma unroll
            for (int r = 0; r < MP; ++r) s[r] = fmaf(cb, dyv[r], s[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < MP; ++r) {
          const float zk = zs[kk * TB + row_of<C>(tp, r)];
#pragma unroll
          for (int gj = 0; gj < C::GJ; ++gj) {
            float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (brow[r] < a.B) hv = ldg4(&hL[brow[r] * NP + gj * C::JSTR + tj * 4]);
            const float hvv[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int f = 0; f < 4; ++f) {
              acc[r][gj * 4 + f] = fmaf(zk, tmp[r][gj * 4 + f], acc[r][gj * 4 + f]);
              s[r] = fmaf(tmp[r][gj * 4 + f], hvv[f], s[r]);
            }
          }
        }
        dz_commit(s, kk);
      }
    }

    // ---- layers H .. 0 --------------------------------------------------------------------------------
    for (int m = H; m >= 0; --m) {
      const float om = plan_omega(pl, m);
      const int res = plan_res(pl, m);
      const float* dsv = a.save + (long long)(H + 1 + m) * a.B * NP;  // d_m
      float* dag = a.da + (long long)m * a.B * NP;
      float daf[MP][MJ];
      if (EXT && a.dh_add) {  // sources of dh_{m+1} left by the local passes
#pragma unroll
        for (int r = 0; r < MP; ++r)
#pragma unroll
          for (int gj = 0; gj < C::GJ; ++gj)
            if (brow[r] < a.B) {
              const float4 sv = ldg4(&a.dh_add[(long long)m * a.B * NP + brow[r] * NP + gj * C::JSTR + tj * 4]);
              acc[r][gj * 4 + 0] += sv.x; acc[r][gj * 4 + 1] += sv.y; acc[r][gj * 4 + 2] += sv.z; acc[r][gj * 4 + 3] += sv.w;
            }
      }
      if (local) {
#pragma unroll
        for (int r = 0; r < MP; ++r)
#pragma unroll
          for (int gj = 0; gj < C::GJ; ++gj) {
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (brow[r] < a.B) {
              const long long o = (long long)m * a.B * NP + brow[r] * NP + gj * C::JSTR + tj * 4;
              float4 sv = *reinterpret_cast<const float4*>(&a.dh_out[o]);
              sv.x += acc[r][gj * 4 + 0]; sv.y += acc[r][gj * 4 + 1]; sv.z += acc[r][gj * 4 + 2]; sv.w += acc[r][gj * 4 + 3];
              *reinterpret_cast<float4*>(&a.dh_out[o]) = sv;
              gv = ldg4(&a.da_in[o]);
            }
            daf[r][gj * 4 + 0] = gv.x; daf[r][gj * 4 + 1] = gv.y; daf[r][gj * 4 + 2] = gv.z; daf[r][gj * 4 + 3] = gv.w;
          }
      } else {
#pragma unroll
      for (int r = 0; r < MP; ++r) {
#pragma unroll
        for (int gj = 0; gj < C::GJ; ++gj) {
          float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (brow[r] < a.B) dv = ldg4(&dsv[brow[r] * NP + gj * C::JSTR + tj * 4]);
          daf[r][gj * 4 + 0] = acc[r][gj * 4 + 0] * dv.x;
          daf[r][gj * 4 + 1] = acc[r][gj * 4 + 1] * dv.y;
          daf[r][gj * 4 + 2] = acc[r][gj * 4 + 2] * dv.z;
          daf[r][gj * 4 + 3] = acc[r][gj * 4 + 3] * dv.w;
          if (EXT && a.ext_add && brow[r] < a.B) {  // + dh'_{m+1} * e_m, written by the tangent-adjoint pass
            const float4 xv = ldg4(&a.ext_add[(long long)m * a.B * NP + brow[r] * NP + gj * C::JSTR + tj * 4]);
            daf[r][gj * 4 + 0] += xv.x; daf[r][gj * 4 + 1] += xv.y; daf[r][gj * 4 + 2] += xv.z; daf[r][gj * 4 + 3] += xv.w;
          }
          if (EXT && a.ext_out && brow[r] < a.B) {
            const float4 ev = ldg4(&a.e_stash[(long long)m * a.B * NP + brow[r] * NP + gj * C::JSTR + tj * 4]);
            float4* xo = reinterpret_cast<float4*>(&a.ext_out[(long long)m * a.B * NP + brow[r] * NP + gj * C::JSTR + tj * 4]);
            float4 xv = make_float4(acc[r][gj * 4 + 0] * ev.x, acc[r][gj * 4 + 1] * ev.y, acc[r][gj * 4 + 2] * ev.z, acc[r][gj * 4 + 3] * ev.w);
            if (a.ext_accumulate) { const float4 o = *xo; xv.x += o.x; xv.y += o.y; xv.z += o.z; xv.w += o.w; }
            *xo = xv;
          }
          if (brow[r] < a.B)
            *reinterpret_cast<float4*>(&dag[brow[r] * NP + gj * C::JSTR + tj * 4]) =
                make_float4(daf[r][gj * 4], daf[r][gj * 4 + 1], daf[r][gj * 4 + 2], daf[r][gj * 4 + 3]);
        }
      }
      }
      if (m >= 1) {
        // residual bookkeeping (mirror of the forward epilogue)
#pragma unroll
        for (int r = 0; r < MP; ++r)
#pragma unroll
          for (int c = 0; c < MJ; ++c) {
            if (local) acc[r][c] = 0.f;  // no chain: every layer's term is a source of the primal pass
            else if (RES) {
              if (res == 3) { carry[RES ? r : 0][RES ? c : 0] = 0.5f * acc[r][c]; acc[r][c] = 0.f; }
              else if (res == 2) acc[r][c] = carry[RES ? r : 0][RES ? c : 0];
              else acc[r][c] = 0.f;
            } else {
              if (res != 1) acc[r][c] = 0.f;  // res == 1: dh_m = T + dh_{m+1}
            }
          }
        // da_m -> shared (k-major) as the GEMM A operand.  All reads of the previous contents are
        // behind the barrier of the last ws.release() / the tile prologue.
#pragma unroll
        for (int c = 0; c < MJ; ++c) {
          const int j = col_of<C>(tj, c);
#pragma unroll
          for (int gp = 0; gp < C::GP; ++gp) {
            const int p0 = gp * C::PSTR + tp * 4;
            *reinterpret_cast<float4*>(&dact[act_idx<C>(j, p0)]) =
                make_float4(daf[gp * 4][c], daf[gp * 4 + 1][c], daf[gp * 4 + 2][c], daf[gp * 4 + 3][c]);
          }
        }
        __syncthreads();
        const float* hm = a.h_stash + (long long)(m - 1) * a.B * NP;  // h_m
        for (int kk = 0; kk < K1; ++kk) {
          float tmp[MP][MJ];
#pragma unroll
          for (int r = 0; r < MP; ++r)
#pragma unroll
            for (int c = 0; c < MJ; ++c) tmp[r][c] = 0.f;
#pragma unroll 1
          for (int hf = 0; hf < C::NH; ++hf) {
            const float* st = ws.acquire();
            mk_gemm<C>(dact, st, hf * C::NIS, tmp, tp, tj);
            // the last chunk of this layer must not release before dact has been re-read below
            if (!(kk == K1 - 1 && hf == C::NH - 1)) ws.release();
          }
          float s[MP];
#pragma unroll
          for (int r = 0; r < MP; ++r) {
            const float zk = zs[kk * TB + row_of<C>(tp, r)];
            float sr = 0.f;
#pragma unroll
            for (int gj = 0; gj < C::GJ; ++gj) {
              float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (brow[r] < a.B) hv = ldg4(&hm[brow[r] * NP + gj * C::JSTR + tj * 4]);
              const float4 cb = ldg4(&C_all[((long long)m * K1 + kk) * NP + gj * C::JSTR + tj * 4]);
              const float hvv[4] = {hv.x, hv.y, hv.z, hv.w};
              const float cvv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
              for (int f = 0; f < 4; ++f) {
                const int c = gj * 4 + f;
                const float tv = om * tmp[r][c];
                acc[r][c] = fmaf(zk, tv, acc[r][c]);
                sr = fmaf(tv, hvv[f], sr);
                if (!(EXT && a.no_bias)) sr = fmaf(cvv[f], dact[act_idx<C>(col_of<C>(tj, c), row_of<C>(tp, r))], sr);
              }
            }
            s[r] = sr;
          }
          dz_commit(s, kk);
          if (kk == K1 - 1) ws.release();
        }
      } else {
        // first matrix: only dz is needed (no gradient w.r.t. the coordinates on this path)
        const float* M0 = a.packed + pl.off_M0;
        for (int kk = 0; kk < K1; ++kk) {
          float s[MP];
#pragma unroll
          for (int r = 0; r < MP; ++r) s[r] = 0.f;
#pragma unroll
          for (int c = 0; c < MJ; ++c) {
            const int j = col_of<C>(tj, c);
            const float cb = (EXT && a.no_bias) ? 0.f : __ldg(&C_all[((long long)0 * K1 + kk) * NP + j]);
            float wv[NIF_MAX_SI];
#pragma unroll
            for (int i = 0; i < NIF_MAX_SI; ++i) wv[i] = (i < si) ? __ldg(&M0[((long long)kk * si + i) * NP + j]) : 0.f;
#pragma unroll
            for (int r = 0; r < MP; ++r) {
              float lin = 0.f;
#pragma unroll
              for (int i = 0; i < NIF_MAX_SI; ++i)
                if (i < si) lin = fmaf(xs[i * TB + row_of<C>(tp, r)], wv[i], lin);
              s[r] = fmaf(daf[r][c], fmaf(om, lin, cb), s[r]);
            }
          }
          dz_commit(s, kk);
        }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < TB * K; idx += NT) {
      const int p = idx / K, kk = idx - p * K;
      const long long b = row0 + p;
      if (b < a.B) a.dz[b * K + kk] = (EXT && a.dz_accumulate) ? a.dz[b * K + kk] + dzs[kk * TB + p] : dzs[kk * TB + p];
    }
    __syncthreads();
  }
}

template <class C>
static int launch_bwd_data(const Plan& pl, const BwdArgs& a, cudaStream_t st) {
  const size_t smem = bwd_smem_bytes<C>(pl.K, pl.si, pl.so);
  if (smem > 227 * 1024) {
    nif_set_error("reverse tile needs %zu B of shared memory (latent_dim too large for this build)", smem);
    return NIF_E_UNSUPPORTED;
  }
  const bool res = pl.variant == NIF_VARIANT_SIREN_RES;
  const bool ext = a.h_stash != a.save || a.ext_add || a.ext_out || a.no_bias || a.dz_accumulate || a.da_in || a.dh_add ||
                   a.zt_last != 1.f;
  auto kern = ext ? (res ? nif_bwd_data_kernel<C, true, true> : nif_bwd_data_kernel<C, false, true>)
                  : (res ? nif_bwd_data_kernel<C, true, false> : nif_bwd_data_kernel<C, false, false>);
  NIF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0, occ = 0;
  NIF_CUDA_CHECK(cudaGetDevice(&dev));
  NIF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  NIF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, smem));
  if (occ < 1) occ = 1;
  long long grid = (long long)sms * occ;
  if (grid > a.total_tiles) grid = a.total_tiles;
  if (grid < 1) return NIF_OK;
  { NIF_PROF("nif_bwd_data_kernel", st); kern<<<(unsigned)grid, C::NT, smem, st>>>(pl, a); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// ---------------------------------------------------------------------------------------------------
// weight-gradient GEMM for the hidden matrices
// ---------------------------------------------------------------------------------------------------
struct WgtArgs {
  long long B, rows_per_split;
  int S;
  const float *z, *save, *da;
  float* part;  // [S][H][K+1][NP][NP]
  float zt_last;  // bias coordinate of zt (1; 0 when z is a latent tangent)
};

// KQ latent coordinates per thread: 4, or 1 for plans whose latent vector is just [1] (K = 0: the trunk)
template <int TS, int KQ>
__global__ void __launch_bounds__((TS / 4) * (TS / 4)) nif_bwd_weight_kernel(const Plan pl, const WgtArgs a) {
  constexpr int TT = TS / 4, NTW = TT * TT, RC = 32;  // RC rows per chunk
  constexpr int LD4 = (RC * TS / 4) / NTW;            // float4 loads per thread per operand per chunk
  __shared__ __align__(16) float hs[RC][TS];
  __shared__ __align__(16) float ds[RC][TS];
  __shared__ __align__(16) float zq[RC][4];
  const int NP = pl.NP, K = pl.K, K1 = pl.K + 1, H = pl.H + pl.wide_last;
  const int nb = NP / TS;
  const int KGN = (K1 + KQ - 1) / KQ;
  int bx = blockIdx.x;
  const int jb = bx % nb; bx /= nb;
  const int ib = bx % nb; bx /= nb;
  const int kg = bx % KGN; bx /= KGN;
  const int h = bx;  // hidden matrix index, layer m = h + 1
  const int s = blockIdx.y;
  const int tid = threadIdx.x, ti = tid / TT, tj = tid % TT;
  const float* hsrc = a.save + (long long)h * a.B * NP + ib * TS;        // h_m, m = h + 1 -> slot h
  const float* dsrc = a.da + (long long)(h + 1) * a.B * NP + jb * TS;    // da_m
  const long long r0 = (long long)s * a.rows_per_split;
  long long r1 = r0 + a.rows_per_split;
  if (r1 > a.B) r1 = a.B;

  float acc[KQ][4][4];
#pragma unroll
  for (int q = 0; q < KQ; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int f = 0; f < 4; ++f) acc[q][e][f] = 0.f;

  constexpr int ZL = (RC * 4 + NTW - 1) / NTW;     // latent values per thread per chunk
  float4 ph[LD4], pd[LD4];
  float pz[ZL];
  auto fetch = [&](long long rb) {
#pragma unroll
    for (int u = 0; u < LD4; ++u) {
      const int e = tid + u * NTW;  // float4 index in the chunk
      const int rr = e / (TS / 4), c4 = e % (TS / 4);
      const long long b = rb + rr;
      ph[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      pd[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < r1) {
        ph[u] = ldg4(&hsrc[b * NP + c4 * 4]);
        pd[u] = ldg4(&dsrc[b * NP + c4 * 4]);
      }
    }
#pragma unroll
    for (int u = 0; u < ZL; ++u) {
      const int e = tid + u * NTW;
      pz[u] = 0.f;
      if (e < RC * 4) {
        const int rr = e / 4, q = e % 4;
        const long long b = rb + rr;
        const int kk = kg * KQ + q;
        if (b < r1 && q < KQ) pz[u] = (kk < K) ? __ldg(&a.z[b * K + kk]) : (kk == K ? a.zt_last : 0.f);
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < LD4; ++u) {
      const int e = tid + u * NTW;
      const int rr = e / (TS / 4), c4 = e % (TS / 4);
      *reinterpret_cast<float4*>(&hs[rr][c4 * 4]) = ph[u];
      *reinterpret_cast<float4*>(&ds[rr][c4 * 4]) = pd[u];
    }
#pragma unroll
    for (int u = 0; u < ZL; ++u) {
      const int e = tid + u * NTW;
      if (e < RC * 4) zq[e / 4][e % 4] = pz[u];
    }
  };

  if (r0 < r1) fetch(r0);
  for (long long rb = r0; rb < r1; rb += RC) {
    __syncthreads();  // previous chunk fully consumed
    stash();
    __syncthreads();
    if (rb + RC < r1) fetch(rb + RC);
#pragma unroll 4
    for (int rr = 0; rr < RC; ++rr) {
      const float4 h4 = *reinterpret_cast<const float4*>(&hs[rr][ti * 4]);
      const float4 d4 = *reinterpret_cast<const float4*>(&ds[rr][tj * 4]);
      const float4 z4 = *reinterpret_cast<const float4*>(&zq[rr][0]);
      const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
      const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
      if (KQ == 1) {
        const float zd[4] = {zv[0] * dv[0], zv[0] * dv[1], zv[0] * dv[2], zv[0] * dv[3]};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) acc[0][e][f] = fmaf(hv[e], zd[f], acc[0][e][f]);
      } else {
        float o[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) o[e][f] = hv[e] * dv[f];
#pragma unroll
        for (int q = 0; q < KQ; ++q)
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[q][e][f] = fmaf(zv[q], o[e][f], acc[q][e][f]);
      }
    }
  }
  const float om = plan_omega(pl, h + 1);  // 1 for the last matrix (h == pl.H, wide_last)
#pragma unroll
  for (int q = 0; q < KQ; ++q) {
    const int kk = kg * KQ + q;
    if (kk < K1) {
      float* dst = a.part + ((((long long)s * H + h) * K1 + kk) * NP + ib * TS + ti * 4) * NP + jb * TS + tj * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        *reinterpret_cast<float4*>(&dst[(long long)e * NP]) =
            make_float4(om * acc[q][e][0], om * acc[q][e][1], om * acc[q][e][2], om * acc[q][e][3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// thin terms: bias rows of every layer, first matrix, last matrix
//   column space Q:  [0,(H+1)NP) dC_m[j] | so: dC_last[c] | si*NP: dM_0[i][j] | NP*so: dM_last[i][c]
// ---------------------------------------------------------------------------------------------------
struct EdgeArgs {
  long long B, rows_per_split;
  int S, Q;
  const float *z, *x, *save, *da, *du;
  float* part;  // [S][K+1][Q]
  int no_bias;  // tangent-adjoint pass: the bias-row columns are zero
  int tiled;    // da / save are in the tiled layout of the tensor-core paths (nif_common.cuh)
  int q_begin, q_end;  // columns handled by this launch (the tensor-core thin-term kernel takes the rest)
  float zt_last = 1.f; // bias coordinate of zt (0 when z is a latent tangent)
};

// Register-tiled batch reduction: a thread owns 4 columns x 4 latent coordinates; a CTA owns 64 columns x KG groups
// of 4 latent coordinates x RS row slices (16 * KG * RS threads).  Per chunk of 32 rows the features F[32][64]
// (generated on the fly from the stashed rows, coalesced along the columns) and the latent coordinates
// zt[32][4 KG] are staged in shared memory; the next chunk's global loads fly while the current one is consumed.
// Row slices (KG < 16) share the columns and split the rows of a chunk; they are summed through shared memory at
// the end, so the kernel always writes one partial per (batch split, kappa, column).
template <int KG, int RS>
__global__ void __launch_bounds__(16 * KG * RS) nif_bwd_edge_kernel(const Plan pl, const EdgeArgs a) {
  constexpr int NTH = 16 * KG * RS, RC = 32, KC = 4 * KG;
  constexpr int FL = NTH / 64;                        // fetch lanes: thread (lane fl, column c) fetches rows fl, fl+FL, ..
  constexpr int EPT = (RC + FL - 1) / FL;             // feature elements per fetch thread per chunk
  constexpr int ZPT = (RC * KC + NTH - 1) / NTH;      // latent coordinates per thread per chunk
  static_assert(NTH >= 64, "block too small");
  __shared__ __align__(16) float Fs[2][RC][64];  // double buffered: one barrier per chunk
  __shared__ __align__(16) float Zs[2][RC][KC];
  __shared__ __align__(16) float red[RS > 1 ? (RS - 1) * KG * 16 * 16 : 1];
  const int NP = pl.NP, K = pl.K, K1 = pl.K + 1, H = pl.H, si = pl.si, so = pl.so;
  const int tid = threadIdx.x;
  const int qg = tid % 16, kg = (tid / 16) % KG, rs = tid / (16 * KG);
  const int s = blockIdx.y;
  const int k0 = blockIdx.z * KC;  // first latent coordinate of this pass
  const long long r0 = (long long)s * a.rows_per_split;
  long long r1 = r0 + a.rows_per_split;
  if (r1 > a.B) r1 = a.B;

  // the feature column this thread fetches:  F[b] = scale * A[row(b)] * (Bp ? Bp[b*sB] : 1), row(b) = b * sA, or the
  // tiled row offset when the source is a tiled stash / da slot
  const int fc = tid % 64, fl = tid / 64;
  const float* Ap = nullptr;
  const float* Bp = nullptr;
  long long sA = 0, sB = 0;
  bool a_tiled = false;
  float scale = 1.f;
  {
    const long long slot = a.tiled ? nif_tiled_rows(a.B) * NP : a.B * NP;  // floats per da / stash slot
    int r = a.q_begin + blockIdx.x * 64 + fc;
    if (fl < FL && r < a.q_end) {
      if (r < (H + 1) * NP) {
        const int m = r / NP, j = r % NP;
        if (!a.no_bias) { Ap = a.da + (long long)m * slot + (a.tiled ? nif_tiled_col(j) : j); sA = NP; a_tiled = a.tiled; }
      } else if ((r -= (H + 1) * NP) < so) {
        if (!a.no_bias) { Ap = a.du + r; sA = so; }
      } else if ((r -= so) < si * NP) {
        const int i = r / NP, j = r % NP;
        Ap = a.da + (a.tiled ? nif_tiled_col(j) : j); sA = NP; a_tiled = a.tiled;
        Bp = a.x + i; sB = si; scale = plan_omega(pl, 0);
      } else {
        r -= si * NP;
        const int i = r / so, c = r % so;
        Ap = a.save + (long long)H * slot + (a.tiled ? nif_tiled_col(i) : i); sA = NP; a_tiled = a.tiled;
        Bp = a.du + c; sB = so;
      }
    }
  }
  float acc[4][4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f < 4; ++f) acc[e][f] = 0.f;

  // Loads are branch-free (clamped row, dummy-but-valid pointers) and their results are first touched when the
  // chunk is staged, so all of a chunk's loads are in flight while the previous chunk is consumed.
  const bool has_col = Ap != nullptr;
  const bool has_b = Bp != nullptr;
  if (!has_col) { Ap = a.da; sA = 0; }
  if (!has_b) { Bp = Ap; sB = 0; }
  const float* zsrc = K > 0 ? a.z : a.da;  // K == 0 (trunk plans): zt = [1], nothing is read from z
  float pa[EPT], pb[EPT], pz[ZPT];
  auto fetch = [&](long long rb) {
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
      long long b = rb + fl + u * FL;
      if (b >= r1) b = r0;
      pa[u] = __ldg(&Ap[a_tiled ? nif_tiled_row_np(b, NP) : b * sA]);
      pb[u] = __ldg(&Bp[b * sB]);
    }
#pragma unroll
    for (int u = 0; u < ZPT; ++u) {
      const int e = tid + u * NTH;
      const int kk = k0 + e % KC;
      long long b = rb + e / KC;
      if (b >= r1) b = r0;
      pz[u] = __ldg(&zsrc[K > 0 ? b * K + (kk < K ? kk : 0) : 0]);
    }
  };
  if (r0 < r1) fetch(r0);
  int buf = 0;
  for (long long rb = r0; rb < r1; rb += RC, buf ^= 1) {
    // buffer `buf` was last read two chunks ago; every thread has passed the barrier of the chunk in between
    if (fl < FL) {
#pragma unroll
      for (int u = 0; u < EPT; ++u) {
        const int rr = fl + u * FL;
        float v = scale * pa[u];
        if (has_b) v *= pb[u];
        if (rr < RC) Fs[buf][rr][fc] = (has_col && rb + rr < r1) ? v : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < ZPT; ++u) {
      const int e = tid + u * NTH;
      const int kk = k0 + e % KC;
      float v = (kk < K) ? pz[u] : (kk == K ? a.zt_last : 0.f);
      if (rb + e / KC >= r1) v = 0.f;
      if (e < RC * KC) Zs[buf][e / KC][e % KC] = v;
    }
    __syncthreads();
    if (rb + RC < r1) fetch(rb + RC);
#pragma unroll 8
    for (int rr = rs; rr < RC; rr += RS) {
      const float4 f4 = *reinterpret_cast<const float4*>(&Fs[buf][rr][4 * qg]);
      const float4 z4 = *reinterpret_cast<const float4*>(&Zs[buf][rr][4 * kg]);
      const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
      const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) acc[e][f] = fmaf(zv[e], fv[f], acc[e][f]);
    }
  }
  if (RS > 1) {  // sum the row slices
    __syncthreads();
    if (rs > 0) {
      float* dst = red + (((rs - 1) * KG + kg) * 16 + qg) * 16;
#pragma unroll
      for (int e = 0; e < 4; ++e) *reinterpret_cast<float4*>(dst + 4 * e) = make_float4(acc[e][0], acc[e][1], acc[e][2], acc[e][3]);
    }
    __syncthreads();
    if (rs > 0) return;
    for (int o = 1; o < RS; ++o) {
      const float* src = red + (((o - 1) * KG + kg) * 16 + qg) * 16;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 v = *reinterpret_cast<const float4*>(src + 4 * e);
        acc[e][0] += v.x; acc[e][1] += v.y; acc[e][2] += v.z; acc[e][3] += v.w;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int kk = k0 + 4 * kg + e;
    if (kk < K1) {
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int q = a.q_begin + blockIdx.x * 64 + 4 * qg + f;
        if (q < a.q_end) a.part[((long long)s * K1 + kk) * a.Q + q] = acc[e][f];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Keras 'mse' seed:  du = 2 (u - t) sw inv_gb / so ;  loss += sum_b sw[b] mean_c (u-t)^2 inv_gb
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nif_mse_seed_kernel(long long B, int so, const float* __restrict__ u,
                                                           const float* __restrict__ t, const float* __restrict__ sw,
                                                           float inv_gb, float* __restrict__ du,
                                                           float* __restrict__ part) {
  float local = 0.f;
  const float c2 = 2.f * inv_gb / so;
  for (long long b = blockIdx.x * 256LL + threadIdx.x; b < B; b += 256LL * gridDim.x) {
    const float w = sw ? sw[b] : 1.f;
    float e2 = 0.f;
    for (int c = 0; c < so; ++c) {
      const float e = u[b * so + c] - t[b * so + c];
      e2 = fmaf(e, e, e2);
      du[b * so + c] = c2 * w * e;
    }
    local = fmaf(w * inv_gb / so, e2, local);
  }
  __shared__ float red[256];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
// one warp, fixed summation order (lane-strided partial sums, then a shuffle tree): deterministic, and the loads of a
// lane are independent instead of one serial chain over every partial
__global__ void nif_loss_final_kernel(int nparts, const float* __restrict__ part, float* __restrict__ loss) {
  float sacc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 32) sacc += part[i];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
  if (threadIdx.x == 0) *loss += sacc;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int nif_unpack_grad_impl(const Plan& pl, int S_h, const float* part_h, int S_e, const float* part_e, int Q,
                         float* dw_h, float* db_h, float beta, cudaStream_t st);
struct TcBwdExt {  // hooks of the reverse-over-forward passes (nif_tc_bwd.cu)
  const float *h_stash, *e_stash, *ext_add;
  float* ext_out;
  int no_bias, dz_accumulate, ext_accumulate;
};
int nif_tc_bwd_data_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                         const float* save, const float* du, float* da, float* dz, unsigned* maxes,
                         cudaStream_t st, const TcBwdExt* ext = nullptr);
int nif_tc_bwd_weight_impl(const Plan& pl, long long B, const float* z, const float* save, const float* da,
                           const unsigned* maxes, int S, long long rows_per_split, float* part, cudaStream_t st);
int nif_tc_bwd_edge_impl(const Plan& pl, long long B, const float* z, const float* x, const float* save, const float* da,
                         const float* du, const unsigned* maxes, int S, long long rows_per_split, int Q, float* part,
                         cudaStream_t st, int no_bias = 0);


bool nif_plan_uses_bf(const Plan& pl);
int nif_bf_bwd_data_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                         const float* save, const float* du, float* da, float* dz, cudaStream_t st);
int nif_bf_bwd_weight_impl(const Plan& pl, long long B, const float* z, const float* save, const float* da, int S,
                           long long rows_per_split, float* part, cudaStream_t st);
int nif_bf_bwd_edge_impl(const Plan& pl, long long B, const float* z, const float* x, const float* save, const float* da,
                         const float* du, int S, long long rows_per_split, int Q, float* part, cudaStream_t st);
// batch splits of the bf16 thin-term kernel: (H + 1 + si + so + 1) feature blocks per split, whole waves of CTAs
static int nif_bf_edge_splits(const Plan& pl, long long B) {
  const long long items = pl.H + 1 + pl.si + pl.so + 1;
  long long s_max = (B + 511) / 512;
  if (s_max < 1) s_max = 1;
  if (s_max > 96) s_max = 96;
  long long best = 1;
  double best_eff = 0.0;
  for (long long S = 1; S <= s_max; ++S) {
    const long long ctas = items * S, waves = (ctas + 147) / 148;
    const double eff = (double)ctas / (double)(waves * 148);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = S; }
  }
  return (int)best;
}

static long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }
// batch splits of the bf16 weight-gradient kernel: one CTA per SM and every CTA costs the same, so the count that
// wastes the least of the last wave (at least 512 rows per split; an explicit acc_rows caps the rows per split)
static int nif_bf_wgt_splits(const Plan& pl, long long B) {
  const int KQ = 4 * (128 / pl.NP);
  const long long items = (long long)pl.H * ((pl.K + 1 + KQ - 1) / KQ);
  long long s_max = (B + 511) / 512;
  if (s_max < 1) s_max = 1;
  if (s_max > 64) s_max = 64;
  long long s_min = 1;
  if (pl.acc_rows > 0) {
    long long rows = pl.acc_rows < 256 ? 256 : pl.acc_rows;
    s_min = (B + rows - 1) / rows;
    if (s_max < s_min) s_max = s_min;
  }
  // every split writes (and the un-packing kernel reads) a full partial image: the smallest count that fills its last
  // wave to 90 %, else the best fill found
  long long best = s_min;
  double best_eff = 0.0;
  for (long long S = s_min; S <= s_max; ++S) {
    const long long ctas = items * S, waves = (ctas + 147) / 148;
    const double eff = (double)ctas / (double)(waves * 148);
    if (eff >= 0.9) return (int)S;
    if (eff > best_eff + 1e-9) { best_eff = eff; best = S; }
  }
  return (int)best;
}
// rows per batch split of the tensor-core batch-reduction kernels (bounds the truncating accumulation chains, see
// nif_grad_ws_layout): nif_desc_t.acc_rows, default 4096
static long long nif_tc_wgt_max_rows(const Plan& pl) {
  long long r = pl.acc_rows > 0 ? pl.acc_rows : 4096;
  if (r < 256) r = 256;
  return round_up(r, 64);
}
#define NIF_TC_WGT_MAX_ROWS nif_tc_wgt_max_rows(pl)
// batch splits of the tensor-core weight-gradient kernel: one CTA per SM and every CTA costs the same, so whole waves of
// CTAs -- a multiple of the splits that fill one wave -- and enough of them to respect the accumulation cap
static int nif_tc_wgt_splits(const Plan& pl, long long B) {
  const long long items = (long long)pl.H * ((pl.KP + 3) / 4);
  long long per_wave = items > 0 ? 148 / items : 1;
  if (per_wave < 1) per_wave = 1;
  const long long s_cap = (B + NIF_TC_WGT_MAX_ROWS - 1) / NIF_TC_WGT_MAX_ROWS;
  long long S = per_wave;
  if (S < s_cap) S = round_up(s_cap, per_wave);
  return (int)S;
}

// batch splits of the tensor-core thin-term kernel: whole waves of CTAs (column-block pairs x splits), and enough of
// them that no TMEM accumulation chain exceeds the cap (the same truncation mechanism as in the weight kernel)
static int nif_tc_edge_splits(const Plan& pl, long long B) {
  const int ncta_x = (pl.H + 1 + pl.si + pl.so + 1 + 1) / 2;
  long long per_wave = 148 / ncta_x;
  if (per_wave < 1) per_wave = 1;
  const long long s_cap = (B + NIF_TC_WGT_MAX_ROWS - 1) / NIF_TC_WGT_MAX_ROWS;
  long long S = per_wave;
  if (S < s_cap) S = round_up(s_cap, per_wave);
  const long long maxs = (B + 63) / 64;
  if (S > maxs) S = maxs;
  if (S < 1) S = 1;
  return (int)S;
}

// latent coordinates per pass of nif_bwd_edge_kernel (4 * KG of its instantiations)
static int nif_edge_kc(int K1) { return K1 <= 4 ? 4 : K1 <= 8 ? 8 : K1 <= 16 ? 16 : K1 <= 32 ? 32 : K1 <= 36 ? 36 : 68; }

GradWs nif_grad_ws_layout(const Plan& pl, long long B) {
  GradWs w;
  const long long NP = pl.NP, K1 = pl.K + 1, H = pl.H;
  w.Q = (int)((H + 1) * NP + pl.so + pl.si * NP + (pl.wide_last ? 0 : NP * pl.so));
  // batch splits: enough CTAs to fill the chip, but at least 256 rows per split
  const long long TS = NP >= 64 ? 64 : 32;
  const long long Hm = H + pl.wide_last;  // matrices handled by the hidden-matrix GEMM
  const long long kq = K1 == 1 ? 1 : 4;  // latent coordinates per thread of nif_bwd_weight_kernel
  const long long base_h = Hm * ((K1 + kq - 1) / kq) * (NP / TS) * (NP / TS);
  long long S_h = base_h > 0 ? (2 * 148 + base_h - 1) / base_h : 1;
  long long maxs = (B + 255) / 256;
  if (maxs < 1) maxs = 1;
  if (S_h > maxs) S_h = maxs;
  if (S_h < 1) S_h = 1;
  w.rows_h = round_up((B + S_h - 1) / S_h, 32);
  if (w.rows_h < 32) w.rows_h = 32;
  w.S_h = (int)((B + w.rows_h - 1) / w.rows_h);
  if (w.S_h < 1) w.S_h = 1;
  if (nif_plan_uses_tc(pl)) {
    // The tensor-core weight kernel keeps its accumulators in TMEM for a whole batch split, and the tensor core
    // truncates its fp32 accumulator on every instruction: an error that grows with the length of the chain.  Measured
    // against the fp64 oracle at 65 536 rows: dw 1.95e-5 / db 2.38e-5 with 16 384 rows per split, 4.8e-6 / 5.9e-6 with
    // 4096 (tools/emulate_batch_reduction.py reproduces both on the CPU).  At most NIF_TC_WGT_MAX_ROWS (4096) rows per
    // split, so the workspace must hold that many partials; they are added in fp32 round-to-nearest by the unpack kernel.
    const long long s_need = nif_tc_wgt_splits(pl, B);
    if (w.S_h < s_need && B > NIF_TC_WGT_MAX_ROWS) {
      w.S_h = (int)s_need;
      w.rows_h = round_up((B + s_need - 1) / s_need, 64);
    }
  }
  if (nif_plan_uses_bf(pl) && pl.K >= 1) {
    const long long S = nif_bf_wgt_splits(pl, B);
    w.rows_h = round_up((B + S - 1) / S, 64);
    if (w.rows_h < 64) w.rows_h = 64;
    w.S_h = (int)((B + w.rows_h - 1) / w.rows_h);
    if (w.S_h < 1) w.S_h = 1;
  }
  const long long kc = nif_edge_kc((int)K1);  // latent coordinates per pass of the thin-term kernel
  const long long base_e = (w.Q + 63) / 64 * ((K1 + kc - 1) / kc);
  long long S_e = (4 * 148 + base_e - 1) / base_e;
  if (S_e > maxs) S_e = maxs;
  if (S_e < 1) S_e = 1;
  w.rows_e = round_up((B + S_e - 1) / S_e, 64);
  if (w.rows_e < 64) w.rows_e = 64;
  w.S_e = (int)((B + w.rows_e - 1) / w.rows_e);
  if (w.S_e < 1) w.S_e = 1;
  w.S_e_ws = w.S_e;  // partials the workspace holds (S_e stays the CUDA-core kernel's split count)
  if (nif_plan_uses_bf(pl) && pl.K >= 1) {  // room for the bf16 thin-term kernel's splits
    const int s_bf = nif_bf_edge_splits(pl, B);
    if (s_bf > w.S_e_ws) w.S_e_ws = s_bf;
  }
  if (nif_plan_uses_tc(pl)) {  // room for the tensor-core thin-term kernel's splits (accumulation cap)
    const int s_tc = nif_tc_edge_splits(pl, B);
    if (s_tc > w.S_e_ws) w.S_e_ws = s_tc;
  }
  long long off = 0;
  w.da = off; off += round_up((H + 1 + pl.wide_last) * nif_tiled_rows(B) * NP, 4);  // (tiled slots hold B rounded up to 64 rows)
  w.du = off; off += round_up(B * pl.so, 4);
  w.part_h = off; off += round_up((long long)w.S_h * Hm * K1 * NP * NP, 4);
  w.part_e = off; off += round_up((long long)w.S_e_ws * K1 * w.Q, 4);
  w.loss_part = off; off += 1024;
  w.maxes = off; off += 256;  // device-side maxima used for tensor-core operand scales
  w.total = off;
  return w;
}

static cudaError_t launch_edge(const Plan& pl, const EdgeArgs& e, const GradWs& w, cudaStream_t st) {
  const int K1 = pl.K + 1;
  const int KC = nif_edge_kc(K1);
  dim3 grid((unsigned)((e.q_end - e.q_begin + 63) / 64), (unsigned)w.S_e, (unsigned)((K1 + KC - 1) / KC));
  switch (KC) {
    case 4: { NIF_PROF("nif_bwd_edge_kernel", st); nif_bwd_edge_kernel<1, 16><<<grid, 256, 0, st>>>(pl, e); } break;
    case 8: { NIF_PROF("nif_bwd_edge_kernel", st); nif_bwd_edge_kernel<2, 8><<<grid, 256, 0, st>>>(pl, e); } break;
    case 16: { NIF_PROF("nif_bwd_edge_kernel", st); nif_bwd_edge_kernel<4, 4><<<grid, 256, 0, st>>>(pl, e); } break;
    case 32: { NIF_PROF("nif_bwd_edge_kernel", st); nif_bwd_edge_kernel<8, 2><<<grid, 256, 0, st>>>(pl, e); } break;
    case 36: { NIF_PROF("nif_bwd_edge_kernel", st); nif_bwd_edge_kernel<9, 1><<<grid, 144, 0, st>>>(pl, e); } break;
    default: { NIF_PROF("nif_bwd_edge_kernel", st); nif_bwd_edge_kernel<17, 1><<<grid, 272, 0, st>>>(pl, e); } break;
  }
  return cudaGetLastError();
}

// hidden-matrix GEMM (CUDA cores) + thin terms + un-packing, for stashes whose da_m already sit in ws.
// Used by the trunk, whose parameter gradients are the same batch reductions with zt = [1].
int nif_weight_grads_impl(const Plan& pl, long long B, const float* z, const float* x, const float* save, const float* du,
                          float* dw_h, float* db_h, float beta, float* ws, cudaStream_t st, int no_bias = 0,
                          float zt_last = 1.f) {
  const GradWs w = nif_grad_ws_layout(pl, B);
  const int Hm = pl.H + pl.wide_last;
  if (Hm > 0) {
    WgtArgs g;
    g.B = B; g.rows_per_split = w.rows_h; g.S = w.S_h;
    g.z = z; g.save = save; g.da = ws + w.da; g.part = ws + w.part_h; g.zt_last = zt_last;
    const int K1 = pl.K + 1;
    const int kq = K1 == 1 ? 1 : 4;
    const int kgn = (K1 + kq - 1) / kq;
    if (pl.NP >= 64) {
      const int nb = pl.NP / 64;
      dim3 grid((unsigned)(Hm * kgn * nb * nb), (unsigned)w.S_h);
      if (kq == 1) { NIF_PROF("nif_bwd_weight_kernel", st); nif_bwd_weight_kernel<64, 1><<<grid, 256, 0, st>>>(pl, g); }
      else { NIF_PROF("nif_bwd_weight_kernel", st); nif_bwd_weight_kernel<64, 4><<<grid, 256, 0, st>>>(pl, g); }
    } else {
      dim3 grid((unsigned)(Hm * kgn), (unsigned)w.S_h);
      if (kq == 1) { NIF_PROF("nif_bwd_weight_kernel", st); nif_bwd_weight_kernel<32, 1><<<grid, 64, 0, st>>>(pl, g); }
      else { NIF_PROF("nif_bwd_weight_kernel", st); nif_bwd_weight_kernel<32, 4><<<grid, 64, 0, st>>>(pl, g); }
    }
    NIF_CUDA_CHECK(cudaGetLastError());
  }
  EdgeArgs e;
  e.B = B; e.rows_per_split = w.rows_e; e.S = w.S_e; e.Q = w.Q;
  e.z = z; e.x = x; e.save = save; e.da = ws + w.da; e.du = du; e.part = ws + w.part_e; e.no_bias = no_bias; e.tiled = 0;
  e.q_begin = 0; e.q_end = w.Q; e.zt_last = zt_last;
  NIF_CUDA_CHECK(launch_edge(pl, e, w, st));
  return nif_unpack_grad_impl(pl, w.S_h, ws + w.part_h, w.S_e, ws + w.part_e, w.Q, dw_h, db_h, beta, st);
}

// Batch reductions of a tensor-core reverse pass whose da_m and maxima sit in ws: hidden-matrix GEMM, thin terms,
// un-packing.  hsave: the stash whose h_m slots pair with da_m (the primal stash, or the tangent activations h'_m of a
// reverse-over-forward pass, then with no_bias = 1 and x = xdot).
static int nif_tc_weight_grads(const Plan& pl, long long B, const float* z, const float* x, const float* hsave,
                               const float* du, float* dw_h, float* db_h, float beta, float* ws, cudaStream_t st,
                               int no_bias) {
  const GradWs w = nif_grad_ws_layout(pl, B);
  int S_used = w.S_h, S_e_used = w.S_e;
  {
    // one CTA per SM: as many batch splits as fit one wave (never more than the workspace was sized for)
    int S = nif_tc_wgt_splits(pl, B);  // whole waves, bounded accumulation chains (see nif_grad_ws_layout)
    if (S > w.S_h) S = w.S_h;
    long long rows = round_up((B + S - 1) / S, 64);
    S = (int)((B + rows - 1) / rows);
    const int rcw = nif_tc_bwd_weight_impl(pl, B, z, hsave, ws + w.da, reinterpret_cast<const unsigned*>(ws + w.maxes),
                                           S, rows, ws + w.part_h, st);
    if (rcw != NIF_OK) return rcw;
    S_used = S;
  }
  {
    EdgeArgs e;
    e.B = B; e.rows_per_split = w.rows_e; e.S = w.S_e; e.Q = w.Q;
    e.z = z; e.x = x; e.save = hsave; e.da = ws + w.da; e.du = du; e.part = ws + w.part_e; e.no_bias = no_bias; e.tiled = 1;
    e.q_begin = 0; e.q_end = w.Q;
    // thin terms on the tensor cores: one wave of CTAs (column-block pairs x batch splits)
    int S_tc = nif_tc_edge_splits(pl, B);
    if (S_tc > w.S_e_ws) S_tc = w.S_e_ws;
    const long long rows_tc = round_up((B + S_tc - 1) / S_tc, 64);
    S_tc = (int)((B + rows_tc - 1) / rows_tc);
    const int rce = nif_tc_bwd_edge_impl(pl, B, z, x, hsave, ws + w.da, du, reinterpret_cast<const unsigned*>(ws + w.maxes),
                                         S_tc, rows_tc, w.Q, ws + w.part_e, st, no_bias);
    if (rce == NIF_OK) S_e_used = S_tc;
    else if (rce != NIF_E_UNSUPPORTED) return rce;
    else NIF_CUDA_CHECK(launch_edge(pl, e, w, st));
  }
  return nif_unpack_grad_impl(pl, S_used, ws + w.part_h, S_e_used, ws + w.part_e, w.Q, dw_h, db_h, beta, st);
}

// dz_ev (optional): recorded on `st` as soon as the data pass has been enqueued, i.e. when dz is final: the caller may start
// the ParameterNet trunk's reverse pass on another stream while the weight-gradient kernels of this call run
int nif_backward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                      const float* save, const float* du, float* dw_h, float* db_h, float beta, float* dz,
                      float* ws, cudaStream_t st, cudaEvent_t dz_ev) {
  const GradWs w = nif_grad_ws_layout(pl, B);
  if (B <= 0) return NIF_OK;
  BwdArgs a;
  a.B = B;
  a.z = z; a.x = x; a.packed = packed; a.save = save; a.du = du;
  a.da = ws + w.da;
  a.dz = dz;
  a.h_stash = save; a.e_stash = nullptr; a.ext_out = nullptr; a.ext_add = nullptr; a.no_bias = 0; a.dz_accumulate = 0;
  a.ext_accumulate = 0; a.da_in = nullptr; a.dh_out = nullptr; a.dh_add = nullptr; a.zt_last = 1.f;
  if (pl.bf) {  // bf16 tensor-core reverse pass: data kernel, weight-gradient GEMM, thin terms on the CUDA cores
    int rcb = nif_bf_bwd_data_impl(pl, B, z, x, packed, save, du, ws + w.da, dz, st);
    if (rcb == NIF_OK) {
      if (dz_ev) { NIF_CUDA_CHECK(cudaEventRecord(dz_ev, st)); dz_ev = nullptr; }
      rcb = nif_bf_bwd_weight_impl(pl, B, z, save, ws + w.da, w.S_h, w.rows_h, ws + w.part_h, st);
      if (rcb != NIF_OK) return rcb;
      int S_e_used = w.S_e;
      {  // thin terms: tensor-core batch reduction; shapes it does not cover use the CUDA-core kernel
        int S_bf = nif_bf_edge_splits(pl, B);
        if (S_bf > w.S_e_ws) S_bf = w.S_e_ws;
        const long long rows_bf = round_up((B + S_bf - 1) / S_bf, 64);
        S_bf = (int)((B + rows_bf - 1) / rows_bf);
        rcb = nif_bf_bwd_edge_impl(pl, B, z, x, save, ws + w.da, du, S_bf, rows_bf, w.Q, ws + w.part_e, st);
        if (rcb == NIF_OK) S_e_used = S_bf;
        else if (rcb != NIF_E_UNSUPPORTED) return rcb;
        else {
          EdgeArgs e;
          e.B = B; e.rows_per_split = w.rows_e; e.S = w.S_e; e.Q = w.Q;
          e.z = z; e.x = x; e.save = save; e.da = ws + w.da; e.du = du; e.part = ws + w.part_e; e.no_bias = 0; e.tiled = 1;
          e.q_begin = 0; e.q_end = w.Q;
          NIF_CUDA_CHECK(launch_edge(pl, e, w, st));
        }
      }
      return nif_unpack_grad_impl(pl, w.S_h, ws + w.part_h, S_e_used, ws + w.part_e, w.Q, dw_h, db_h, beta, st);
    }
    if (rcb != NIF_E_UNSUPPORTED) return rcb;
  }
  int rc = NIF_E_UNSUPPORTED;
  if (pl.tc)  // tensor-core data pass; shapes it does not cover use the CUDA-core kernel below
    rc = nif_tc_bwd_data_impl(pl, B, z, x, packed, save, du, ws + w.da, dz, reinterpret_cast<unsigned*>(ws + w.maxes), st);
  const bool tc_data = (rc == NIF_OK);
  if (rc == NIF_E_UNSUPPORTED)
  switch (pl.NP) {
    case 32: a.total_tiles = (B + Cfg32::TB - 1) / Cfg32::TB; rc = launch_bwd_data<Cfg32>(pl, a, st); break;
    case 64: a.total_tiles = (B + Cfg64::TB - 1) / Cfg64::TB; rc = launch_bwd_data<Cfg64>(pl, a, st); break;
    case 128: a.total_tiles = (B + Cfg128::TB - 1) / Cfg128::TB; rc = launch_bwd_data<Cfg128>(pl, a, st); break;
    default: nif_set_error("unsupported padded width %d", pl.NP); return NIF_E_UNSUPPORTED;
  }
  if (rc != NIF_OK) return rc;
  if (dz_ev) NIF_CUDA_CHECK(cudaEventRecord(dz_ev, st));

  if (tc_data && pl.H > 0)  // tensor-core batch reductions (they need the maxima the TC data pass recorded)
    return nif_tc_weight_grads(pl, B, z, x, save, du, dw_h, db_h, beta, ws, st, 0);
  return nif_weight_grads_impl(pl, B, z, x, save, du, dw_h, db_h, beta, ws, st);  // CUDA cores
}

// ---------------------------------------------------------------------------------------------------
// Reverse-over-forward pass for Sobolev training (JacobianLayer inside the loss, tutorial/8 cell 20;
// nif/layers/gradient.py:207-231 differentiated once more by Keras' tape).  One tangent direction xdot on the ShapeNet
// inputs (z' = 0).  With a_m the pre-activations, a_m' their tangents, and seeds du = dL/du, dud = dL/du':
//   tangent adjoint:  da_m' = dh'_{m+1} * d_m                                   (same weights, same d_m)
//   primal adjoint :  da_m  = dh_{m+1}  * d_m + dh'_{m+1} * e_m,   e_m = alpha act''(a_m) a_m'
//   dW_m = omega (h_m (x) da_m + h_m' (x) da_m'),  dC_m = da_m,  dz = dz(primal pass) + dz(tangent pass, no bias rows)
// i.e. two passes of the plain reverse machinery: first the tangent adjoint (stash slots h', inputs xdot, bias rows
// dropped; it leaves X_m = dh'_{m+1} * e_m in the workspace), then the primal adjoint (adds X_m), each followed by the
// batch-reduction kernels; the second pass accumulates.  Runs on the fp32 CUDA-core kernels.
// ws: nif_grad_ws_layout(pl, B).total + (H+1) * B * NP floats.
// ---------------------------------------------------------------------------------------------------
int nif_sobolev_backward_impl(const Plan& pl, long long B, const float* z, const float* x, int n_dir, const float* zdot,
                              unsigned zdot_dirs, const float* xdot, const float* packed, const float* save, const float* du,
                              const float* dud, float* dw_h, float* db_h, float beta, float* dz, float* dzdot, float* ws,
                              cudaStream_t st) {
  // n_dir directions (xdot [n_dir][B][si], zdot [n_dir][B][K] or null = the latent code does not move, dud
  // [n_dir][B][so], dzdot [n_dir][B][K] out).  save: slots [0, 2(H+1)) h, d; direction d: [2(H+1)(1+d), ...) h', e.
  // Per direction: the tangent adjoint (chain through G = sum_k zt_k w h' M_k), then -- when zdot moves -- the local pass
  // over F = sum_k zdot_k (w h M_k + C_k) with the adjoints the first pass left in ws.da; last the primal adjoint,
  // which picks up X (through act'') and S (through F's h_m).
  if (B <= 0) return NIF_OK;
  const GradWs w = nif_grad_ws_layout(pl, B);
  if (!zdot && nif_plan_tc_sobolev(pl)) {
    // Tensor-core form (ShapeNet-input directions only; the stash is tiled, written by the same branch of
    // nif_tangent_impl): per direction the tangent adjoint -- nif_tc_bwd_data_kernel<ext> with h' as activations, xdot
    // as input and the bias rows dropped; it leaves X_m = dh'_{m+1} e_m in the workspace -- then the primal adjoint,
    // which adds X_m to da_m; each followed by the FP16x3 batch reductions.
    const long long tslot = nif_tiled_rows(B) * 64;
    float* Xt = ws + w.total;
    unsigned* maxes = reinterpret_cast<unsigned*>(ws + w.maxes);
    for (int d = 0; d <= n_dir; ++d) {
      const bool primal = d == n_dir;
      TcBwdExt e;
      const float *xin, *seed;
      if (!primal) {
        xin = xdot + (long long)d * B * pl.si; seed = dud + (long long)d * B * pl.so;
        e.h_stash = save + (2LL + 2 * d) * (pl.H + 1) * tslot; e.e_stash = e.h_stash + (pl.H + 1) * tslot;
        e.ext_out = Xt; e.ext_add = nullptr; e.ext_accumulate = d > 0; e.no_bias = 1; e.dz_accumulate = d > 0;
      } else {
        xin = x; seed = du;
        e.h_stash = save; e.e_stash = nullptr; e.ext_out = nullptr; e.ext_add = Xt; e.ext_accumulate = 0; e.no_bias = 0;
        e.dz_accumulate = 1;
      }
      int rc = nif_tc_bwd_data_impl(pl, B, z, xin, packed, save, seed, ws + w.da, dz, maxes, st, &e);
      if (rc != NIF_OK) return rc;
      rc = nif_tc_weight_grads(pl, B, z, xin, e.h_stash, seed, dw_h, db_h, d == 0 ? beta : 1.0f, ws, st, e.no_bias);
      if (rc != NIF_OK) return rc;
    }
    return NIF_OK;
  }
  const long long slot = B * (long long)pl.NP;
  float* X = ws + w.total;
  float* S = X + (pl.H + 1) * slot;
  // zdot_dirs: bit d set = direction d moves the latent code (its zdot rows are not identically zero); the others skip the
  // local pass and get dzdot = 0
  if (zdot) {
    NIF_CUDA_CHECK(cudaMemsetAsync(S, 0, sizeof(float) * (pl.H + 1) * slot, st));
    for (int d = 0; d < n_dir; ++d)
      if (!((zdot_dirs >> d) & 1u)) NIF_CUDA_CHECK(cudaMemsetAsync(dzdot + (long long)d * B * pl.K, 0, sizeof(float) * B * pl.K, st));
  }
  auto run = [&](BwdArgs& a) -> int {
    switch (pl.NP) {
      case 32: a.total_tiles = (B + Cfg32::TB - 1) / Cfg32::TB; return launch_bwd_data<Cfg32>(pl, a, st);
      case 64: a.total_tiles = (B + Cfg64::TB - 1) / Cfg64::TB; return launch_bwd_data<Cfg64>(pl, a, st);
      case 128: a.total_tiles = (B + Cfg128::TB - 1) / Cfg128::TB; return launch_bwd_data<Cfg128>(pl, a, st);
    }
    nif_set_error("unsupported padded width %d", pl.NP);
    return NIF_E_UNSUPPORTED;
  };
  for (int d = 0; d <= n_dir; ++d) {
    const bool primal = d == n_dir;
    BwdArgs a;
    a.B = B;
    a.packed = packed; a.save = save;
    a.da = ws + w.da;
    a.ext_accumulate = 0; a.da_in = nullptr; a.dh_out = nullptr; a.dh_add = nullptr; a.zt_last = 1.f;
    a.z = z; a.dz = dz;
    if (!primal) {  // tangent adjoint of direction d
      a.x = xdot + (long long)d * B * pl.si; a.du = dud + (long long)d * B * pl.so;
      a.h_stash = save + (2LL + 2 * d) * (pl.H + 1) * slot; a.e_stash = a.h_stash + (pl.H + 1) * slot;
      a.ext_out = X; a.ext_add = nullptr; a.ext_accumulate = d > 0;
      a.no_bias = 1; a.dz_accumulate = d > 0;
    } else {        // primal adjoint
      a.x = x; a.du = du; a.h_stash = save; a.e_stash = nullptr; a.ext_out = nullptr; a.ext_add = X;
      a.dh_add = zdot ? S : nullptr;
      a.no_bias = 0; a.dz_accumulate = 1;
    }
    int rc = run(a);
    if (rc != NIF_OK) return rc;
    rc = nif_weight_grads_impl(pl, B, z, a.x, a.h_stash, a.du, dw_h, db_h, d == 0 ? beta : 1.0f, ws, st, a.no_bias);
    if (rc != NIF_OK) return rc;
    if (!primal && zdot && ((zdot_dirs >> d) & 1u)) {  // local pass: ws.da still holds da'_m of this direction
      const float* zd = zdot + (long long)d * B * pl.K;
      BwdArgs l = a;
      l.z = zd; l.zt_last = 0.f; l.x = x; l.h_stash = save; l.e_stash = nullptr; l.ext_out = nullptr; l.ext_add = nullptr;
      l.no_bias = 0; l.dz = dzdot + (long long)d * B * pl.K; l.dz_accumulate = 0;
      l.da_in = ws + w.da; l.dh_out = S;
      rc = run(l);
      if (rc != NIF_OK) return rc;
      rc = nif_weight_grads_impl(pl, B, zd, x, save, a.du, dw_h, db_h, 1.0f, ws, st, 0, 0.f);
      if (rc != NIF_OK) return rc;
    }
  }
  return NIF_OK;
}

int nif_mse_backward_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed,
                          const float* u, const float* save, const float* target, const float* sw, float inv_gb,
                          float* loss, float* dw_h, float* db_h, float beta, float* dz, float* ws,
                          cudaStream_t st, cudaEvent_t dz_ev) {
  const GradWs w = nif_grad_ws_layout(pl, B);
  if (B <= 0) return NIF_OK;
  int nblk = (int)((B + 255) / 256);
  if (nblk > 1024) nblk = 1024;
  { NIF_PROF("nif_mse_seed_kernel", st); nif_mse_seed_kernel<<<nblk, 256, 0, st>>>(B, pl.so, u, target, sw, inv_gb, ws + w.du, ws + w.loss_part); }
  NIF_CUDA_CHECK(cudaGetLastError());
  { NIF_PROF("nif_loss_final_kernel", st); nif_loss_final_kernel<<<1, 32, 0, st>>>(nblk, ws + w.loss_part, loss); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return nif_backward_impl(pl, B, z, x, packed, save, ws + w.du, dw_h, db_h, beta, dz, ws, st, dz_ev);
}
