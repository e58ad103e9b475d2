// Weight image of the bf16 tensor-core path: the reference's pnet_output column slices (nif/model.py:253-300, 769-846,
// 883-933) applied once to [w_h; b_h] and written as bf16 UMMA operand tiles (layout: nif_bf.cuh).
#include "nif_bf.cuh"

__device__ __forceinline__ float bf_src_at(const Plan& pl, const float* __restrict__ w_h, const float* __restrict__ b_h,
                                           long long g, int kappa, int col) {
  return (kappa < pl.K) ? w_h[(long long)kappa * pl.P + col] : b_h[g * pl.P + col];
}

// logical value of small tile T at (row, col)
__device__ __forceinline__ float bf_small_value(const Plan& pl, const float* __restrict__ w_h, const float* __restrict__ b_h,
                                                long long g, int T, int row, int col) {
  const int K1 = pl.K + 1, n = pl.n, si = pl.si, so = pl.so, H = pl.H;
  if (T < bf_n_jk(pl)) {  // [NP rows x KZ]: col = kappa
    if (col >= K1 || row >= n) return 0.f;
    if (T < si) return bf_src_at(pl, w_h, b_h, g, col, T * n + row);                       // X0[i']: M0[kappa][i'][j]
    if (T <= si + H) return bf_src_at(pl, w_h, b_h, g, col, plan_b_off(pl, T - si) + row);  // C_m[kappa][j], m = T - si
    const int c = T - (si + 1 + H);                                                         // BLT[c]: ML[kappa][i][c]
    return bf_src_at(pl, w_h, b_h, g, col, plan_w_off(pl, H + 1) + row * so + c);
  }
  const int Tk = T - bf_n_jk(pl);  // [KZ rows x NP]: row = kappa
  if (row >= K1 || col >= n) return 0.f;
  if (Tk < so) return bf_src_at(pl, w_h, b_h, g, row, plan_w_off(pl, H + 1) + col * so + Tk);  // XL[c]: ML[kappa][i][c]
  if (Tk < so + H + 1) return bf_src_at(pl, w_h, b_h, g, row, plan_b_off(pl, Tk - so) + col);   // BC[m]: C_m[kappa][j]
  return bf_src_at(pl, w_h, b_h, g, row, (Tk - so - H - 1) * n + col);                          // B0[i]: M0[kappa][i][j]
}

// one thread per float slot (two bf16 along K)
__global__ void __launch_bounds__(256) nif_pack_bf_kernel(const Plan pl, long long G, const float* __restrict__ w_h,
                                                          const float* __restrict__ b_h, float* __restrict__ packed) {
  const long long per = pl.packed_floats - pl.off_WF;  // W floats of one image (including tail padding)
  const long long total = G * per;
  const int NP = pl.NP, n = pl.n, K1 = pl.K + 1, CK = bf_ck(pl), NCHW = bf_nchw(pl), KZ = pl.KZ;
  const long long n_main = (long long)pl.H * NCHW * bf_chunk_floats(pl);
  const long long n_small = (long long)(bf_n_jk(pl) + bf_n_kj(pl)) * bf_small_floats(pl);
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    const long long g = e / per;
    long long r = e - g * per;
    float a0 = 0.f, a1 = 0.f;
    if (r < 2 * n_main) {
      const bool fwd = r < n_main;
      if (!fwd) r -= n_main;
      int t = (int)(r % bf_chunk_floats(pl));
      r /= bf_chunk_floats(pl);
      const int ck = (int)(r % NCHW), h = (int)(r / NCHW);
      // slot t of a [128 x NP] K-major tile: row group of 8 rows = (NP/8) k-groups x 32 slots
      const int per_rg = (NP / 8) * 32;
      const int rg = t / per_rg; t %= per_rg;
      const int kc = t / 32; t %= 32;
      const int row = rg * 8 + t / 4, k = kc * 8 + (t % 4) * 2;
      const int kl = row / NP, a = row % NP;
      const int kk = ck * CK + kl;
      if (kk < K1 && a < n) {
        // forward rows are j (K = i); reverse rows are i (K = j)
        const int col0 = plan_w_off(pl, h + 1);
        if (fwd) {
          if (k < n) a0 = bf_src_at(pl, w_h, b_h, g, kk, col0 + k * n + a);
          if (k + 1 < n) a1 = bf_src_at(pl, w_h, b_h, g, kk, col0 + (k + 1) * n + a);
        } else {
          if (k < n) a0 = bf_src_at(pl, w_h, b_h, g, kk, col0 + a * n + k);
          if (k + 1 < n) a1 = bf_src_at(pl, w_h, b_h, g, kk, col0 + a * n + k + 1);
        }
      }
    } else if (r < 2 * n_main + n_small) {
      r -= 2 * n_main;
      const int T = (int)(r / bf_small_floats(pl));
      int t = (int)(r % bf_small_floats(pl));
      const int KD = T < bf_n_jk(pl) ? KZ : NP;  // K extent of this tile
      const int per_rg = (KD / 8) * 32;
      const int rg = t / per_rg; t %= per_rg;
      const int kc = t / 32; t %= 32;
      const int row = rg * 8 + t / 4, k = kc * 8 + (t % 4) * 2;
      a0 = bf_small_value(pl, w_h, b_h, g, T, row, k);
      a1 = bf_small_value(pl, w_h, b_h, g, T, row, k + 1);
    }
    packed[g * pl.packed_floats + pl.off_WF + (e - g * per)] = __uint_as_float(bf_pack2(a0, a1));
  }
}

int nif_pack_bf_impl(const Plan& pl, long long G, const float* w_h, const float* b_h, float* packed, cudaStream_t st) {
  const long long total = G * (pl.packed_floats - pl.off_WF);
  long long nblk = (total + 255) / 256;
  if (nblk > 148 * 32) nblk = 148 * 32;
  { NIF_PROF("nif_pack_bf_kernel", st); nif_pack_bf_kernel<<<(unsigned)nblk, 256, 0, st>>>(pl, G, w_h, b_h, packed); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
