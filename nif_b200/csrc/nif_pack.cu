// Plan construction, weight packing (reference column layout -> kernel layout) and gradient
// un-packing (kernel partials -> reference layout).
//
// The reference slices pnet_output[:, a:b] into per-sample matrices (nif/model.py:253-300,
// 769-846, 883-933).  Here the same offsets are applied ONCE to the shared [K,P] / [P] parameters.
#include <cstdarg>
#include <cstdio>
#include <cuda_fp16.h>
#include "nif_common.cuh"

static thread_local char g_err[512] = "";
void nif_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* nif_last_error(void) { return g_err; }
extern "C" int nif_version(void) { return 100; }

static long long up4(long long v) { return (v + 3) / 4 * 4; }

// offsets of the packed-image sections for a plan whose shape fields (variant .. NP, P, wide_last, tc) are set
void nif_plan_layout(Plan* pp) {
  Plan& p = *pp;
  const long long K1 = p.K + 1, NP = p.NP;
  const long long HT = p.H + p.wide_last;
  long long off = 0;
  p.off_MH = off;  off += up4(HT * K1 * NP * NP);
  p.off_MHT = off; off += up4(HT * K1 * NP * NP);
  p.off_M0 = off;  off += up4(K1 * p.si * NP);
  p.off_ML = off;  off += up4(K1 * NP * p.so);
  p.off_C = off;   off += up4((long long)p.Lm * K1 * NP);
  p.KP = (p.K + 2) / 2 * 2;
  p.NCH = p.KP / 2;
  p.KZ = (p.K + 1 + 15) / 16 * 16;
  p.LPC = (2 * p.KZ <= 128) ? 2 : 1;
  p.NLC = (p.so + p.LPC - 1) / p.LPC;
  p.off_TCF = p.off_TCB = p.off_TCS = p.off_TCX = p.off_TCS2 = 0;
  if (p.tc) {
    off = (off + 31) / 32 * 32;
    p.off_TCF = off; off += (long long)p.H * p.NCH * NIF_TC_CHUNK_FLOATS;
    p.off_TCB = off; off += (long long)p.H * p.NCH * NIF_TC_CHUNK_FLOATS;
    p.off_TCS = off; off += up4((long long)p.H * p.KP);
    p.off_TCX = off; off += (p.si + 1 + p.H) * plan_x0_floats(p) + p.NLC * plan_xl_floats(p) + (p.H + 1 + p.si) * plan_x0_floats(p);
    p.off_TCS2 = off; off += up4(plan_n_small(p));
  }
  p.off_WF = p.off_WB = p.off_WX = 0;
  if (p.bf) {
    off = (off + 255) / 256 * 256;  // bulk-copy sources: 1 KB aligned
    const long long nchw = (p.K + 1 + (128 / p.NP) - 1) / (128 / p.NP);
    p.off_WF = off; off += (long long)p.H * nchw * 64 * p.NP;
    p.off_WB = off; off += (long long)p.H * nchw * 64 * p.NP;
    p.off_WX = off; off += (long long)(2 * (p.si + p.so + p.H + 1)) * p.NP * p.KZ / 2;
  }
  p.packed_floats = (off + 255) / 256 * 256;  // keep every image 1 KB aligned
}

int nif_make_plan(const nif_desc_t* d, Plan* out) {
  if (!d) { nif_set_error("null descriptor"); return NIF_E_BAD_DESC; }
  if (d->variant < 0 || d->variant > 2) { nif_set_error("variant %d not in {0,1,2}", d->variant); return NIF_E_BAD_DESC; }
  if (d->si < 1 || d->si > NIF_MAX_SI) { nif_set_error("si=%d outside [1,%d]", d->si, NIF_MAX_SI); return NIF_E_BAD_DESC; }
  if (d->so < 1 || d->so > NIF_MAX_SO) { nif_set_error("so=%d outside [1,%d]", d->so, NIF_MAX_SO); return NIF_E_BAD_DESC; }
  if (d->n < 1 || d->n > 128) { nif_set_error("units=%d outside [1,128]", d->n); return NIF_E_BAD_DESC; }
  if (d->l < 0 || d->l > 64) { nif_set_error("nlayers=%d outside [0,64]", d->l); return NIF_E_BAD_DESC; }
  if (d->K < 0 || d->K > 256) { nif_set_error("latent_dim=%d outside [0,256]", d->K); return NIF_E_BAD_DESC; }
  if (d->act < 0 || d->act > NIF_ACT_SIGMOID) { nif_set_error("activation id %d unknown", d->act); return NIF_E_BAD_DESC; }
  if (d->dtype_compute < 0 || d->dtype_compute > 2) {
    nif_set_error("dtype_compute=%d: built paths are 0 (fp32 CUDA cores), 1 (tensor cores, bf16 operands) and 2 (tensor cores, FP16x3 split)", d->dtype_compute);
    return NIF_E_UNSUPPORTED;
  }
  if (d->dtype_compute == 1 && d->n <= 32) {
    nif_set_error("dtype_compute=1 (bf16 tensor cores) is built for units > 32 (got %d)", d->n);
    return NIF_E_UNSUPPORTED;
  }
  if (d->dtype_compute == 2 && (d->n <= 32 || d->n > 64)) {
    nif_set_error("dtype_compute=2 (tensor cores) is built for 32 < units <= 64 only (got %d)", d->n);
    return NIF_E_UNSUPPORTED;
  }
  Plan p;
  p.variant = d->variant;
  p.act = d->variant == NIF_VARIANT_NIF ? d->act : NIF_ACT_SINE;
  p.si = d->si; p.so = d->so; p.n = d->n; p.l = d->l; p.K = d->K;
  p.omega0 = d->variant == NIF_VARIANT_NIF ? 1.0f : d->omega0;
  p.H = d->variant == NIF_VARIANT_SIREN_RES ? 2 * d->l : d->l;
  p.Lm = p.H + 2;
  p.NP = d->n <= 32 ? 32 : (d->n <= 64 ? 64 : 128);
  p.P = p.H * p.n * p.n + (p.si + p.so + 1 + p.H) * p.n + p.so;
  p.wide_last = 0;
  p.tc = d->dtype_compute == 2 ? 1 : 0;
  p.bf = d->dtype_compute == 1 ? 1 : 0;
  if (d->acc_rows < 0) { nif_set_error("acc_rows=%d is negative", d->acc_rows); return NIF_E_BAD_DESC; }
  p.acc_rows = d->acc_rows;
  nif_plan_layout(&p);
  *out = p;
  return NIF_OK;
}

// value of [w_h; b_h] at (kappa, reference column col) for group g
__device__ __forceinline__ float src_at(const Plan& pl, const float* __restrict__ w_h, const float* __restrict__ b_h,
                                        long long g, int kappa, int col) {
  return (kappa < pl.K) ? w_h[(long long)kappa * pl.P + col] : b_h[g * pl.P + col];
}

// logical value of small tile T at (row, col); see Plan for the tile list
__device__ __forceinline__ float tcx_value(const Plan& pl, const float* __restrict__ w_h, const float* __restrict__ b_h,
                                           int T, int row, int col) {
  const int K1 = pl.K + 1, n = pl.n;
  if (T <= pl.si + pl.H) {  // X0[i'] / XC[m]: row = j, col = kappa
    if (col >= K1 || row >= n) return 0.f;
    if (T < pl.si) return src_at(pl, w_h, b_h, 0, col, T * n + row);
    const int m = T - pl.si;  // 0 (bias of the first layer) .. H
    return src_at(pl, w_h, b_h, 0, col, plan_b_off(pl, m) + row);
  }
  const int q = T - (pl.si + 1 + pl.H);  // XL[q]: row = c_l * KZ + kappa, col = i
  if (q < pl.NLC) {
    const int cl = row / pl.KZ, kk = row % pl.KZ, c = pl.LPC * q + cl;
    if (kk >= K1 || c >= pl.so || col >= n) return 0.f;
    return src_at(pl, w_h, b_h, 0, kk, plan_w_off(pl, pl.H + 1) + col * pl.so + c);
  }
  // reverse-pass tiles, row = kappa, col = j:  BCt[m] = C_m[kappa][j],  B0t[i] = M0[kappa][i][j]
  if (row >= K1 || col >= n) return 0.f;
  const int t2 = q - pl.NLC;
  if (t2 <= pl.H) return src_at(pl, w_h, b_h, 0, row, plan_b_off(pl, t2) + col);
  return src_at(pl, w_h, b_h, 0, row, (t2 - pl.H - 1) * n + col);
}

// The thin sections of the image (M0, ML, C, the small tensor-core tiles): one thread per float slot.  The hidden
// matrices (MH, MHT and the main tensor-core tiles) are written by nif_pack_mats_kernel, one block per matrix.
// Slots visited per group: [off_M0, end1) and [off_TCX, end2)  (end1 = off_TCF | off_WF | packed_floats).
__global__ void __launch_bounds__(256) nif_pack_kernel(const Plan pl, long long G, const float* __restrict__ w_h,
                                                       const float* __restrict__ b_h, float* __restrict__ packed,
                                                       long long end1, long long end2) {
  const long long len1 = end1 - pl.off_M0, len2 = end2 > 0 ? end2 - pl.off_TCX : 0;
  const long long per_group = len1 + len2, total = G * per_group;
  const int K1 = pl.K + 1, NP = pl.NP, n = pl.n, H = pl.H;
  for (long long e0 = blockIdx.x * 256LL + threadIdx.x; e0 < total; e0 += 256LL * gridDim.x) {
    const long long g = e0 / per_group;
    const long long q = e0 - g * per_group;
    long long r = q < len1 ? pl.off_M0 + q : pl.off_TCX + (q - len1);
    const long long e = g * pl.packed_floats + r;
    float v = 0.f;
    if (pl.bf && r >= pl.off_WF) continue;  // bf16 operand tiles: written by nif_pack_bf_kernel
    if (r < pl.off_MHT) {  // MH [H (+1 if wide_last)][K1][NP][NP]
      if (r < (long long)(H + pl.wide_last) * K1 * NP * NP) {
        const int j = r % NP; r /= NP;
        const int i = r % NP; r /= NP;
        const int kk = r % K1; const int h = (int)(r / K1);
        if (h < H) { if (i < n && j < n) v = src_at(pl, w_h, b_h, g, kk, plan_w_off(pl, h + 1) + i * n + j); }
        else if (i < n && j < pl.so) v = src_at(pl, w_h, b_h, g, kk, plan_w_off(pl, H + 1) + i * pl.so + j);  // last matrix [n][so]
      }
    } else if (r < pl.off_M0) {  // MHT: the same, transposed per matrix
      r -= pl.off_MHT;
      if (r < (long long)(H + pl.wide_last) * K1 * NP * NP) {
        const int i = r % NP; r /= NP;
        const int j = r % NP; r /= NP;
        const int kk = r % K1; const int h = (int)(r / K1);
        if (h < H) { if (i < n && j < n) v = src_at(pl, w_h, b_h, g, kk, plan_w_off(pl, h + 1) + i * n + j); }
        else if (i < n && j < pl.so) v = src_at(pl, w_h, b_h, g, kk, plan_w_off(pl, H + 1) + i * pl.so + j);
      }
    } else if (r < pl.off_ML) {  // M0 [K1][si][NP]
      r -= pl.off_M0;
      if (r < (long long)K1 * pl.si * NP) {
        const int j = r % NP; r /= NP;
        const int i = r % pl.si; const int kk = (int)(r / pl.si);
        if (j < n) v = src_at(pl, w_h, b_h, g, kk, i * n + j);
      }
    } else if (r < pl.off_C) {  // ML [K1][NP][so]
      r -= pl.off_ML;
      if (r < (long long)K1 * NP * pl.so) {
        const int c = r % pl.so; r /= pl.so;
        const int i = r % NP; const int kk = (int)(r / NP);
        if (i < n) v = src_at(pl, w_h, b_h, g, kk, plan_w_off(pl, H + 1) + i * pl.so + c);
      }
    } else if (!pl.tc || r < pl.off_TCF) {  // C [Lm][K1][NP]
      r -= pl.off_C;
      if (r < (long long)pl.Lm * K1 * NP) {
        const int j = r % NP; r /= NP;
        const int kk = r % K1; const int m = (int)(r / K1);
        const int width = (m == pl.Lm - 1) ? pl.so : n;
        if (j < width) v = src_at(pl, w_h, b_h, g, kk, plan_b_off(pl, m) + j);
      }
    } else if (r < pl.off_TCS) {
      // tensor-core operand tiles (fp16 pairs stored in one float slot).  Within a 128x64 fp16 tile, element
      // (row nrow, col kcol) lives at byte offset (nrow/8)*1024 + (kcol/8)*128 + (nrow%8)*16 + (kcol%8)*2.
      const bool fwd = r < pl.off_TCB;
      r -= fwd ? pl.off_TCF : pl.off_TCB;
      int t = (int)(r % 4096); r /= 4096;           // float slot inside a 16 KB tile
      const int lo = (int)(r % 2); r /= 2;
      const int c = (int)(r % pl.NCH); const int h = (int)(r / pl.NCH);
      const int kq2 = t % 4; t /= 4;
      const int rq = t % 8; t /= 8;
      const int kc = t % 8; const int rg = t / 8;
      const int nrow = rg * 8 + rq, kcol = kc * 8 + kq2 * 2;
      const int kk = 2 * c + nrow / 64, a = nrow % 64;
      float w[2] = {0.f, 0.f};
      if (kk < K1) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int i = fwd ? kcol + q : a, j = fwd ? a : kcol + q;  // fwd rows are j (K = i); reverse rows are i (K = j)
          if (i < n && j < n) w[q] = src_at(pl, w_h, b_h, g, kk, plan_w_off(pl, h + 1) + i * n + j);
        }
      }
      const float sc = 1.0f / packed[g * pl.packed_floats + pl.off_TCS + h * pl.KP + kk];  // exact power of two
      __half2 out;
      if (!lo) {
        out = __floats2half2_rn(w[0] * sc, w[1] * sc);
      } else {
        const float a0 = w[0] * sc, a1 = w[1] * sc;
        out = __floats2half2_rn(a0 - __half2float(__float2half_rn(a0)), a1 - __half2float(__float2half_rn(a1)));
      }
      v = __uint_as_float(*reinterpret_cast<uint32_t*>(&out));
    } else if (r >= pl.off_TCX && r < pl.off_TCS2) {
      // small forward operand tiles; float slot -> (tile, hi/lo, row, col pair)
      r -= pl.off_TCX;
      const long long nz = (long long)(pl.si + 1 + pl.H) * plan_x0_floats(pl);
      const long long nl = (long long)pl.NLC * plan_xl_floats(pl);
      int T, lo, row, col;
      if (r < nz) {
        T = (int)(r / plan_x0_floats(pl));
        int t = (int)(r % plan_x0_floats(pl));
        lo = t / (32 * pl.KZ); t %= 32 * pl.KZ;
        const int rg = t / (4 * pl.KZ); t %= 4 * pl.KZ;   // 8-row group: KZ/8 chunks x 32 float slots
        const int kc = t / 32; t %= 32;
        row = rg * 8 + t / 4; col = kc * 8 + (t % 4) * 2;
      } else if (r < nz + nl) {
        r -= nz;
        T = pl.si + 1 + pl.H + (int)(r / plan_xl_floats(pl));
        int t = (int)(r % plan_xl_floats(pl));
        lo = t / (pl.LPC * pl.KZ * 32); t %= pl.LPC * pl.KZ * 32;  // hi tile: LPC*KZ rows x 64 fp16
        const int rg = t / 256; t %= 256;
        const int kc = t / 32; t %= 32;
        row = rg * 8 + t / 4; col = kc * 8 + (t % 4) * 2;
      } else {  // reverse-pass tiles [KZ rows x 64]: 8-row group = 8 k-chunks x 32 float slots
        r -= nz + nl;
        T = pl.si + 1 + pl.H + pl.NLC + (int)(r / plan_x0_floats(pl));
        int t = (int)(r % plan_x0_floats(pl));
        lo = t / (pl.KZ * 32); t %= pl.KZ * 32;
        const int rg = t / 256; t %= 256;
        const int kc = t / 32; t %= 32;
        row = rg * 8 + t / 4; col = kc * 8 + (t % 4) * 2;
      }
      const float sc = 1.0f / packed[pl.off_TCS2 + T];
      const float a0 = tcx_value(pl, w_h, b_h, T, row, col) * sc, a1 = tcx_value(pl, w_h, b_h, T, row, col + 1) * sc;
      __half2 out;
      if (!lo) out = __floats2half2_rn(a0, a1);
      else out = __floats2half2_rn(a0 - __half2float(__float2half_rn(a0)), a1 - __half2float(__float2half_rn(a1)));
      v = __uint_as_float(*reinterpret_cast<uint32_t*>(&out));
    } else {
      continue;  // scale tables are written by nif_pack_scales_kernel before this kernel runs
    }
    packed[e] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// Hidden matrices: one block per (group, matrix h, latent coordinate kappa).  The n x n block of the reference layout is
// read once (coalesced) into shared memory and written as MH[h][kappa] (row-major, zero padded to NP x NP), MHT (its
// transpose) and -- for the FP16x3 tensor-core plans -- the forward and reverse operand tiles [hi | lo] of chunk kappa / 2
// with the slab's power-of-two scale (which the block computes itself: max |M| over the matrix).  The per-element kernel
// this replaces spent ~150 instructions of index arithmetic per float: 80 -> ~15 microseconds at C2, where it sits in
// front of every step of a small batch.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nif_pack_mats_kernel(const Plan pl, const float* __restrict__ w_h,
                                                            const float* __restrict__ b_h, float* __restrict__ packed_all) {
  extern __shared__ float Ms[];  // [NP][NP + 1]
  const int K1 = pl.K + 1, NP = pl.NP, n = pl.n, H = pl.H, HT = pl.H + pl.wide_last;
  const int KPAD = pl.tc ? pl.KP : K1;  // the tensor-core image also holds (zero) tiles of the padding coordinates
  const int LD = NP + 1;
  int bid = blockIdx.x;
  const int kk = bid % KPAD; bid /= KPAD;
  const int h = bid % HT;
  const long long g = bid / HT;
  float* packed = packed_all + g * pl.packed_floats;
  const int tid = threadIdx.x;
  const bool real = kk < K1;
  const int rows = n, cols = h < H ? n : pl.so;
  const float* src = nullptr;
  if (real) src = (kk < pl.K ? w_h + (long long)kk * pl.P : b_h + g * pl.P) + plan_w_off(pl, h + 1);
  float mx = 0.f;
  for (int e = tid; e < NP * NP; e += 256) {
    const int i = e / NP, j = e - i * NP;
    float v = 0.f;
    if (real && i < rows && j < cols) v = __ldg(&src[i * cols + j]);
    Ms[i * LD + j] = v;
    mx = fmaxf(mx, fabsf(v));
  }
  __shared__ float red[256];
  __shared__ float sc_s;
  red[tid] = mx;
  __syncthreads();
  if (pl.tc && h < H) {
    for (int o = 128; o >= 1; o >>= 1) {
      if (tid < o) red[tid] = fmaxf(red[tid], red[tid + o]);
      __syncthreads();
    }
    if (tid == 0) {
      float inv = 1.0f;
      const float m = red[0];
      if (m > 0.f && m < 3.0e38f) {
        int ex = (int)((__float_as_uint(m) >> 23) & 0xFF) - 127;  // floor(log2(max)) for normal numbers
        ex = max(-100, min(100, ex));
        inv = __uint_as_float((uint32_t)(127 + ex - 13) << 23);    // 2^(ex-13): scaled max lands in [2^13, 2^14)
      }
      packed[pl.off_TCS + h * pl.KP + kk] = inv;
      sc_s = 1.0f / inv;  // exact power of two
    }
    __syncthreads();
  }
  if (real) {
    float* mh = packed + pl.off_MH + ((long long)h * K1 + kk) * NP * NP;
    float* mht = packed + pl.off_MHT + ((long long)h * K1 + kk) * NP * NP;
    for (int e = tid; e < NP * NP / 4; e += 256) {
      const int i = e / (NP / 4), j4 = (e - i * (NP / 4)) * 4;
      *reinterpret_cast<float4*>(&mh[i * NP + j4]) = make_float4(Ms[i * LD + j4], Ms[i * LD + j4 + 1], Ms[i * LD + j4 + 2], Ms[i * LD + j4 + 3]);
      // transposed: row i of MHT is column i of M
      *reinterpret_cast<float4*>(&mht[i * NP + j4]) = make_float4(Ms[j4 * LD + i], Ms[(j4 + 1) * LD + i], Ms[(j4 + 2) * LD + i], Ms[(j4 + 3) * LD + i]);
    }
  }
  if (pl.tc && h < H) {
    // operand tiles (NP == 64): 16-byte units (row nrow = 64 (kappa % 2) + a, 8 consecutive k) in memory order
    const float sc = sc_s;
    const int c = kk >> 1, half = kk & 1;
    float* tf = packed + pl.off_TCF + ((long long)h * pl.NCH + c) * NIF_TC_CHUNK_FLOATS;
    float* tb = packed + pl.off_TCB + ((long long)h * pl.NCH + c) * NIF_TC_CHUNK_FLOATS;
    for (int u = tid; u < 512; u += 256) {
      const int rg = u >> 6, kc = (u >> 3) & 7, rq = u & 7;
      const int a = rg * 8 + rq;
      const int slot = (((half * 8 + rg) * 8 + kc) * 8 + rq) * 4;  // float slot of the unit inside a [128 x 64] fp16 tile
      uint32_t fh[4], fl[4], bh[4], bl[4];
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) {
        const int k0 = kc * 8 + 2 * q2;
        // forward tile: rows are j = a, K = i;   reverse tile: rows are i = a, K = j
        const float f0 = Ms[k0 * LD + a] * sc, f1 = Ms[(k0 + 1) * LD + a] * sc;
        const float b0 = Ms[a * LD + k0] * sc, b1 = Ms[a * LD + k0 + 1] * sc;
        __half2 o = __floats2half2_rn(f0, f1);
        fh[q2] = *reinterpret_cast<uint32_t*>(&o);
        o = __floats2half2_rn(f0 - __half2float(__float2half_rn(f0)), f1 - __half2float(__float2half_rn(f1)));
        fl[q2] = *reinterpret_cast<uint32_t*>(&o);
        o = __floats2half2_rn(b0, b1);
        bh[q2] = *reinterpret_cast<uint32_t*>(&o);
        o = __floats2half2_rn(b0 - __half2float(__float2half_rn(b0)), b1 - __half2float(__float2half_rn(b1)));
        bl[q2] = *reinterpret_cast<uint32_t*>(&o);
      }
      *reinterpret_cast<uint4*>(&tf[slot]) = make_uint4(fh[0], fh[1], fh[2], fh[3]);
      *reinterpret_cast<uint4*>(&tf[4096 + slot]) = make_uint4(fl[0], fl[1], fl[2], fl[3]);
      *reinterpret_cast<uint4*>(&tb[slot]) = make_uint4(bh[0], bh[1], bh[2], bh[3]);
      *reinterpret_cast<uint4*>(&tb[4096 + slot]) = make_uint4(bl[0], bl[1], bl[2], bl[3]);
    }
  }
}

// one block per small tensor-core tile: inverse power-of-two scale of the tile so that max|.| * 2^e is in [2^13, 2^14)
// (the scales of the main slabs are computed by nif_pack_mats_kernel, which reads those matrices anyway)
__global__ void __launch_bounds__(256) nif_pack_scales_kernel(const Plan pl, const float* __restrict__ w_h,
                                                              const float* __restrict__ b_h, float* __restrict__ packed) {
  const int T = blockIdx.x;
  float m = 0.f;
  {
    const bool fwd_jk = T <= pl.si + pl.H;                      // X0 / XC: [64 (j) x KZ]
    const bool is_xl = !fwd_jk && T < pl.si + 1 + pl.H + pl.NLC;  // XL: [LPC * KZ x 64]; else BCt / B0t: [KZ x 64]
    const int rows = fwd_jk ? 64 : (is_xl ? pl.LPC * pl.KZ : pl.KZ), cols = fwd_jk ? pl.KZ : 64;
    for (int e = threadIdx.x; e < rows * cols; e += 256) m = fmaxf(m, fabsf(tcx_value(pl, w_h, b_h, T, e / cols, e % cols)));
  }
  __shared__ float red[256];
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float inv = 1.0f;
    const float mx = red[0];
    if (mx > 0.f && mx < 3.0e38f) {
      int ex = (int)((__float_as_uint(mx) >> 23) & 0xFF) - 127;  // floor(log2(max)) for normal numbers
      ex = max(-100, min(100, ex));
      inv = __uint_as_float((uint32_t)(127 + ex - 13) << 23);    // 2^(ex-13): scaled max lands in [2^13, 2^14)
    }
    packed[pl.off_TCS2 + T] = inv;
  }
}

int nif_pack_bf_impl(const Plan& pl, long long G, const float* w_h, const float* b_h, float* packed, cudaStream_t st);
int nif_pack_impl(const Plan& pl, long long G, const float* w_h, const float* b_h, float* packed, cudaStream_t st) {
  if (pl.bf) {
    const int rc = nif_pack_bf_impl(pl, G, w_h, b_h, packed, st);
    if (rc) return rc;
  }
  if (pl.tc) {
    { NIF_PROF("nif_pack_scales_kernel", st); nif_pack_scales_kernel<<<(unsigned)plan_n_small(pl), 256, 0, st>>>(pl, w_h, b_h, packed); }
    NIF_CUDA_CHECK(cudaGetLastError());
  }
  const int HT = pl.H + pl.wide_last;
  if (HT > 0) {
    const size_t smem = (size_t)pl.NP * (pl.NP + 1) * sizeof(float);
    NIF_CUDA_CHECK(cudaFuncSetAttribute(nif_pack_mats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long nb = G * HT * (pl.tc ? pl.KP : pl.K + 1);
    { NIF_PROF("nif_pack_mats_kernel", st); nif_pack_mats_kernel<<<(unsigned)nb, 256, smem, st>>>(pl, w_h, b_h, packed); }
    NIF_CUDA_CHECK(cudaGetLastError());
  }
  const long long end1 = pl.tc ? pl.off_TCF : (pl.bf ? pl.off_WF : pl.packed_floats);
  const long long end2 = pl.tc ? pl.off_TCS2 : 0;
  const long long total = G * ((end1 - pl.off_M0) + (end2 > 0 ? end2 - pl.off_TCX : 0));
  long long nblk = (total + 255) / 256;
  if (nblk > 148 * 16) nblk = 148 * 16;
  if (nblk < 1) nblk = 1;
  { NIF_PROF("nif_pack_kernel", st); nif_pack_kernel<<<(unsigned)nblk, 256, 0, st>>>(pl, G, w_h, b_h, packed, end1, end2); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}

// one thread per (kappa, reference column): sum the batch-split partials, write dw_h / db_h
__global__ void __launch_bounds__(256) nif_unpack_grad_kernel(const Plan pl, int S_h, const float* __restrict__ part_h,
                                                              int S_e, const float* __restrict__ part_e, int Q,
                                                              float* __restrict__ dw_h, float* __restrict__ db_h,
                                                              float beta) {
  const int K1 = pl.K + 1, NP = pl.NP, n = pl.n, H = pl.H, P = pl.P;
  // (the hidden matrices, columns [w_hid0, w_last0), are summed by nif_unpack_mats_kernel: this kernel visits the rest)
  const int hid = H * n * n, PT = P - hid;
  const long long total = (long long)K1 * PT;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    const int kk = (int)(e / PT);
    const int tcol = (int)(e - (long long)kk * PT);
    const int w_hid0 = pl.si * n, w_last0 = w_hid0 + hid, b0 = w_last0 + n * pl.so;
    const int col = tcol < w_hid0 ? tcol : tcol + hid;
    float v = 0.f;
    if (col >= w_hid0 && col < w_last0) {
      const int r = col - w_hid0;
      const int h = r / (n * n), ij = r % (n * n), i = ij / n, j = ij % n;
      const long long stride = (long long)(H + pl.wide_last) * K1 * NP * NP;
      const float* src = part_h + (((long long)h * K1 + kk) * NP + i) * NP + j;
      for (int s = 0; s < S_h; ++s) v += src[s * stride];
    } else {
      int q;
      if (col < w_hid0) {  // first matrix
        const int i = col / n, j = col % n;
        q = (H + 1) * NP + pl.so + i * NP + j;
      } else if (col < b0) {  // last matrix
        const int r = col - w_last0;
        if (pl.wide_last) {  // produced by the hidden-matrix GEMM as matrix index H: [i][c] inside an NP x NP tile
          const int i = r / pl.so, c = r % pl.so;
          const long long stride = (long long)(H + 1) * K1 * NP * NP;
          const float* src = part_h + (((long long)H * K1 + kk) * NP + i) * NP + c;
          for (int s = 0; s < S_h; ++s) v += src[s * stride];
          float* dst = (kk < pl.K) ? &dw_h[(long long)kk * P + col] : &db_h[col];
          *dst = (beta != 0.f) ? (*dst * beta + v) : v;
          continue;
        }
        q = (H + 1) * NP + pl.so + pl.si * NP + r;  // r = i * so + c
      } else {
        const int r = col - b0;
        const int m = r / n;
        if (m <= H) q = m * NP + (r - m * n);
        else q = (H + 1) * NP + (r - (H + 1) * n);
      }
      const long long stride = (long long)K1 * Q;
      const float* src = part_e + (long long)kk * Q + q;
      for (int s = 0; s < S_e; ++s) v += src[s * stride];
    }
    float* dst = (kk < pl.K) ? &dw_h[(long long)kk * P + col] : &db_h[col];
    *dst = (beta != 0.f) ? (*dst * beta + v) : v;
  }
}

// hidden matrices: one block per (matrix h, kappa): sum the batch-split partials of the NP x NP tile (coalesced), write the
// n x n block of the reference layout (contiguous in dw_h / db_h)
__global__ void __launch_bounds__(256) nif_unpack_mats_kernel(const Plan pl, int S_h, const float* __restrict__ part_h,
                                                              float* __restrict__ dw_h, float* __restrict__ db_h, float beta) {
  const int K1 = pl.K + 1, NP = pl.NP, n = pl.n, P = pl.P;
  const int RB = (n + 15) / 16;  // blocks of 16 rows per matrix (enough blocks in flight to hide the partial-sum latency)
  int bid = blockIdx.x;
  const int rb = bid % RB; bid /= RB;
  const int kk = bid % K1, h = bid / K1;
  const int i0 = rb * 16, i1 = min(n, i0 + 16);
  const long long stride = (long long)(pl.H + pl.wide_last) * K1 * NP * NP;
  const float* src = part_h + ((long long)h * K1 + kk) * NP * NP;
  float* dst = ((kk < pl.K) ? dw_h + (long long)kk * P : db_h) + pl.si * n + h * n * n;
  if ((n & 3) == 0) {  // rows of both layouts are 16-byte aligned (P's sections are multiples of n)
    const int n4 = n / 4;
    for (int e = i0 * n4 + threadIdx.x; e < i1 * n4; e += 256) {
      const int i = e / n4, j4 = (e - i * n4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < S_h; ++s) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(src + s * stride + i * NP + j4));
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      float* d = dst + i * n + j4;
      if (((reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
        float4* d4 = reinterpret_cast<float4*>(d);
        if (beta != 0.f) { const float4 o = *d4; v.x += beta * o.x; v.y += beta * o.y; v.z += beta * o.z; v.w += beta * o.w; }
        *d4 = v;
      } else {
        const float vv[4] = {v.x, v.y, v.z, v.w};
        for (int q = 0; q < 4; ++q) d[q] = (beta != 0.f) ? (d[q] * beta + vv[q]) : vv[q];
      }
    }
  } else {
    for (int e = i0 * n + threadIdx.x; e < i1 * n; e += 256) {
      const int i = e / n, j = e - i * n;
      float v = 0.f;
      for (int s = 0; s < S_h; ++s) v += src[s * stride + i * NP + j];
      float* d = dst + e;
      *d = (beta != 0.f) ? (*d * beta + v) : v;
    }
  }
}

int nif_unpack_grad_impl(const Plan& pl, int S_h, const float* part_h, int S_e, const float* part_e, int Q,
                         float* dw_h, float* db_h, float beta, cudaStream_t st) {
  if (pl.H > 0) {
    { NIF_PROF("nif_unpack_mats_kernel", st); nif_unpack_mats_kernel<<<(unsigned)(pl.H * (pl.K + 1) * ((pl.n + 15) / 16)), 256, 0, st>>>(pl, S_h, part_h, dw_h, db_h, beta); }
    NIF_CUDA_CHECK(cudaGetLastError());
  }
  const long long total = (long long)(pl.K + 1) * (pl.P - (long long)pl.H * pl.n * pl.n);
  long long nblk = (total + 255) / 256;
  if (nblk > 148 * 16) nblk = 148 * 16;
  { NIF_PROF("nif_unpack_grad_kernel", st); nif_unpack_grad_kernel<<<(unsigned)nblk, 256, 0, st>>>(pl, S_h, part_h, S_e, part_e, Q, dw_h, db_h, beta); }
  NIF_CUDA_CHECK(cudaGetLastError());
  return NIF_OK;
}
