// Building blocks of the bf16 single-product tensor-core path (dtype_compute = 1; what the reference's
// `mixed_bfloat16` policy allows, nif/model.py:101-105, 146, 530-533, 954): bf16 operands, fp32 accumulation in TMEM,
// ONE tcgen05.mma per algorithmic MAC (the FP16x3 path of nif_tc.cuh issues three).  Padded widths NP = 64 and 128.
//
// Operand tiles are bf16 in the UMMA "K-major, no swizzle" core-matrix layout with a K extent KD (a multiple of 16):
//   byte offset(row r, col k) = (r/8) * SBO + (k/8) * 128 + (r%8) * 16 + (k%8) * 2,   SBO = (KD/8) * 128
// bf16 has the exponent range of fp32, so there are no operand scales anywhere on this path.
//
// Work split of the row-owning kernels: TWO threads per row.  Thread (r, half) owns TMEM lane r and the column half
// [half * NP/2, (half+1) * NP/2) of every accumulator it drains; warps w and w + 4 share the TMEM lanes 32 (w & 3) ..
// (a warp may only touch the lane quarter given by its index modulo 4).  That is what lifts the register ceiling of
// the one-thread-per-row kernels of nif_tc_fwd.cu (acc[64] + h[64] per thread) to NP = 128.
//
// Activation stash / da slots use the tiled layout of nif_common.cuh generalised to NP columns:
//   offset(b, j) = (b >> 5) * 32 NP + (j >> 2) * 128 + (b & 31) * 4 + (j & 3)
// so the 32 lanes of a warp (32 consecutive rows, same column quad) move 512 contiguous bytes per access.
#pragma once
#include <cuda_bf16.h>
#include "nif_tc.cuh"

#define BF_STAGE_BYTES 32768u  // one weight-stream stage: a main chunk (128 x NP bf16) or a small tile (NP x KZ bf16)

__host__ __device__ inline long long bf_tiled_row(long long b, int NP) { return (b >> 5) * (32LL * NP) + (b & 31) * 4; }
__host__ __device__ inline long long bf_slot_floats(long long B, int NP) { return nif_tiled_rows(B) * NP; }

// shared-memory matrix descriptor, K-major, no swizzle, LBO = 128, caller's SBO
__device__ __forceinline__ uint64_t bf_make_desc(uint32_t saddr, uint32_t sbo) { return tc_make_desc(saddr, sbo); }
// instruction descriptor, kind::f16 with bf16 inputs, D = fp32, M = 128; a_mn / b_mn: operand is MN-major
__host__ __device__ constexpr uint32_t bf_idesc(int N, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((128u >> 4) << 24);
}
__device__ __forceinline__ uint32_t bf_pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// 8 columns (one fp32 per lane per column) -> 8 registers
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// The W image: bf16 operand tiles appended to the packed fp32 image (offsets in floats; one float slot = 2 bf16).
//   WF [H][NCHW]  forward main chunks  [128 rows (kappa_l, j) x NP (i)]   kappa = CK * chunk + kappa_l, CK = 128 / NP
//   WB [H][NCHW]  reverse main chunks  [128 rows (kappa_l, i) x NP (j)]
//   small tiles, NP * KZ bf16 each (KZ = K + 1 rounded up to 16):
//     "JK" geometry [NP rows x KZ]:  X0[i'] i' <= si  (M0[kappa][i'][j]; i' = si: C_0[kappa][j]),  XC[m] m = 1..H
//                                    (C_m[kappa][j]),  BLT[c] c < so  (rows i: ML[kappa][i][c])
//     "KJ" geometry [KZ rows x NP]:  XL[c] c < so  (ML[kappa][i][c], K = i),  BC[m] m = 0..H  (C_m[kappa][j], K = j),
//                                    B0[i] i < si  (M0[kappa][i][j], K = j)
__host__ __device__ inline int bf_ck(const Plan& p) { return 128 / p.NP; }
__host__ __device__ inline int bf_nchw(const Plan& p) { return (p.K + 1 + bf_ck(p) - 1) / bf_ck(p); }
__host__ __device__ inline long long bf_chunk_floats(const Plan& p) { return 64LL * p.NP; }       // 128 x NP bf16
__host__ __device__ inline long long bf_small_floats(const Plan& p) { return (long long)p.NP * p.KZ / 2; }
__host__ __device__ inline int bf_n_jk(const Plan& p) { return p.si + 1 + p.H + p.so; }
__host__ __device__ inline int bf_n_kj(const Plan& p) { return p.so + p.H + 1 + p.si; }
// indices into the small-tile list
__host__ __device__ inline int bf_t_x0(const Plan& p, int i) { return i; }
__host__ __device__ inline int bf_t_xc(const Plan& p, int m) { return p.si + m; }                  // m = 1..H
__host__ __device__ inline int bf_t_blt(const Plan& p, int c) { return p.si + 1 + p.H + c; }
__host__ __device__ inline int bf_t_xl(const Plan& p, int c) { return bf_n_jk(p) + c; }
__host__ __device__ inline int bf_t_bc(const Plan& p, int m) { return bf_n_jk(p) + p.so + m; }     // m = 0..H
__host__ __device__ inline int bf_t_b0(const Plan& p, int i) { return bf_n_jk(p) + p.so + p.H + 1 + i; }

// static shape test: does this plan run on the bf16 tensor-core kernels?
bool nif_plan_uses_bf(const Plan& pl);
