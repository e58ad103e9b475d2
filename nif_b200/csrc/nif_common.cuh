// nif_b200 — shared device/host definitions for the fused NIF kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nif_b200.h"

#define NIF_MAX_SO 8        // max ShapeNet output width handled in registers
#define NIF_MAX_SI 8        // max ShapeNet input width
#define NIF_MAX_DIR 4       // max tangent directions per launch

// ---------------------------------------------------------------------------------------------
// Plan: everything the kernels need to know about one descriptor, computed on the host.
//
// Packed weight image (one per group), all fp32, every section 16-byte aligned:
//   MH  [H][K+1][NP][NP]   hidden matrices, M[kappa][i][j]  (i = input index, j = output index)
//   MHT [H][K+1][NP][NP]   the same, transposed per matrix: MT[kappa][j][i]   (for the reverse pass)
//   M0  [K+1][si][NP]      first matrix
//   ML  [K+1][NP][so]      last matrix
//   C   [Lm][K+1][NP]      bias rows of every layer (last layer uses the first `so` entries)
// kappa < K indexes rows of w_h; kappa == K is the row taken from b_h (its latent coordinate is 1).
// Padding (i,j >= n) is zero.
// ---------------------------------------------------------------------------------------------
struct Plan {
  int variant, act, si, so, n, l, K;
  int H;    // hidden matrices (l or 2l)
  int Lm;   // total matrices H + 2
  int NP;   // padded width
  int P;    // po_dim
  float omega0;
  long long off_MH, off_MHT, off_M0, off_ML, off_C, packed_floats;
  // tensor-core operand images (dtype_compute == 2, NP == 64).  fp32-grade accuracy on the fp16 tensor-core
  // path: every value is scaled by a power of two and split  x * 2^e = hi + lo  (two fp16), products use the
  // three terms hi*hi + lo*hi + hi*lo accumulated in fp32.  Per hidden matrix h and per chunk c of 2 latent
  // coordinates the image holds a [hi | lo] pair of 128x64 fp16 tiles (16 KB each) in the UMMA K-major
  // core-matrix layout (nif_tc.cuh):
  //   TCF: rows (kappa_l, j), K = i  (forward)        TCB: rows (kappa_l, i), K = j  (reverse data pass)
  //   TCS: [H][KP] fp32 inverse scales 2^-e of slab (h, kappa)   (KP = K+1 rounded up to even, NCH = KP/2)
  int tc, KP, NCH;
  long long off_TCF, off_TCB, off_TCS;
  // small forward operand tiles (same split / layout, one inverse scale per tile in TCS2):
  //   X0[i'] (i' = 0..si)  [64 (j) x KZ (kappa)]   M0[kappa][i'][j], or C_0[kappa][j] for i' = si
  //   XC[m]  (m = 1..H)    [64 (j) x KZ (kappa)]   C_m[kappa][j]
  //   XL[q]  (q < NLC)     [LPC*KZ rows x 64 (i)]  row (c_l * KZ + kappa) = ML[kappa][i][LPC*q + c_l]
  // KZ = K+1 rounded up to 16; LPC = outputs per XL tile (2 if 2*KZ <= 128 else 1); NLC = ceil(so / LPC).
  int KZ, LPC, NLC;
  long long off_TCX, off_TCS2;
  // wide_last (trunk plans): the last matrix [n x so] is wide (so up to 256); its gradient runs through the
  // hidden-matrix batch-reduction GEMM as matrix index H (rows h_{H+1}, columns = the seed du padded to NP)
  // instead of the column-per-thread edge kernel.
  int wide_last;
  // rows per accumulation chain of the tensor-core batch reductions (nif_desc_t.acc_rows; 0 = default 4096)
  int acc_rows;
  // bf16 single-product tensor-core path (dtype_compute == 1, nif_bf.cuh): bf16 operand tiles appended to the image
  int bf;
  long long off_WF, off_WB, off_WX;
};
__host__ __device__ inline long long plan_x0_floats(const Plan& p) { return 64LL * p.KZ; }            // one X0 / XC chunk [hi|lo]
__host__ __device__ inline long long plan_xl_floats(const Plan& p) { return (long long)p.LPC * p.KZ * 64; }       // one XL chunk [hi|lo]
// small tiles: X0[si+1], XC[H], XL[NLC] (forward), then BCt[m] m = 0..H and B0t[i] i < si (reverse data pass: [KZ (kappa) x 64 (j)]
// tiles of C_m[kappa][j] and M0[kappa][i][j], the thin dz terms)
__host__ __device__ inline int plan_n_small(const Plan& p) { return p.si + 1 + p.H + p.NLC + p.H + 1 + p.si; }
#define NIF_TC_CHUNK_FLOATS 8192   // [hi | lo] x 128 x 64 fp16 = 32 KB, counted in floats

__host__ __device__ inline int plan_w_off(const Plan& p, int m) {  // reference column offset of matrix m
  if (m == 0) return 0;
  if (m <= p.H) return p.si * p.n + (m - 1) * p.n * p.n;
  return p.si * p.n + p.H * p.n * p.n;
}
__host__ __device__ inline int plan_b_off(const Plan& p, int m) {  // reference column offset of bias m
  int nW = p.si * p.n + p.H * p.n * p.n + p.n * p.so;
  return nW + m * p.n;
}

// layer semantics shared by forward and backward ------------------------------------------------
// out_m = alpha * act(pre_m) + (residual), pre_m = omega * lin + bias
__host__ __device__ inline float plan_omega(const Plan& p, int m) {
  return (p.variant == NIF_VARIANT_NIF || m == p.Lm - 1) ? 1.0f : p.omega0;
}
__host__ __device__ inline float plan_alpha(const Plan& p, int m) {
  // second layer of a res-block: 0.5 * (u + sin(...))
  return (p.variant == NIF_VARIANT_SIREN_RES && m >= 2 && m <= p.H && (m & 1) == 0) ? 0.5f : 1.0f;
}
// residual bookkeeping: 0 none, 1 "out += in" (NIF hidden), 2 "remember input" (first of res-block),
// 3 "out += 0.5 * remembered" (second of res-block)
__host__ __device__ inline int plan_res(const Plan& p, int m) {
  if (m < 1 || m > p.H) return 0;
  if (p.variant == NIF_VARIANT_NIF) return 1;
  if (p.variant == NIF_VARIANT_SIREN_RES) return (m & 1) ? 2 : 3;
  return 0;
}

// activations -----------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// sin and cos together, fp32-accurate (~1 ulp), compact: three-constant Cody-Waite reduction by pi/2 (good to
// |v| < 1e5, far beyond omega_0-scaled SIREN pre-activations) and the Cephes minimax polynomials on [-pi/4, pi/4].
// ~25 instructions inline (the library's inlined Payne-Hanek slow path bloats a 64-wide unrolled epilogue past the
// instruction cache).
// huge / non-finite arguments (never seen for sane SIREN pre-activations): fold into [-pi, pi] in double precision
// (exact quadrant up to ~1e15; NaN / inf stay NaN).  Out of line and by value, so the hot path carries neither the
// fp64 instructions (the compiler would predicate them into every element) nor address-taken outputs.
static __device__ __noinline__ float nif_fold_2pi(float v) {
  const double q = rint((double)v * 0.15915494309189535);
  return (float)fma(-q, 6.283185307179586, (double)v);
}
// branch-free core, valid for |v| < 1e5 (NaN propagates).  Callers that evaluate many elements test the whole group
// once (nif_sincos_fold8) and then run the cores back to back: without a branch per element the compiler interleaves
// the independent polynomial chains instead of executing them one after the other.
__device__ __forceinline__ void nif_sincosf_core(float v, float& s, float& c);
__device__ __forceinline__ void nif_sincosf(float v, float& s, float& c) {
  if (!(fabsf(v) < 1.0e5f)) v = nif_fold_2pi(v);
  nif_sincosf_core(v, s, c);
}
// folds the huge / non-finite members of a group of 8 arguments (one test for the group)
__device__ __forceinline__ void nif_sincos_fold8(float* v) {
  float mx = 0.f;
  bool bad = false;
#pragma unroll
  for (int e = 0; e < 8; ++e) { mx = fmaxf(mx, fabsf(v[e])); bad |= (v[e] != v[e]); }
  if (bad || !(mx < 1.0e5f)) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (!(fabsf(v[e]) < 1.0e5f)) v[e] = nif_fold_2pi(v[e]);
  }
}
__device__ __forceinline__ void nif_sincosf_core(float v, float& s, float& c) {
  const float kf = rintf(v * 0.636619747f);
  const int k = __float2int_rn(kf);
  float r = fmaf(kf, -1.57079601e+00f, v);
  r = fmaf(kf, -3.13916473e-07f, r);
  r = fmaf(kf, -5.39030253e-15f, r);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  cp = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const float a = (k & 1) ? cp : sp, b = (k & 1) ? sp : cp;
  s = (k & 2) ? -a : a;
  c = ((k + 1) & 2) ? -b : b;
}

// value and first derivative
__device__ __forceinline__ void act_fd(int act, float v, float& f, float& d) {
  switch (act) {
    case NIF_ACT_SINE: nif_sincosf(v, f, d); break;
    case NIF_ACT_SWISH: { float sg = sigmoidf_(v); f = v * sg; d = sg * (1.0f + v * (1.0f - sg)); } break;
    case NIF_ACT_TANH: { float t = tanhf(v); f = t; d = 1.0f - t * t; } break;
    case NIF_ACT_RELU: f = v > 0.f ? v : 0.f; d = v > 0.f ? 1.f : 0.f; break;
    case NIF_ACT_SIGMOID: { float sg = sigmoidf_(v); f = sg; d = sg * (1.0f - sg); } break;
    default: f = v; d = 1.0f; break;
  }
}
// second derivative, given v and the already computed f = act(v), d = act'(v)  (reverse-over-forward pass)
__device__ __forceinline__ float act_dd(int act, float v, float f, float d) {
  switch (act) {
    case NIF_ACT_SINE: return -f;
    case NIF_ACT_SWISH: { const float sg = sigmoidf_(v); return sg * (1.0f - sg) * (2.0f + v * (1.0f - 2.0f * sg)); }
    case NIF_ACT_TANH: return -2.0f * f * d;
    case NIF_ACT_SIGMOID: return d * (1.0f - 2.0f * f);
    default: return 0.0f;  // relu, linear
  }
}
// four at a time, out of line: ONE copy of the activation switch per kernel instead of one per unrolled element
struct ActFd4 { float4 f, d; };
static __device__ __noinline__ ActFd4 act_fd4v(int act, float4 v) {
  ActFd4 r;
  act_fd(act, v.x, r.f.x, r.d.x);
  act_fd(act, v.y, r.f.y, r.d.y);
  act_fd(act, v.z, r.f.z, r.d.z);
  act_fd(act, v.w, r.f.w, r.d.w);
  return r;
}
__device__ __forceinline__ void act_fd4(int act, const float (&v)[4], float (&f)[4], float (&d)[4]) {
  const ActFd4 r = act_fd4v(act, make_float4(v[0], v[1], v[2], v[3]));
  f[0] = r.f.x; f[1] = r.f.y; f[2] = r.f.z; f[3] = r.f.w;
  d[0] = r.d.x; d[1] = r.d.y; d[2] = r.d.z; d[3] = r.d.w;
}
__device__ __forceinline__ float act_f(int act, float v) {
  switch (act) {
    case NIF_ACT_SINE: { float s, c; nif_sincosf(v, s, c); return s; }
    case NIF_ACT_SWISH: return v * sigmoidf_(v);
    case NIF_ACT_TANH: return tanhf(v);
    case NIF_ACT_RELU: return v > 0.f ? v : 0.f;
    case NIF_ACT_SIGMOID: return sigmoidf_(v);
    default: return v;
  }
}

// PTX helpers: mbarrier + 1-D bulk async copy (TMA unit, no tensor map needed) --------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// 256-bit global accesses (sm_100: LDG.256 / STG.256): a thread that owns a row moves whole 32-byte sectors, half the
// instructions of float4 accesses.  p must be 32-byte aligned.
__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg8(float* p, float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4),
               "f"(a5), "f"(a6), "f"(a7)
               : "memory");
}

// Tiled activation layout of the tensor-core path (NP = 64; stash slots h_m, d_m and the reverse pass's da_m): the
// kernels that own a row per thread would otherwise touch one 32-byte sector per lane and instruction.  Rows are
// grouped by 32; inside a group the 16 column quads follow each other, and inside a quad the 32 rows' float4:
//   offset(b, j) = (b >> 5) * 2048 + (j >> 2) * 128 + (b & 31) * 4 + (j & 3)
// so a warp that reads one quad of 32 consecutive rows moves 512 contiguous bytes, and a 64-row sub-tile is one
// contiguous 16 KB block (a single bulk copy) whose rows are conflict-free in shared memory.  Slots hold
// nif_tiled_rows(B) rows (B rounded up to 64).
__host__ __device__ inline long long nif_tiled_rows(long long B) { return (B + 63) / 64 * 64; }
__host__ __device__ inline long long nif_tiled_row(long long b) { return (b >> 5) * 2048 + (b & 31) * 4; }
__host__ __device__ inline int nif_tiled_col(int j) { return (j >> 2) * 128 + (j & 3); }
// the same layout for a slot of NP columns (the bf16 tensor-core path, nif_bf.cuh): row groups are 32 NP floats apart
__host__ __device__ inline long long nif_tiled_row_np(long long b, int NP) { return (b >> 5) * (32LL * NP) + (b & 31) * 4; }
// static shape test shared by forward and reverse: does this plan run on the tensor-core kernels (tiled stash)?
bool nif_plan_uses_tc(const Plan& pl);
// ... and its Sobolev step (forward tangents with a stash, reverse-over-forward) for ShapeNet-input directions
bool nif_plan_tc_sobolev(const Plan& pl);

// reverse-pass workspace layout (offsets in floats), shared by nif_bwd.cu / nif_api.cu / nif_trunk.cu
struct GradWs {
  long long da, du, part_h, part_e, loss_part, maxes, total;
  int S_h, S_e, S_e_ws, Q;
  long long rows_h, rows_e;
};

// host-side error plumbing ------------------------------------------------------------------------
void nif_set_error(const char* fmt, ...);
int nif_make_plan(const nif_desc_t* d, Plan* out);

// per-kernel timing for the benchmark's kernel table (nif_profile_begin / nif_profile_end): while a profile is open,
// every launch of the library is bracketed by two CUDA events on its own stream
extern int g_nif_prof_on;
void nif_prof_push(const char* name, cudaStream_t st, bool begin);
struct NifProfScope {
  const char* name;
  cudaStream_t st;
  bool on;
  NifProfScope(const char* n, cudaStream_t s) : name(n), st(s), on(g_nif_prof_on != 0) { if (on) nif_prof_push(name, st, true); }
  ~NifProfScope() { if (on) nif_prof_push(name, st, false); }
};
#define NIF_PROF(name, st) NifProfScope nif_prof_scope_(name, st)

#define NIF_CUDA_CHECK(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      nif_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NIF_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)
