// tcgen05 / TMEM building blocks for the tensor-core realisation of the shared-operand GEMMs
// (descriptor encodings validated on hardware by tc_probe.cu).
//
// Operand tiles are fp16 [128 rows x 64] in the UMMA "K-major, no swizzle" core-matrix layout:
//   byte offset(row r, col k) = (r/8) * SBO + (k/8) * LBO + (r%8) * 16 + (k%8) * 2,   LBO = 128, SBO = 1024
// i.e. 8-row x 16-byte core matrices, contiguous along K first, then along rows (16 KB per tile).  A thread
// that owns one row writes 16-byte chunks; the 8 rows of a quarter warp land in 128 contiguous bytes.
//
// fp32-grade accuracy on the fp16 pipe ("FP16x3"): every operand row is scaled by a power of two so that its
// largest magnitude lies in [2^13, 2^14), then split  x * 2^e = hi + lo  with hi = fp16(x 2^e), lo = fp16(x 2^e - hi)
// (22 significant bits for the large entries, absolute error 2^-25 relative to the row maximum for the rest).
// Products use  hi*hi + lo*hi + hi*lo  with fp32 accumulation in TMEM (cross terms first, see tc_mma_split).
// The scales are powers of two, so un-scaling in the epilogue is exact.
#pragma once
#include <cuda_fp16.h>
#include "nif_common.cuh"

#define TC_LBO 128u
#define TC_SBO 1024u
#define TC_TILE_BYTES 16384u  // 128 rows x 64 fp16

__device__ __forceinline__ uint64_t tc_make_desc(uint32_t saddr, uint32_t sbo = TC_SBO) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((TC_LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;                // layout type 0 (no swizzle), base offset 0
}
// instruction descriptor, kind::f16 with fp16 inputs, D = fp32, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t tc_idesc_f16(int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (the `.ts` form): lane = row of A, 32-bit column c holds the 16-bit elements k = 2c (low half)
// and 2c + 1; one instruction covers K = 16 = 8 columns
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns <- 32 registers per thread; tc_wait_st() before the data is handed over
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"
      "%31,%32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D = Alo*Bhi + Ahi*Blo + Ahi*Bhi over a K extent of 16 * ksteps, ONE accumulator: the small cross terms of all k-steps
// are accumulated first and the dominant hi*hi terms last, so only `ksteps` instructions truncate (the tensor core rounds
// its accumulator toward zero on every instruction) at the magnitude of the result -- the same count a dedicated hi*hi
// accumulator would see -- while the cross terms' truncations are 2^-11 smaller.
__device__ __forceinline__ void tc_mma_split(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                             uint32_t idesc, int ksteps) {
  for (int ks = 0; ks < ksteps; ++ks) {  // 16 elements of K = 2 core matrices = 256 B = 16 descriptor units
    const uint64_t adv = (uint64_t)(ks * 16);
    tc_mma_f16(d, a_lo + adv, b_hi + adv, idesc, ks > 0 ? 1u : 0u);
    tc_mma_f16(d, a_hi + adv, b_lo + adv, idesc, 1u);
  }
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t adv = (uint64_t)(ks * 16);
    tc_mma_f16(d, a_hi + adv, b_hi + adv, idesc, 1u);
  }
}
__device__ __forceinline__ void tc_mma_split_k64(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                                 uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t adv = (uint64_t)(ks * 16);
    tc_mma_f16(d, a_lo + adv, b_hi + adv, idesc, ks > 0 ? 1u : 0u);
    tc_mma_f16(d, a_hi + adv, b_lo + adv, idesc, 1u);
  }
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d, a_hi + (uint64_t)(ks * 16), b_hi + (uint64_t)(ks * 16), idesc, 1u);
}
// One lane of a converged warp.  The MMA-issuer warps run their whole loop warp-uniformly and put only the tcgen05.mma /
// commit instructions under this predicate: descriptors and addresses then live in uniform registers.  (Inside an
// `if (lane == 0)` region the compiler cannot use the uniform datapath and wraps EVERY tcgen05.mma in an
// elect / 5 x R2UR.BROADCAST / branch loop, ~90 cycles of issue per 64-cycle instruction.)
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.b32 %0, 1, 0, q;\n\t}\n" : "=r"(p));
  return p != 0;
}
// arrive on an mbarrier when every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t tmem, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive columns (one fp32 per lane per column) -> 32 registers per thread
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// warp-group register re-allocation (all 4 warps of an aligned warp group must execute it)
template <int N> __device__ __forceinline__ void tc_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void tc_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// barrier among a subset of warps (id 1..15), all `nthreads` must call it
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// power-of-two scale for a row whose largest magnitude is `amax`: returns sc = 2^e with amax*sc in [2^13, 2^14)
// and inv = 2^-e (both exact); (1, 1) for an all-zero or non-finite row
__device__ __forceinline__ void tc_row_scale(float amax, float& sc, float& inv) {
  sc = 1.f;
  inv = 1.f;
  if (amax > 0.f && amax < 3.0e38f) {
    int ex = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127;
    ex = max(-100, min(100, ex));
    sc = __uint_as_float((uint32_t)(127 + 13 - ex) << 23);
    inv = __uint_as_float((uint32_t)(127 + ex - 13) << 23);
  }
}

// write one row of an operand tile: 64 values * sc -> hi tile and lo tile (8 x st.shared.v4 each)
__device__ __forceinline__ void tc_store_row_split(unsigned char* tile_hi, unsigned char* tile_lo, int r,
                                                   const float (&h)[64], float sc) {
  const uint32_t row_off = (uint32_t)(r >> 3) * TC_SBO + (uint32_t)(r & 7) * 16u;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a0 = h[8 * c + 2 * e] * sc, a1 = h[8 * c + 2 * e + 1] * sc;
      const __half2 hh = __floats2half2_rn(a0, a1);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(tile_hi + row_off + c * TC_LBO) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(tile_lo + row_off + c * TC_LBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// same for a row of `w8 * 8` values held in shared memory style access v(k) (used for the zt operand tile)
template <class F>
__device__ __forceinline__ void tc_store_row_split_fn(unsigned char* tile_hi, unsigned char* tile_lo, int r, uint32_t sbo,
                                                      int w8, float sc, F value) {
  const uint32_t row_off = (uint32_t)(r >> 3) * sbo + (uint32_t)(r & 7) * 16u;
  for (int c = 0; c < w8; ++c) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a0 = value(8 * c + 2 * e) * sc, a1 = value(8 * c + 2 * e + 1) * sc;
      const __half2 hh = __floats2half2_rn(a0, a1);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(tile_hi + row_off + c * TC_LBO) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(tile_lo + row_off + c * TC_LBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Optional timeline trace (make -C nif_b200/csrc trace): CTA 0 records clock64() at the hand-offs of its first tile
// pairs (per translation unit); read back with the reader each kernel file instantiates.  Compiled out of the product
// build.  Roles: 0 / 1 = epilogue of tile 0 / 1, 2 = MMA issuer of tile 0, 3 = weight-stream producer.
#ifdef NIF_TRACE
static __device__ long long g_trace[4][2048];
static __device__ int g_trace_n[4];
#define TRACE(role, tag)                                                                  \
  do {                                                                                    \
    if (blockIdx.x == 0 && trace_n < 1023) {                                              \
      g_trace[role][2 * trace_n] = (long long)(tag);                                      \
      g_trace[role][2 * trace_n + 1] = clock64();                                         \
      ++trace_n;                                                                          \
      g_trace_n[role] = 2 * trace_n;                                                      \
    }                                                                                     \
  } while (0)
#define NIF_TRACE_READER(name)                                                            \
  extern "C" int name(long long* host, int* counts) {                                     \
    cudaMemcpyFromSymbol(host, g_trace, sizeof(g_trace));                                 \
    cudaMemcpyFromSymbol(counts, g_trace_n, sizeof(g_trace_n));                           \
    return 0;                                                                             \
  }
#else
#define TRACE(role, tag) do {} while (0)
#define NIF_TRACE_READER(name)
#endif
