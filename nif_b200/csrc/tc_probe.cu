// Standalone tcgen05 probe (development tool, not part of libnif_b200.so):
//   1. correctness of the UMMA descriptors used by the fused kernels (K-major, no-swizzle core-matrix layout,
//      kind::tf32 and kind::f16) against a CPU product,
//   2. cycles per tcgen05.mma for M=128, N in {64,128,256},
//   3. tcgen05.ld throughput (TMEM -> registers), which bounds the per-row epilogue.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout_type 0 = no swizzle, base_offset 0
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                 "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                 "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// mode 0: tf32 (4-byte elements), mode 1: fp16 (2-byte elements), mode 2: fp16 with MN-major operands
// A [128 x KD], B [N x KD] row-major in global (as floats); D [128 x N] out.  KD = 64.
template <int MODE, int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, int reps, long long* cyc,
                                                   int ld_reps) {
  constexpr int KD = 64;
  constexpr int ES = MODE == 0 ? 4 : 2;            // element size
  constexpr int EPC = 16 / ES;                     // elements per 16-byte chunk
  constexpr int NCH = KD / EPC;                    // k-chunks per row
  constexpr int LBO = 128, SBO = NCH * 128;        // core matrices adjacent along K; row groups after that
  constexpr int KSTEP_BYTES = 2 * LBO;             // one MMA consumes 32 bytes of K = 2 chunks
  constexpr int NK = NCH / 2;                      // MMAs per tile
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned char* sA = sm;
  unsigned char* sB = sm + 128 * KD * ES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // operands -> shared, core-matrix layout: addr(r, k) = (r/8)*SBO + (k/EPC)*LBO + (r%8)*16 + (k%EPC)*ES
  auto put = [&](unsigned char* base, int r, int k, float v) {
    // K-major: 8 rows x 16 B core matrices; MN-major (mode 2): 8 k x 16 B (8 consecutive rows) core matrices
    const int off = MODE == 2 ? (r / 8) * SBO + (k / 8) * LBO + (k % 8) * 16 + (r % 8) * 2
                              : (r / 8) * SBO + (k / EPC) * LBO + (r % 8) * 16 + (k % EPC) * ES;
    if (MODE == 0) *reinterpret_cast<float*>(base + off) = v;
    else *reinterpret_cast<__half*>(base + off) = __float2half_rn(v);
  };
  for (int i = tid; i < 128 * KD; i += 128) put(sA, i / KD, i % KD, A[i]);
  for (int i = tid; i < N * KD; i += 128) put(sB, i / KD, i % KD, B[i]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_base_s;

  // instruction descriptor: D fp32, A/B format, K-major both, N>>3 at bit 17, M>>4 at bit 24
  const uint32_t fmt = MODE == 0 ? 2u : 0u;  // TF32 = 2, F16 = 0
  const uint32_t mn = MODE == 2 ? 1u : 0u;   // a_major (bit 15) / b_major (bit 16): 1 = MN-major
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (mn << 15) | (mn << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);

  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < NK; ++ks) {
        const uint64_t da = make_desc(s32(sA) + ks * KSTEP_BYTES, LBO, SBO);
        const uint64_t db = make_desc(s32(sB) + ks * KSTEP_BYTES, LBO, SBO);
        if (MODE == 0) mma_tf32(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        else mma_f16(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) { t1 = clock64(); cyc[0] = t1 - t0; }
  fence_after();

  // epilogue: warp w owns lanes [32w, 32w+32)
  uint32_t v[32];
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < N; c0 += 32) {
    tmem_ld32(tmem + lane_base + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  // TMEM read throughput: every warp re-reads its N columns ld_reps times
  __syncthreads();
  long long u0 = clock64();
  float sink = 0.f;
  uint32_t v2[32];
  for (int r = 0; r < ld_reps; ++r) {
    for (int c0 = 0; c0 < N; c0 += 64) {
      tmem_ld32(tmem + lane_base + c0, v);
      tmem_ld32(tmem + lane_base + c0 + 32, v2);
      tmem_wait_ld();
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < 32; j += 2) x ^= (v[j] ^ v[j + 1]) ^ (v2[j] ^ v2[j + 1]);  // tree-able, no serial FP chain
      sink += __uint_as_float(x & 0x3fffffffu);
    }
  }
  __syncthreads();
  long long u1 = clock64();
  if (tid == 0) cyc[1] = u1 - u0;
  if (sink == 1.2345f) D[0] = sink;
  fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int MODE, int N>
static int run(const char* name) {
  const int KD = 64;
  std::vector<float> A(128 * KD), B(N * KD), D(128 * N), R(128 * N);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX) * 2 - 1;
  for (auto& x : B) x = (rand() / (float)RAND_MAX) * 2 - 1;
  float *dA, *dB, *dD; long long* dc;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dc, 16));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const int ES = MODE == 0 ? 4 : 2;
  const size_t smem = (128 + N) * KD * ES;
  auto k = probe_kernel<MODE, N>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<1, 128, smem>>>(dA, dB, dD, 1, dc, 1);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int kk = 0; kk < KD; ++kk) {
        float a = A[r * KD + kk], b = B[n * KD + kk];
        if (MODE == 0) { a = tf32_trunc(a); b = tf32_trunc(b); }
        else { a = __half2float(__float2half_rn(a)); b = __half2float(__float2half_rn(b)); }
        s += (double)a * b;
      }
      maxerr = fmax(maxerr, fabs(s - D[r * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("%s N=%d: max|err| = %.3e (max|ref| %.3f)  %s\n", name, N, maxerr, maxref, maxerr < 2e-3 * maxref ? "OK" : "MISMATCH");
  // timing
  long long c[2];
  const int reps = 2000, ld_reps = 2000;
  k<<<1, 128, smem>>>(dA, dB, dD, reps, dc, ld_reps);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost));
  const int nk = KD * ES / 32;
  printf("   %d MMAs (M=128,N=%d,K=%dB): %.1f cyc/MMA -> %.0f MAC/clk/SM;  tcgen05.ld: %.1f cyc per 32 columns per warp -> %.1f B/clk/SM (4 warps)\n",
         reps * nk, N, 32, (double)c[0] / (reps * nk), 128.0 * N * (32 / ES) / ((double)c[0] / (reps * nk)),
         (double)c[1] / (ld_reps * (N / 32)), 4.0 * 32 * 32 * 4 / ((double)c[1] / (ld_reps * (N / 32))));
  int bad = !(maxerr < 2e-3 * maxref);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc);
  return bad;
}

int main() {
  int bad = 0;
  bad += run<0, 64>("tf32");
  bad += run<0, 128>("tf32");
  bad += run<0, 256>("tf32");
  bad += run<1, 64>("fp16");
  bad += run<1, 256>("fp16");
  bad += run<2, 64>("fp16-MNmajor");
  bad += run<2, 128>("fp16-MNmajor");
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad;
}
