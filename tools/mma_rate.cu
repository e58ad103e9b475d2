// Micro-benchmark: tcgen05.mma (kind::f16, M=128, K=16 per instruction, operands in SMEM, no-swizzle K-major core-matrix
// layout as used by nif_tc_*.cu) -- sustained rate and commit->wake latency for N = 64 / 128 / 256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
                   smem_u32(bar)),
               "r"(parity)
               : "memory");
}

// per "chunk": 12 MMAs (4 k-steps x 3 products) on [128 x 64] x [N x 64] tiles; `per_wait` chunks between waits
__global__ void k(int N, int iters, int per_wait, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0 && lane == 0) {
    const uint32_t tm = slot;
    const uint64_t a_hi = make_desc(smem_u32(smem), 1024), a_lo = make_desc(smem_u32(smem + 16384), 1024);
    const uint64_t b_hi = make_desc(smem_u32(smem + 32768), 1024), b_lo = make_desc(smem_u32(smem + 32768 + 32768), 1024);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int c = 0; c < per_wait; ++c) {
        const uint32_t d1 = tm + (uint32_t)((c & 1) * 2 * (N > 128 ? 0 : N)), d2 = d1 + (N > 128 ? 256 : N);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 16);
          mma(d2, a_lo + adv, b_hi + adv, idesc, ks > 0);
          mma(d2, a_hi + adv, b_lo + adv, idesc, 1);
          mma(d1, a_hi + adv, b_hi + adv, idesc, ks > 0);
        }
      }
      commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int Ns[3] = {64, 128, 256};
  const int pw[4] = {1, 2, 8, 32};
  for (int n = 0; n < 3; ++n)
    for (int p = 0; p < 4; ++p) {
      const int iters = 2048 / pw[p];
      k<<<148, 128, 98304>>>(Ns[n], iters, pw[p], out);
      k<<<148, 128, 98304>>>(Ns[n], iters, pw[p], out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      const double per_chunk = (double)h[0] / (iters * pw[p]);
      const double ideal = 12.0 * 128.0 * Ns[n] / 256.0;
      printf("N=%3d chunks/wait %2d: %8.1f cycles per 12-MMA chunk (ideal %6.1f, %5.1f%% of tensor peak)  %s\n", Ns[n], pw[p],
             per_chunk, ideal, 100.0 * ideal / per_chunk, cudaGetErrorString(e));
    }
  return 0;
}
