// Micro-benchmark: how fast can every SM stream the SAME L2-resident weight image (32 KB chunks, cp.async.bulk)
// with D chunks in flight?  Models the weight stream of nif_tc_fwd / nif_tc_bwd_data (68 chunks of 32 KB per tile pair).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream tools/l2_stream.cu && ./l2_stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
                   smem_u32(bar)),
               "r"(parity)
               : "memory");
}

// depth D loads in flight; chunk sequence c = (i + rot * blockIdx.x) % nchunks
__global__ void k(const float* img, int nchunks, int chunk_bytes, int iters, int D, int rot, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    auto issue = [&](int i) {
      const int s = i % D;
      const int c = (i + rot * (int)blockIdx.x) % nchunks;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(chunk_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem + (size_t)s * chunk_bytes)),
                   "l"(reinterpret_cast<const unsigned char*>(img) + (size_t)c * chunk_bytes), "r"(chunk_bytes), "r"(smem_u32(&bar[s]))
                   : "memory");
    };
    for (int i = 0; i < D && i < iters; ++i) issue(i);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&bar[i % D], (uint32_t)((i / D) & 1));
      if (i + D < iters) issue(i + D);
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int nchunks = 68;
  float* img;
  long long* out;
  cudaMalloc(&img, (size_t)nchunks * 32768);
  cudaMemset(img, 0, (size_t)nchunks * 32768);
  cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 68 * 8;
  for (int bytes = 32768; bytes >= 8192; bytes /= 2)
    for (int rot = 0; rot <= 1; ++rot)
      for (int D = 1; D <= 6; D += (D < 4 ? 1 : 2)) {
        if ((size_t)D * bytes > 196608) continue;
        const int nch = nchunks * (32768 / bytes);
        k<<<148, 32, (size_t)D * bytes>>>(img, nch, bytes, iters, D, rot, out);
        k<<<148, 32, (size_t)D * bytes>>>(img, nch, bytes, iters, D, rot, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("chunk %5d B  depth %d  rot %d: %8.1f cycles/chunk (slowest SM) = %6.1f B/cycle/SM  %s\n", bytes, D, rot,
               (double)mx / iters, (double)bytes * iters / mx, cudaGetErrorString(e));
      }
  return 0;
}
