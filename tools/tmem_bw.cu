// Micro-benchmark: tcgen05.ld throughput / latency per SM (how fast can epilogue warps drain TMEM?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tools/tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]);
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// mode 0: every load followed by wait (latency-exposed); mode 1: 4 loads in flight before each wait
template <int X, int MODE>
__global__ void k(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t a[32], b[32], c[32], d[32];
    if (MODE == 0) {
      ld<X>(tm + (uint32_t)((i * 4 * X) & 255), a);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += a[0] + a[X - 1];
    } else {
      const uint32_t base = (uint32_t)((i * 4 * X) & 255);
      ld<X>(tm + base, a);
      ld<X>(tm + ((base + X) & 511), b);
      ld<X>(tm + ((base + 2 * X) & 511), c);
      ld<X>(tm + ((base + 3 * X) & 511), d);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += a[0] + b[X - 1] + c[1] + d[2];
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int X, int MODE>
void run(int threads, const char* name) {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 4096;
  k<X, MODE><<<148, threads>>>(iters, out, sink);
  k<X, MODE><<<148, threads>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const int loads = MODE ? 4 : 1;
  const double cyc = (double)h[0] / iters;
  const double bytes = (double)loads * X * 128.0 * (threads / 32);  // per iteration per SM
  printf("%-28s warps/SM %2d  x%-2d  %7.1f cycles/iter  -> %7.1f B/cycle/SM  (%s)\n", name, threads / 32, X, cyc, bytes / cyc,
         cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(sink);
}

int main() {
  run<32, 0>(32, "1 load then wait");
  run<16, 0>(32, "1 load then wait");
  run<32, 1>(32, "4 loads then wait");
  run<32, 1>(128, "4 loads then wait");
  run<32, 1>(256, "4 loads then wait");
  run<16, 1>(128, "4 loads then wait");
  run<16, 1>(256, "4 loads then wait");
  run<32, 0>(128, "1 load then wait");
  run<32, 0>(256, "1 load then wait");
  return 0;
}
