// MUFU.SIN / MUFU.EX2 / F2FP issue rates per SM per clock (B200): nvcc -arch=sm_100a -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
template <int OP>
__global__ void k(float* out, int iters, long long* clk) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3f + i;
  unsigned pk = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) v[i] = __sinf(v[i]);                       // FMUL + MUFU.SIN
      if (OP == 1) v[i] = exp2f(v[i]) ;                        // MUFU.EX2 (+ range handling)
      if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) { __nv_bfloat162 h = __floats2bfloat162_rn(v[i], v[(i + 1) & 7]); pk ^= *reinterpret_cast<unsigned*>(&h); v[i] += 1.f; }
      if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 5) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + pk;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int OP>
void run(const char* name, int threads) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 4096;
  k<OP><<<148, threads>>>(out, iters, clk); cudaDeviceSynchronize();
  k<OP><<<148, threads>>>(out, iters, clk); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads/SM %4d : %.2f ops/clk/SM\n", name, threads, (double)threads * iters * 8 / c);
}
int main() {
  for (int th : {128, 512, 1024}) {
    run<0>("__sinf (FMUL+MUFU.SIN)", th); run<2>("ex2.approx", th); run<4>("rcp.approx", th); run<5>("tanh.approx", th);
    run<3>("F2FP.BF16 pack (+FADD)", th);
  }
  return 0;
}
