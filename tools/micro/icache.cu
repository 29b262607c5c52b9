// icache.cu -- developer microbenchmark: cost of run-once straight-line code (cold instruction cache) versus the same
// code on its second pass, for 1 / 4 / 8 warps per CTA.
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__device__ __forceinline__ float chain(float a, float b, float c, float d, float x) {
#pragma unroll
  for (int i = 0; i < N; ++i) {          // 4 independent FFMA chains
    a = fmaf(a, x, 1.0f); b = fmaf(b, x, 2.0f); c = fmaf(c, x, 3.0f); d = fmaf(d, x, 4.0f);
  }
  return (a + b) + (c + d);
}

__global__ void k(float* out, long long* st, float x, int reps) {
  float acc = 0.f;
  for (int r = 0; r < reps; ++r) {
    const long long t0 = clock64();
    acc += chain<1024>(acc, 1.f, 2.f, 3.f, x);    // 4096 FFMA, fully unrolled: 64 KB of code
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) st[r] = t1 - t0;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  long long* st; cudaMalloc(&st, 64 * 8);
  long long h[4];
  for (int threads : {32, 128, 256}) {
    k<<<148, threads>>>(out, st, 0.999f, 3);
    cudaDeviceSynchronize();
    cudaMemcpy(h, st, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%3d threads/CTA: 4096 unrolled FFMA: pass 1 %lld cycles (%.2f / instr), pass 2 %lld, pass 3 %lld\n", threads, h[0], h[0] / 4096.0, h[1], h[2]);
  }
  printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
