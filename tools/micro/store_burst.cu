// store_burst.cu -- developer microbenchmark: how long does ONE burst of 128 CTAs x 128 KB take for
// different epilogue store patterns?  In-kernel %globaltimer stamps (min start / max end over CTAs).
#include <cstdio>
#include <cuda_runtime.h>

__device__ unsigned long long g_t0, g_t1;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// mode 0: per 32-col block, 32 x STG.32 rows 1 KB apart (current epilogue)
// mode 1: thread-per-row float4 stores (each thread writes its own row, 16 B at a time; uncoalesced rows)
// mode 2: warp writes whole rows: 2 x STG.128 per 1 KB row
// mode 3: bulk smem->global 1 KB per row (cp.async.bulk), one lane per row
// mode 4: bulk smem->global 32 KB per warp (contiguous 32 rows)
__global__ void burst(float* out, int mode) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* base = out + (size_t)blockIdx.x * (128 * 256);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sf = reinterpret_cast<float*>(sm);
  for (int i = threadIdx.x; i < 128 * 256; i += blockDim.x) sf[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) atomicMin(&g_t0, gtime());
  __syncthreads();
  float* wbase = base + (size_t)warp * 32 * 256;
  if (mode == 0) {
    for (int cb = 0; cb < 8; ++cb)
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) wbase[rr * 256 + cb * 32 + lane] = sf[(warp * 32 + rr) * 256 + cb * 32 + lane];
  } else if (mode == 1) {
    for (int c = 0; c < 64; ++c)
      reinterpret_cast<float4*>(wbase + lane * 256)[c] = reinterpret_cast<float4*>(sf + (warp * 32 + lane) * 256)[c];
  } else if (mode == 2) {
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr)
      for (int h = 0; h < 2; ++h)
        reinterpret_cast<float4*>(wbase + rr * 256)[h * 32 + lane] = reinterpret_cast<float4*>(sf + (warp * 32 + rr) * 256)[h * 32 + lane];
  } else if (mode == 3) {
    unsigned saddr = (unsigned)__cvta_generic_to_shared(sf + (warp * 32 + lane) * 256);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(wbase + lane * 256), "r"(saddr), "r"(1024) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (mode == 4) {
    if (lane == 0) {
      unsigned saddr = (unsigned)__cvta_generic_to_shared(sf + (warp * 32) * 256);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(wbase), "r"(saddr), "r"(32768) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(&g_t1, gtime());
}

int main() {
  float* buf; cudaMalloc(&buf, (size_t)1 << 30);
  float* fl; cudaMalloc(&fl, (size_t)1 << 28);
  cudaFuncSetAttribute(burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const char* names[] = {"32 x STG.32 per 32-col block (rows 1 KB apart)", "thread-per-row STG.128", "warp-per-row 2 x STG.128", "bulk 1 KB per row", "bulk 32 KB per warp"};
  for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < 5; ++mode) {
      for (int ctas : {64, 128}) {
        cudaMemset(fl, 1, (size_t)1 << 28);          // flush: L2 full of dirty lines
        unsigned long long big = ~0ull, zero = 0;
        cudaMemcpyToSymbol(g_t0, &big, 8); cudaMemcpyToSymbol(g_t1, &zero, 8);
        burst<<<ctas, 128, 128 * 1024>>>(buf + (size_t)(rep * 5 + mode) * 128 * 32768, mode);
        cudaDeviceSynchronize();
        unsigned long long t0, t1; cudaMemcpyFromSymbol(&t0, g_t0, 8); cudaMemcpyFromSymbol(&t1, g_t1, 8);
        printf("%-52s ctas %3d: %6.2f us  (%6.1f GB/s)\n", names[mode], ctas, (t1 - t0) * 1e-3, ctas * 131072.0 / (t1 - t0));
      }
    }
  printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
