// store_bw.cu -- developer microbenchmark: global-store throughput for the epilogue patterns used by the
// tensor-core kernels (few warps per SM, 128-byte rows 1 KB apart) versus wide / many-warp stores.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_bw store_bw.cu && ./store_bw
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>
__global__ void store_kernel(float* out, size_t floats_per_cta, int iters, int row_floats, size_t iter_stride = 0) {
  // each warp owns consecutive 32-row groups; a store instruction writes 32*VEC floats of one row
  float* base = out + (size_t)blockIdx.x * floats_per_cta;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t rows = floats_per_cta / row_floats;
  const int segs = row_floats / (32 * VEC);
  for (int it = 0; it < iters; ++it, base += iter_stride) {
    for (size_t r0 = (size_t)warp * 32; r0 < rows; r0 += (size_t)nw * 32) {
      for (int sg = 0; sg < segs; ++sg) {
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          float* p = base + (r0 + rr) * row_floats + sg * 32 * VEC + lane * VEC;
          if (VEC == 1) *p = (float)it;
          else *reinterpret_cast<float4*>(p) = make_float4(it, it, it, it);
        }
      }
    }
  }
}

__global__ void bulk_store_kernel(float* out, size_t floats_per_cta, int iters) {
  // TMA-style bulk copies smem -> global, 16 KB per instruction, one thread issues
  extern __shared__ __align__(128) unsigned char sm[];
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    char* base = reinterpret_cast<char*>(out + (size_t)blockIdx.x * floats_per_cta);
    const size_t bytes = floats_per_cta * 4;
    unsigned saddr = (unsigned)__cvta_generic_to_shared(sm);
    for (int it = 0; it < iters; ++it) {
      for (size_t off = 0; off < bytes; off += 16384) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(saddr), "r"(16384) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F>
static float time_it(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  const int iters = 50;
  float* buf; cudaMalloc(&buf, (size_t)1 << 30);
  for (int per_cta_kb : {128, 1024}) {
    for (int ctas : {64, 128, 148, 296}) {
      const size_t fpc = (size_t)per_cta_kb * 1024 / 4;
      const double gb = (double)ctas * per_cta_kb * 1024 * iters / 1e9;
      for (int threads : {128, 256, 1024}) {
        float m1 = time_it([&] { store_kernel<1><<<ctas, threads>>>(buf, fpc, iters, 256); });
        float m4 = time_it([&] { store_kernel<4><<<ctas, threads>>>(buf, fpc, iters, 256); });
        printf("ctas %3d x %4d KB  threads %4d : STG.32 %7.1f GB/s   STG.128 %7.1f GB/s\n", ctas, per_cta_kb, threads, gb / (m1 * 1e-3), gb / (m4 * 1e-3));
      }
      cudaFuncSetAttribute(bulk_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
      float mb = time_it([&] { bulk_store_kernel<<<ctas, 128, 16384>>>(buf, fpc, iters); });
      printf("ctas %3d x %4d KB  bulk smem->global              : %7.1f GB/s\n", ctas, per_cta_kb, gb / (mb * 1e-3));
    }
  }
  // streaming writes: every iteration targets a fresh 16 MB region (lines not resident in L2)
  {
    const int ctas = 128; const size_t fpc = 128 * 1024 / 4; const size_t stride = (size_t)ctas * fpc;
    for (int it_n : {1, 4, 16, 60}) {
      const double gb = (double)ctas * 128 * 1024 * it_n / 1e9;
      cudaMemset(buf, 0, (size_t)1 << 30);     // dirty lines everywhere
      cudaDeviceSynchronize();
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a); store_kernel<1><<<ctas, 128>>>(buf, fpc, it_n, 256, stride); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("fresh lines, %2d x 16 MB, 128 ctas x 128 thr STG.32: %7.1f GB/s (%.1f us)\n", it_n, gb / (ms * 1e-3), ms * 1e3);
    }
  }
  printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
