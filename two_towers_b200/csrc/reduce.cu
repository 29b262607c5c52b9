#include "reduce.cuh"

namespace tt {

__global__ void __launch_bounds__(256) reduce_parts_kernel(const ReduceJobs jobs) {
  __shared__ float s_tot[8][32];
  pdl_trigger();
  pdl_wait();
  const ReduceJob& j = jobs.job[blockIdx.y];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  if ((int64_t)blockIdx.x * 32 >= j.n) return;                   // block-uniform
  float acc = 0.f;
  if (i < j.n) {
    const float* base = j.part + i;
    int p = w;
    for (; p + 24 < j.nparts; p += 32) {                         // 4 independent loads in flight
      const float a = __ldg(base + (int64_t)p * j.stride), b = __ldg(base + (int64_t)(p + 8) * j.stride);
      const float c = __ldg(base + (int64_t)(p + 16) * j.stride), d = __ldg(base + (int64_t)(p + 24) * j.stride);
      acc += a; acc += b; acc += c; acc += d;
    }
    for (; p < j.nparts; p += 8) acc += __ldg(base + (int64_t)p * j.stride);
  }
  s_tot[w][lane] = acc;
  __syncthreads();
  if (w == 0 && i < j.n) {
    float v = s_tot[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) v += s_tot[k][lane];
    int64_t o = i;
    if (j.ncols > 0) {
      const int64_t r = i / j.ncols;
      const int c = (int)(i % j.ncols);
      if (j.bias) v += j.bias[c];
      if (j.act == 1) v = fmaxf(v, 0.f);
      if (j.mask) v = (j.mask[r * j.ldmask + c] > 0.f) ? v : 0.f;
      o = r * j.ldo + c;
    }
    j.out[o] = v;
  }
}

int reduce_parts(const ReduceJobs& jobs, cudaStream_t s) {
  if (jobs.njobs <= 0) return TT_OK;
  int64_t nmax = 0;
  for (int k = 0; k < jobs.njobs; ++k) nmax = jobs.job[k].n > nmax ? jobs.job[k].n : nmax;
  dim3 grid((unsigned)ceil_div(nmax, 32), (unsigned)jobs.njobs);
  TT_CUDA(launch_kernel(reduce_parts_kernel, grid, dim3(256), 0, s, true, jobs));
  TT_LAUNCH_CHECK("reduce_parts_kernel");
  return TT_OK;
}

}  // namespace tt
