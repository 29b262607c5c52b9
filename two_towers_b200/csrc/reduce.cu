#include "reduce.cuh"

namespace tt {

__global__ void __launch_bounds__(256) reduce_parts_kernel(const ReduceJobs jobs) {
  __shared__ float s_tot[8][128];
  pdl_trigger();
  pdl_wait();
  int k = 0;
#pragma unroll
  for (int t = 1; t < 4; ++t) k += (t < jobs.njobs && (int)blockIdx.x >= jobs.unit_base[t]) ? 1 : 0;
  const ReduceJob& j = jobs.job[k];
  const int unit = (int)blockIdx.x - jobs.unit_base[k];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int width = j.vec ? 128 : 32;                            // outputs per block
  const int64_t base_i = (int64_t)unit * width;
  if (j.vec) {
    // 128 consecutive outputs: each warp sums its slices with 16-byte loads, 4 in flight
    const int64_t i = base_i + lane * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < j.n) {
      const float* base = j.part + i;
      int p = w;
      for (; p + 24 < j.nparts; p += 32) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(base + (int64_t)p * j.stride));
        const float4 b = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(p + 8) * j.stride));
        const float4 c = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(p + 16) * j.stride));
        const float4 d = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(p + 24) * j.stride));
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
        acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      }
      for (; p < j.nparts; p += 8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(base + (int64_t)p * j.stride));
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
    }
    *reinterpret_cast<float4*>(&s_tot[w][lane * 4]) = acc;
  } else {
    const int64_t i = base_i + lane;
    float acc = 0.f;
    if (i < j.n) {
      const float* base = j.part + i;
      int p = w;
      for (; p + 56 < j.nparts; p += 64) {                       // 8 independent loads in flight
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __ldg(base + (int64_t)(p + 8 * u) * j.stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += t[u];
      }
      for (; p + 24 < j.nparts; p += 32) {                       // 4 independent loads in flight
        const float a = __ldg(base + (int64_t)p * j.stride), b = __ldg(base + (int64_t)(p + 8) * j.stride);
        const float c = __ldg(base + (int64_t)(p + 16) * j.stride), d = __ldg(base + (int64_t)(p + 24) * j.stride);
        acc += a; acc += b; acc += c; acc += d;
      }
      for (; p < j.nparts; p += 8) acc += __ldg(base + (int64_t)p * j.stride);
    }
    s_tot[w][lane] = acc;
  }
  __syncthreads();
  const int64_t i = base_i + threadIdx.x;
  if ((int)threadIdx.x < width && i < j.n) {
    float v = s_tot[0][threadIdx.x];
#pragma unroll
    for (int t = 1; t < 8; ++t) v += s_tot[t][threadIdx.x];
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

int reduce_parts(const ReduceJobs& jobs_in, cudaStream_t s) {
  if (jobs_in.njobs <= 0) return TT_OK;
  ReduceJobs jobs = jobs_in;
  int units = 0;
  for (int k = 0; k < jobs.njobs; ++k) {
    ReduceJob& j = jobs.job[k];
    // wide (16-byte) units for big, aligned jobs; 32-output units keep small many-slice jobs spread over more blocks
    j.vec = (j.n >= 32768 && (j.n & 3) == 0 && (j.stride & 3) == 0 && (reinterpret_cast<uintptr_t>(j.part) & 15) == 0) ? 1 : 0;
    jobs.unit_base[k] = units;
    units += (int)ceil_div(j.n, j.vec ? 128 : 32);
  }
  for (int k = jobs.njobs; k < 5; ++k) jobs.unit_base[k] = units;
  if (units == 0) return TT_OK;
  TT_CUDA(launch_kernel(reduce_parts_kernel, dim3((unsigned)units), dim3(256), 0, s, true, jobs));
  TT_LAUNCH_CHECK("reduce_parts_kernel");
  return TT_OK;
}

}  // namespace tt
