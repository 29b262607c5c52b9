// reduce.cuh -- deterministic "sum of partial slices" kernel shared by every split reduction in the
// library (split-K GEMM partials, loss-gradient splits, per-block column sums -> bias gradients).
//
//   out[i] = epilogue( sum_p part[p * stride + i] ),  i < n
//
// Fixed summation order (bitwise reproducible): a block owns 32 (or, for large aligned jobs, 128) consecutive
// outputs; its 8 warps each sum the slices p = w, w+8, w+16, ... (independent, coalesced loads, 4 in flight per
// thread -- a naive one-thread-per-output loop serialises on every load), then the warp totals are added in
// warp order.  All jobs of a call share one flat grid (no empty blocks).
#pragma once
#include "common.cuh"

namespace tt {

struct ReduceJob {
  const float* part;     // slices
  float* out;
  int64_t n;             // outputs
  int64_t stride;        // elements between slices
  int nparts;
  // optional GEMM-style epilogue on out (all nullable / 0)
  const float* bias;     // [ncols]
  int ncols;             // row length for bias / mask indexing (0 = no epilogue indexing)
  int act;               // 1 = relu
  const float* mask;     // [n / ncols, ldmask] : out *= (mask > 0)
  int ldmask;
  int ldo;               // output row pitch when ncols > 0 (else contiguous)
  int vec;               // set by reduce_parts: 128-output blocks with 16-byte loads
};
struct ReduceJobs {
  ReduceJob job[4];
  int njobs;
  int unit_base[5];      // set by reduce_parts: first block of each job in the flat grid
};

int reduce_parts(const ReduceJobs& jobs, cudaStream_t s);
int plan_reduce(ReduceJobs& jobs);          // fills vec / unit_base, returns the number of 256-thread blocks
inline ReduceJob make_job(const float* part, int nparts, int64_t n, int64_t stride, float* out) {
  ReduceJob j{};
  j.part = part; j.out = out; j.n = n; j.stride = stride; j.nparts = nparts;
  return j;
}

#ifdef __CUDACC__
// one block (256 threads) of the flat reduction grid; callable from other kernels that append work of their own
__device__ __forceinline__ void reduce_parts_block(const ReduceJobs& jobs, int block) {
  __shared__ float s_tot[8][128];
  int k = 0;
#pragma unroll
  for (int t = 1; t < 4; ++t) k += (t < jobs.njobs && block >= jobs.unit_base[t]) ? 1 : 0;
  const ReduceJob& j = jobs.job[k];
  const int unit = block - jobs.unit_base[k];
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
#endif

}  // namespace tt
