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
inline ReduceJob make_job(const float* part, int nparts, int64_t n, int64_t stride, float* out) {
  ReduceJob j{};
  j.part = part; j.out = out; j.n = n; j.stride = stride; j.nparts = nparts;
  return j;
}

}  // namespace tt
