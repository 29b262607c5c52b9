#include "reduce.cuh"

namespace tt {

__global__ void __launch_bounds__(256) reduce_parts_kernel(const ReduceJobs jobs) {
  pdl_trigger();
  pdl_wait();
  reduce_parts_block(jobs, (int)blockIdx.x);
}

int plan_reduce(ReduceJobs& jobs) {
  int units = 0;
  for (int k = 0; k < jobs.njobs; ++k) {
    ReduceJob& j = jobs.job[k];
    // wide (16-byte) units for big, aligned jobs; 32-output units keep small many-slice jobs spread over more blocks
    j.vec = (j.n >= 32768 && (j.n & 3) == 0 && (j.stride & 3) == 0 && (reinterpret_cast<uintptr_t>(j.part) & 15) == 0) ? 1 : 0;
    jobs.unit_base[k] = units;
    units += (int)ceil_div(j.n, j.vec ? 128 : 32);
  }
  for (int k = jobs.njobs; k < 5; ++k) jobs.unit_base[k] = units;
  return units;
}

int reduce_parts(const ReduceJobs& jobs_in, cudaStream_t s) {
  if (jobs_in.njobs <= 0) return TT_OK;
  ReduceJobs jobs = jobs_in;
  const int units = plan_reduce(jobs);
  if (units == 0) return TT_OK;
  TT_CUDA(launch_kernel(reduce_parts_kernel, dim3((unsigned)units), dim3(256), 0, s, true, jobs));
  TT_LAUNCH_CHECK("reduce_parts_kernel");
  return TT_OK;
}

}  // namespace tt
