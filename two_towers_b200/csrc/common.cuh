// common.cuh -- shared device/host helpers for libtt_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tt_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtt_b200 is written for sm_100a (B200) only"
#endif

namespace tt {

// ---- error plumbing (thread-local message, no exceptions across the C ABI) ----------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int require_device();                       // TT_OK iff current device is sm_100
void count_launch(int n = 1);
long long* bad_id_word();                   // device alias of the mapped host word that records an out-of-range token id (may be null)

#define TT_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      ::tt::set_error(__VA_ARGS__);                               \
      return TT_ERR_INVALID;                                      \
    }                                                             \
  } while (0)

#define TT_CUDA(call)                                             \
  do {                                                            \
    int _rc = ::tt::check_cuda((call), #call);                    \
    if (_rc != TT_OK) return _rc;                                 \
  } while (0)

#define TT_LAUNCH_CHECK(name)                                     \
  do {                                                            \
    ::tt::count_launch();                                         \
    int _rc = ::tt::check_cuda(cudaGetLastError(), name);         \
    if (_rc != TT_OK) return _rc;                                 \
  } while (0)

#define TT_REQUIRE_DEVICE()                                       \
  do {                                                            \
    int _rc = ::tt::require_device();                             \
    if (_rc != TT_OK) return _rc;                                 \
  } while (0)

__host__ __device__ inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// bump allocator over a caller-provided workspace
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (base == nullptr || off + bytes > size) return nullptr;
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

constexpr int kNumSMs = 148;    // B200: 2 dies x 74 SMs

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Kernels on the training path call pdl_trigger() as their first instruction and pdl_wait() before their
// first access to global memory produced by the previous kernel.  Launched with the programmatic-stream-
// serialization attribute, kernel N+1 is scheduled as soon as every CTA of kernel N has started, so its launch
// latency and prologue (barrier init, TMEM allocation, tensor-map prefetch) overlap the tail of kernel N.
// Dependents launch only after ALL CTAs of the primary have triggered, so no primary CTA can be starved.
bool pdl_enabled();             // env TT_PDL=0 switches the attribute off (plain stream order)

// cluster_y > 1: thread-block clusters of (1, cluster_y, 1) CTAs (grid.y must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                                         unsigned cluster_y, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl && pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_y > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = cluster_y; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl && pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    na = 1;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- device helpers --------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// streaming 128-bit load that does not pollute L1 (index scan, read-once data)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t packed) {
  return __half22float2(*reinterpret_cast<const __half2*>(&packed));
}
#endif

}  // namespace tt
