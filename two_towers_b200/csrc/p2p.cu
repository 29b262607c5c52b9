// p2p.cu -- all-gather over NVLink peer memory, one kernel per exchange.
//
// The data-parallel training step has three exchanges per step: the bf16 tower outputs [Q_r | D_r] of every rank
// (global in-batch negatives), the per-query logsumexp, and the gradients.  Each is a few KB to a few MB, so a
// collective's fixed cost dominates its wire time.  Here every rank owns one cudaMalloc'ed exchange buffer
// (header + world slots), maps every peer's buffer through CUDA IPC, and ONE kernel does the whole exchange:
//   * each CTA reads its slice of the local source once and stores it into slot `rank` of EVERY rank's buffer
//     (16-byte stores straight over NVLink / NVSwitch, the local copy included),
//   * fence.sys, then one system-scope atomic per peer bumps that peer's arrival counter for this rank,
//   * the CTA then spins (ld.acquire.sys) until every rank's counter in the LOCAL header has reached this round's
//     target -- when the kernel retires, the local buffer holds all world slots and the next kernel in the stream can
//     read it.  No host involvement, no second launch, capturable in a CUDA graph.
// Counters only ever grow (round r expects r * ctas arrivals per peer), so nothing has to be reset between rounds;
// the round number lives in the local header and is advanced by the last CTA to leave.  The CTA count is a property of
// the EXCHANGE (tt_p2p_t.ctas, fixed when it is created), not of a call: calls may move any number of bytes up to the
// slot size and the arrival arithmetic stays consistent on every rank.
// A peer that never arrives (it died, or sits in a long host-side phase) must not hang the GPU or poison the CUDA
// context: after tt_p2p_t.timeout_s seconds (default 600) the waiting CTA records 1 + the missing rank in header word 34
// and returns; the host reads it with tt_p2p_status() at its next synchronisation point and raises.
// Re-use of a slot by a fast rank cannot overtake a slow reader because every step ends with the gradient exchange:
// a rank can only leave step k after all ranks have pushed their step-k gradients, i.e. after they are done reading
// the step-k gather buffers (stream order).  The gradient exchange itself alternates between two slot sets
// (double_buffered): round k+2 re-uses the slots of round k, and nobody can push round k+2 before everybody has
// pushed round k+1, i.e. has finished summing round k.
// Gradients: all-gather + a fixed rank-order sum on every rank (tt_p2p_sum_slots) instead of an all-reduce: the summed
// gradient is bitwise identical on all ranks.
#include <stdlib.h>

#include "common.cuh"

namespace tt {

constexpr int kP2PMaxWorld = 8;
constexpr size_t kP2PHeaderBytes = 256;   // u32 words: [0,32) arrival counters per source rank, [32] round, [33] exit ticket, [34] time-out, [36] CTAs started

struct P2PArgs {
  int world, rank, halves;                // halves == 2: rounds alternate between two slot sets (see tt_p2p_t.double_buffered)
  unsigned timeout_s;
  unsigned char* base[kP2PMaxWorld];      // every rank's exchange buffer as mapped in this process
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned* p, unsigned v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256)
p2p_allgather_kernel(const P2PArgs a, const uint4* __restrict__ src, size_t n16, size_t slot_bytes) {
  unsigned* hdr = reinterpret_cast<unsigned*>(a.base[a.rank]);
  // header word 36 counts the CTAs that have STARTED, over all rounds.  It is bumped before the dependent-launch trigger:
  // a kernel launched programmatically behind this one (tt_inbatch_ce_fwd_dq_p2p) starts only after every CTA here has
  // triggered, so what it reads from word 36 is exactly this round's arrival target (round x CTAs) -- it can then consume
  // the slots rank by rank as their counters reach that target, while this kernel is still pushing / waiting for the
  // slowest peer.  The trigger comes AFTER griddepcontrol.wait: such a dependent also knows that everything before this
  // kernel in the stream has retired (it may read `src` itself).
  pdl_wait();                                                        // the producers of `src` have retired ...
  if (threadIdx.x == 0) { atomicAdd(hdr + 36, 1u); __threadfence(); }
  __syncthreads();
  pdl_trigger();                                                     // ... before any dependent of this kernel may start
  const unsigned round = hdr[32] + 1;                                // every CTA reads it before any CTA can bump it
  const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
  const size_t lo = (size_t)blockIdx.x * per, hi = min(n16, lo + per);
  const size_t half_off = (a.halves == 2 && (round & 1u)) ? (size_t)a.world * slot_bytes : 0;
  const size_t off16 = (kP2PHeaderBytes + half_off + (size_t)a.rank * slot_bytes) / 16;
  for (size_t i0 = lo; i0 < hi; i0 += 256 * 4) {                      // 4 loads in flight per thread, stored to every rank
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = i0 + u * 256 + threadIdx.x;
      if (i < hi) v[u] = __ldg(src + i);
    }
    for (int p = 0; p < a.world; ++p) {
      uint4* dst = reinterpret_cast<uint4*>(a.base[p]) + off16;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t i = i0 + u * 256 + threadIdx.x;
        if (i < hi) dst[i] = v[u];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.world)                                          // one arrival per CTA on every rank's counter for this source
    red_release_sys_add(reinterpret_cast<unsigned*>(a.base[threadIdx.x]) + a.rank, 1u);
  if (threadIdx.x < a.world) {
    const unsigned target = round * gridDim.x;                        // every rank launches the same grid for this exchange
    const unsigned* c = hdr + threadIdx.x;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(c) < target) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > (unsigned long long)a.timeout_s * 1000000000ull) {   // give up, tell the host, keep the context alive
        hdr[34] = 1u + threadIdx.x;
        __threadfence_system();
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(hdr + 33, 1u);
    if (t == gridDim.x - 1) { hdr[33] = 0u; hdr[32] = round; }
  }
}

// out[i] = sum over ranks (rank order) of slot_r[i]  -- the gradient "all-reduce" after an all-gather
__global__ void __launch_bounds__(256)
p2p_sum_slots_kernel(const unsigned char* __restrict__ base, int halves, int world, size_t slot_floats, size_t n, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const unsigned round = reinterpret_cast<const unsigned*>(base)[32];          // the round that has just completed
  const float* slots = reinterpret_cast<const float*>(base + kP2PHeaderBytes) + ((halves == 2 && (round & 1u)) ? (size_t)world * slot_floats : 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v[kP2PMaxWorld];
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r) v[r] = r < world ? __ldg(slots + (size_t)r * slot_floats + i) : 0.f;
    float s = v[0];
#pragma unroll
    for (int r = 1; r < kP2PMaxWorld; ++r) s += v[r];
    out[i] = s;
  }
}

}  // namespace tt

extern "C" {

size_t tt_p2p_buffer_bytes(int world, size_t slot_bytes, int double_buffered) {
  return tt::kP2PHeaderBytes + (size_t)(double_buffered ? 2 : 1) * (size_t)(world > 0 ? world : 1) * tt::align_up(slot_bytes, 256);
}

int tt_p2p_alloc(size_t bytes, void** ptr) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(ptr && bytes >= tt::kP2PHeaderBytes, "p2p_alloc: bad arguments");
  TT_CUDA(cudaMalloc(ptr, bytes));
  TT_CUDA(cudaMemset(*ptr, 0, bytes));
  TT_CUDA(cudaDeviceSynchronize());
  return TT_OK;
}

int tt_p2p_free(void* ptr) {
  if (ptr) TT_CUDA(cudaFree(ptr));
  return TT_OK;
}

int tt_p2p_export(void* ptr, void* handle64) {
  TT_CHECK_ARG(ptr && handle64, "p2p_export: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  TT_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr));
  return TT_OK;
}

int tt_p2p_import(const void* handle64, void** ptr) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(ptr && handle64, "p2p_import: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  TT_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return TT_OK;
}

int tt_p2p_unimport(void* ptr) {
  if (ptr) TT_CUDA(cudaIpcCloseMemHandle(ptr));
  return TT_OK;
}

int tt_p2p_allgather_ctas(size_t bytes) {
  const size_t per_cta = 256 * 16 * 8;                                 // 32 KB per CTA
  size_t c = (bytes + per_cta - 1) / per_cta;
  if (c < 1) c = 1;
  // leave SMs for whatever overlaps.  Measured at 8 ranks (bench.py, 20 steps): 64 / 148 / 296 CTAs give 390.5 / 391.0 / 391.3 us per
  // step -- the exchanges are bound by their fences and by the wait for the slowest rank, not by store issue.  TT_P2P_CTAS
  // overrides the cap (same value on every rank!).
  static const size_t cap = [] { const char* e = getenv("TT_P2P_CTAS"); const long v = e ? atol(e) : 0; return (size_t)(v > 0 ? v : 64); }();
  if (c > cap) c = cap;
  return (int)c;
}

int tt_p2p_status(const tt_p2p_t* x, int* timed_out_rank) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(x && timed_out_rank && x->rank >= 0 && x->rank < tt::kP2PMaxWorld && x->base[x->rank], "p2p_status: bad arguments");
  unsigned w = 0;
  TT_CUDA(cudaMemcpy(&w, static_cast<const unsigned*>(x->base[x->rank]) + 34, sizeof(w), cudaMemcpyDeviceToHost));
  *timed_out_rank = w ? (int)w - 1 : -1;
  return TT_OK;
}

int tt_p2p_allgather(const tt_p2p_t* x, const void* src, size_t bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(x && src && x->world >= 1 && x->world <= tt::kP2PMaxWorld && x->rank >= 0 && x->rank < x->world,
               "p2p_allgather: bad arguments");
  TT_CHECK_ARG(bytes % 16 == 0 && bytes <= x->slot_bytes && x->slot_bytes % 256 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
               "p2p_allgather: bytes must be a multiple of 16 and fit the slot; src 16-byte aligned");
  tt::P2PArgs a{};
  a.world = x->world; a.rank = x->rank; a.halves = x->double_buffered ? 2 : 1;
  a.timeout_s = x->timeout_s > 0 ? (unsigned)x->timeout_s : 600u;
  for (int p = 0; p < x->world; ++p) {
    TT_CHECK_ARG(x->base[p] != nullptr, "p2p_allgather: peer %d not mapped", p);
    a.base[p] = static_cast<unsigned char*>(x->base[p]);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the grid belongs to the exchange, not to this call's byte count (see the header comment)
  const int ctas = x->ctas > 0 ? x->ctas : tt_p2p_allgather_ctas(x->slot_bytes);
  TT_CUDA(tt::launch_kernel(tt::p2p_allgather_kernel, dim3((unsigned)ctas), dim3(256), 0, s, true, a, static_cast<const uint4*>(src),
                            bytes / 16, x->slot_bytes));
  TT_LAUNCH_CHECK("p2p_allgather_kernel");
  return TT_OK;
}

int tt_p2p_sum_slots(const tt_p2p_t* x, size_t n_floats, float* out, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(x && out && x->world >= 1 && x->world <= tt::kP2PMaxWorld && n_floats * 4 <= x->slot_bytes, "p2p_sum_slots: bad arguments");
  if (n_floats == 0) return TT_OK;
  size_t blocks = (n_floats + 255) / 256;
  if (blocks > 4 * (size_t)tt::kNumSMs) blocks = 4 * tt::kNumSMs;
  TT_CUDA(tt::launch_kernel(tt::p2p_sum_slots_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), true,
                            static_cast<const unsigned char*>(x->base[x->rank]), x->double_buffered ? 2 : 1, x->world, x->slot_bytes / 4, n_floats, out));
  TT_LAUNCH_CHECK("p2p_sum_slots_kernel");
  return TT_OK;
}

}  // extern "C"
