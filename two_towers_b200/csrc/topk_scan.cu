// topk_scan.cu -- K7: brute-force dot-product scan of the document index + exact top-k.
//
// Reference: TwoTowerSearch.search, inference/search/two_tower.py:98-102
// (F.cosine_similarity(q[1,1,H], D[1,N,H], dim=2) -> [N]) and :105 (torch.topk).  The
// reference materialises an N x H product plus norms (about 2x the index in temporaries);
// here the index is read exactly once, in 128-bit streaming loads, and nothing of size N is
// written.
//
// Layout / algorithm (HBM-bound: N*H*s bytes per query batch):
//   * index [N,H] row-major fp32 (or bf16); 8 lanes own one row (lane j reads 16-byte chunks
//     j, j+8, ...: a warp instruction covers 4 rows x 128 contiguous bytes), 2 row-quads in
//     flight per warp -> 16 independent LDG.128 per lane.
//   * the query (<= 2 per pass) lives in registers; partial dots are combined with 3 xor
//     shuffles inside the 8-lane group.
//   * top-k: every warp keeps a private candidate buffer in shared memory (CAP >= 2k keys)
//     and a running threshold = score of its current k-th best.  A row is appended only if
//     score >= threshold (ballot + prefix popcount, no atomics); a full buffer is pruned by an
//     in-warp bitonic sort.  Expected appends per warp ~ k (1 + ln(rows_per_warp / k)).
//   * keys are 64-bit: (order-preserving score bits << 32) | ~row, so "larger key" ==
//     "higher score, ties -> LOWER row index" (BASELINE tie rule) and keys are unique.
//   * block epilogue: bitonic sort over all warps' buffers, the block's top-k goes to global;
//     a second kernel (one block per query) radix-selects the k-th key over all block lists
//     and sorts the survivors.  The same kernel implements tt_topk_merge for the
//     row-sharded multi-GPU search.
#include <math_constants.h>

#include "common.cuh"

namespace tt {

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t score_bits(float f) {
  f += 0.0f;                                              // -0 -> +0
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float bits_score(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ u64 make_key(float score, uint32_t row) {
  return ((u64)score_bits(score) << 32) | (u64)(0xffffffffu - row);
}

// descending bitonic sort of buf[0..n) (n power of two) by `nthreads` cooperating threads
template <bool BLOCK>
__device__ __forceinline__ void bitonic_desc(u64* buf, int n, int tid, int nthreads) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (n >> 1); i += nthreads) {
        const int pos = 2 * i - (i & (stride - 1));
        const u64 a = buf[pos], b = buf[pos + stride];
        const bool desc = (pos & size) == 0;
        if ((a < b) == desc) { buf[pos] = b; buf[pos + stride] = a; }
      }
      if (BLOCK) __syncthreads(); else __syncwarp();
    }
  }
}

struct WarpTopK {
  u64* buf;        // [cap] in shared memory
  int cap, k, count;
  float thr;
  __device__ __forceinline__ void init(u64* b, int cap_, int k_) { buf = b; cap = cap_; k = k_; count = 0; thr = -CUDART_INF_F; }
  __device__ __forceinline__ void prune(int lane) {
    __syncwarp();
    for (int i = count + lane; i < cap; i += 32) buf[i] = 0ull;
    __syncwarp();
    bitonic_desc<false>(buf, cap, lane, 32);
    if (count >= k) { count = k; thr = bits_score((uint32_t)(buf[k - 1] >> 32)); }
  }
  // warp-collective append of (score,row) from the lanes with want == true
  __device__ __forceinline__ void push(bool want, float score, uint32_t row, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return;
    if (count + __popc(m) > cap) {
      prune(lane);
      want = want && (score >= thr);
      m = __ballot_sync(0xffffffffu, want);
      if (m == 0) return;
    }
    if (want) buf[count + __popc(m & ((1u << lane) - 1u))] = make_key(score, row);
    count += __popc(m);
  }
};

template <bool BF16> struct ChunkT;
template <> struct ChunkT<false> { typedef float4 type; };
template <> struct ChunkT<true> { typedef uint4 type; };

// dot of one 16-byte index chunk with the matching query registers
__device__ __forceinline__ float dot_chunk(const float4& d, const float* q, float acc) {
  acc = fmaf(d.x, q[0], acc); acc = fmaf(d.y, q[1], acc); acc = fmaf(d.z, q[2], acc); acc = fmaf(d.w, q[3], acc);
  return acc;
}
__device__ __forceinline__ float dot_chunk(const uint4& d, const float* q, float acc) {
  acc = fmaf(bf16lo_to_f32(d.x), q[0], acc); acc = fmaf(bf16hi_to_f32(d.x), q[1], acc);
  acc = fmaf(bf16lo_to_f32(d.y), q[2], acc); acc = fmaf(bf16hi_to_f32(d.y), q[3], acc);
  acc = fmaf(bf16lo_to_f32(d.z), q[4], acc); acc = fmaf(bf16hi_to_f32(d.z), q[5], acc);
  acc = fmaf(bf16lo_to_f32(d.w), q[6], acc); acc = fmaf(bf16hi_to_f32(d.w), q[7], acc);
  return acc;
}
__device__ __forceinline__ float sq_chunk(const float4& d, float acc) {
  acc = fmaf(d.x, d.x, acc); acc = fmaf(d.y, d.y, acc); acc = fmaf(d.z, d.z, acc); acc = fmaf(d.w, d.w, acc);
  return acc;
}
__device__ __forceinline__ float sq_chunk(const uint4& d, float acc) {
  float v;
  v = bf16lo_to_f32(d.x); acc = fmaf(v, v, acc); v = bf16hi_to_f32(d.x); acc = fmaf(v, v, acc);
  v = bf16lo_to_f32(d.y); acc = fmaf(v, v, acc); v = bf16hi_to_f32(d.y); acc = fmaf(v, v, acc);
  v = bf16lo_to_f32(d.z); acc = fmaf(v, v, acc); v = bf16hi_to_f32(d.z); acc = fmaf(v, v, acc);
  v = bf16lo_to_f32(d.w); acc = fmaf(v, v, acc); v = bf16hi_to_f32(d.w); acc = fmaf(v, v, acc);
  return acc;
}
__device__ __forceinline__ float4 ld_chunk(const float4* p) { return ld_stream_f4(p); }
__device__ __forceinline__ uint4 ld_chunk(const uint4* p) { return ld_stream_u4(p); }
__device__ __forceinline__ void zero_chunk(float4& c) { c = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void zero_chunk(uint4& c) { c = make_uint4(0u, 0u, 0u, 0u); }

__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

constexpr int kScanThreads = 512;            // upper bound; large k launches fewer warps so the buffers fit

// CPL = 16-byte chunks per lane (row = 8 lanes x CPL chunks), NQ queries per pass.
template <int CPL, int NQ, bool BF16, bool COSINE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_kernel(const void* __restrict__ index, const float* __restrict__ queries, int64_t N, int H,
                 int k, int cap, int q0, u64* __restrict__ cand) {
  typedef typename ChunkT<BF16>::type Chunk;
  constexpr int EPC = BF16 ? 8 : 4;                        // elements per chunk
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* sm = reinterpret_cast<u64*>(smem_raw);              // [NQ][warps][cap]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kScanWarps = blockDim.x >> 5;
  const int sub = lane & 7, grp = lane >> 3;
  const int chunks_per_row = H / EPC;

  // query fragments -> registers
  float qf[NQ][CPL][EPC];
  float qinv[NQ];
#pragma unroll
  for (int n = 0; n < NQ; ++n) {
    const float* qp = queries + (int64_t)(q0 + n) * H;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int e = 0; e < EPC; ++e) qf[n][c][e] = qp[(sub + 8 * c) * EPC + e];
    qinv[n] = 1.0f;
    if (COSINE) {
      float ss = 0.f;
      for (int e = lane; e < H; e += 32) ss = fmaf(qp[e], qp[e], ss);
      ss = warp_sum(ss);
      qinv[n] = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
    }
  }

  WarpTopK tk[NQ];
#pragma unroll
  for (int n = 0; n < NQ; ++n) tk[n].init(sm + ((size_t)n * kScanWarps + warp) * cap, cap, k);

  const Chunk* base = reinterpret_cast<const Chunk*>(index);
  const int64_t quads = (N + 3) >> 2;
  const int64_t gw = (int64_t)blockIdx.x * kScanWarps + warp;
  const int64_t nw = (int64_t)gridDim.x * kScanWarps;

  for (int64_t it = gw; it < quads; it += 2 * nw) {
    const int64_t rowA = it * 4 + grp, rowB = (it + nw) * 4 + grp;
    const bool okA = rowA < N, okB = (it + nw < quads) && rowB < N;
    Chunk a[CPL], b[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      if (okA) a[c] = ld_chunk(base + rowA * chunks_per_row + sub + 8 * c); else zero_chunk(a[c]);
    }
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      if (okB) b[c] = ld_chunk(base + rowB * chunks_per_row + sub + 8 * c); else zero_chunk(b[c]);
    }
    float invA = 1.0f, invB = 1.0f;
    if (COSINE) {
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) { sa = sq_chunk(a[c], sa); sb = sq_chunk(b[c], sb); }
      invA = 1.0f / fmaxf(sqrtf(group8_sum(sa)), 1e-8f);
      invB = 1.0f / fmaxf(sqrtf(group8_sum(sb)), 1e-8f);
    }
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
      float da = 0.f, db = 0.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) { da = dot_chunk(a[c], qf[n][c], da); db = dot_chunk(b[c], qf[n][c], db); }
      da = group8_sum(da); db = group8_sum(db);
      if (COSINE) { da = da * (qinv[n] * invA); db = db * (qinv[n] * invB); }
      tk[n].push(okA && sub == 0 && da >= tk[n].thr, da, (uint32_t)rowA, lane);
      tk[n].push(okB && sub == 0 && db >= tk[n].thr, db, (uint32_t)rowB, lane);
    }
  }

  // block epilogue: per-warp sort, then sort across the block's warps, emit the block top-k
#pragma unroll
  for (int n = 0; n < NQ; ++n) {
    tk[n].prune(lane);
    __syncthreads();
    u64* region = sm + (size_t)n * kScanWarps * cap;
    bitonic_desc<true>(region, kScanWarps * cap, threadIdx.x, blockDim.x);
    u64* out = cand + ((size_t)(q0 + n) * gridDim.x + blockIdx.x) * k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = region[i];
    __syncthreads();
  }
}

// generic-H fallback: one warp per row, scalar loads (used only when H is not a multiple of 32/64)
template <bool BF16, bool COSINE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_generic_kernel(const void* __restrict__ index, const float* __restrict__ queries, int64_t N, int H,
                         int k, int cap, int q0, u64* __restrict__ cand) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* sm = reinterpret_cast<u64*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kScanWarps = blockDim.x >> 5;
  const float* qp = queries + (int64_t)q0 * H;
  float qinv = 1.0f;
  if (COSINE) {
    float ss = 0.f;
    for (int e = lane; e < H; e += 32) ss = fmaf(qp[e], qp[e], ss);
    qinv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-8f);
  }
  WarpTopK tk;
  tk.init(sm + (size_t)warp * cap, cap, k);
  const int64_t gw = (int64_t)blockIdx.x * kScanWarps + warp, nw = (int64_t)gridDim.x * kScanWarps;
  for (int64_t row = gw; row < N; row += nw) {
    float d = 0.f, ss = 0.f;
    for (int e = lane; e < H; e += 32) {
      float v = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(index)[row * H + e])
                     : reinterpret_cast<const float*>(index)[row * H + e];
      d = fmaf(v, qp[e], d); ss = fmaf(v, v, ss);
    }
    d = warp_sum(d);
    if (COSINE) d = d * qinv / fmaxf(sqrtf(warp_sum(ss)), 1e-8f);
    tk.push(lane == 0 && d >= tk.thr, d, (uint32_t)row, lane);
  }
  tk.prune(lane);
  __syncthreads();
  bitonic_desc<true>(sm, kScanWarps * cap, threadIdx.x, blockDim.x);
  u64* out = cand + ((size_t)q0 * gridDim.x + blockIdx.x) * k;
  for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = sm[i];
}

// ---- merge: one block per query; radix-select the k-th largest key, sort the survivors --------
struct RawKeys {
  const u64* keys; int64_t per_query;
  __device__ __forceinline__ u64 get(int q, int64_t i) const { return keys[(int64_t)q * per_query + i]; }
};
struct BlockedKeys {            // per query: `blocks` buffers of `cap` keys of which the first k are live (tc_topk.cu)
  const u64* keys; int64_t per_query; int k, cap;
  __device__ __forceinline__ u64 get(int q, int64_t i) const { return keys[(int64_t)q * per_query + (i / k) * cap + (i % k)]; }
};
struct ScoreIdKeys {            // R lists of [nq, k]; list r starts at r * rank_stride elements
  const float* scores; const int64_t* ids; int R, nq, k; int64_t score_stride, id_stride;
  __device__ __forceinline__ u64 get(int q, int64_t i) const {
    const int64_t r = i / k, j = i % k;
    const int64_t id = ids[r * id_stride + (int64_t)q * k + j];
    if (id < 0) return 0ull;
    return make_key(scores[r * score_stride + (int64_t)q * k + j], (uint32_t)id);
  }
};

constexpr int kMergeThreads = 1024;

// Exact top-k of L unique 64-bit keys (R sorted lists of k) by an MSB-first radix select, one CTA per query.
// Keys are read from global memory ONCE into registers (up to kMergeRegKeys per thread; longer inputs fall back to
// re-reading), the eight digit passes then only touch shared-memory histograms, and the "which bin holds the k-th key"
// search is a warp-parallel suffix scan instead of a 256-step serial loop.
constexpr int kMergeRegKeys = 16;          // 1024 threads x 16 = 16384 keys in registers (148 block lists x k <= 110)

template <typename Keys>
__global__ void __launch_bounds__(kMergeThreads)
merge_topk_kernel(Keys src, int64_t L, int k, int64_t id_offset, float* __restrict__ out_scores,
                  int64_t* __restrict__ out_ids) {
  __shared__ unsigned int hist[256];
  __shared__ u64 sel[TT_TOPK_MAX];
  __shared__ u64 s_prefix, s_mask;
  __shared__ int s_need, s_count, s_all;
  const int q = blockIdx.x, tid = threadIdx.x;
  const bool in_regs = L <= (int64_t)kMergeRegKeys * kMergeThreads;
  u64 mine[kMergeRegKeys];
  if (in_regs) {
#pragma unroll
    for (int u = 0; u < kMergeRegKeys; ++u) {
      const int64_t i = (int64_t)u * kMergeThreads + tid;
      mine[u] = i < L ? src.get(q, i) : 0ull;               // 0 never matches a live prefix: see below
    }
  }
  if (tid == 0) { s_prefix = 0ull; s_mask = 0ull; s_need = k; s_count = 0; s_all = 0; }
  for (int i = tid; i < TT_TOPK_MAX; i += kMergeThreads) sel[i] = 0ull;
  __syncthreads();
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += kMergeThreads) hist[i] = 0u;
    __syncthreads();
    const u64 prefix = s_prefix, mask = s_mask;
    if (in_regs) {
#pragma unroll
      for (int u = 0; u < kMergeRegKeys; ++u) {
        const u64 key = mine[u];
        if (key != 0ull && (key & mask) == prefix) atomicAdd(&hist[(unsigned)((key >> shift) & 0xffull)], 1u);
      }
    } else {
      for (int64_t i = tid; i < L; i += kMergeThreads) {
        const u64 key = src.get(q, i);
        if (key != 0ull && (key & mask) == prefix) atomicAdd(&hist[(unsigned)((key >> shift) & 0xffull)], 1u);
      }
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns bins [8l, 8l+8).  Walking from bin 255 down, find the bin in which the running count reaches `need`.
      unsigned c[8], lane_sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[8 * tid + j]; lane_sum += c[j]; }
      unsigned above = lane_sum;                               // inclusive suffix sum over lanes tid..31
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_down_sync(0xffffffffu, above, o);
        if (tid + o < 32) above += t;
      }
      const int need = s_need;
      const unsigned excl = above - lane_sum;                   // keys in bins above this lane's
      const bool here = excl < (unsigned)need && above >= (unsigned)need;
      const unsigned who = __ballot_sync(0xffffffffu, here);
      if (who == 0u) {                                          // fewer than k live keys in total (first pass only):
        if (tid == 0) { s_all = 1; s_prefix = 0ull; }           // every live key is selected, the rest is padding
      } else if (here) {
        int rem = need - (int)excl, bin = 7;
        for (; bin > 0; --bin) {
          if ((int)c[bin] >= rem) break;
          rem -= (int)c[bin];
        }
        s_need = rem;
        s_prefix = prefix | ((u64)(8 * tid + bin) << shift);
        s_mask = mask | (0xffull << shift);
        // the chosen bin holds exactly the keys still needed: every key >= prefix (lower digits zero) is selected and the
        // remaining digit passes cannot change that -- stop (typically after the 3rd of 8 passes)
        if ((int)c[bin] == rem) s_all = 1;
      }
    }
    __syncthreads();
    if (s_all) break;                                        // CTA-uniform
  }
  const u64 kth = s_prefix;                                // exact k-th largest key (keys are unique); 0 = take all
  if (in_regs) {
#pragma unroll
    for (int u = 0; u < kMergeRegKeys; ++u) {
      const u64 key = mine[u];
      if (key >= kth && key != 0ull) {
        const int p = atomicAdd(&s_count, 1);
        if (p < TT_TOPK_MAX) sel[p] = key;
      }
    }
  } else {
    for (int64_t i = tid; i < L; i += kMergeThreads) {
      const u64 key = src.get(q, i);
      if (key >= kth && key != 0ull) {
        const int p = atomicAdd(&s_count, 1);
        if (p < TT_TOPK_MAX) sel[p] = key;
      }
    }
  }
  __syncthreads();
  int n2 = 32;
  while (n2 < k) n2 <<= 1;
  bitonic_desc<true>(sel, n2, tid, kMergeThreads);
  for (int i = tid; i < k; i += kMergeThreads) {
    const u64 key = sel[i];
    if (key == 0ull) { out_scores[(int64_t)q * k + i] = -CUDART_INF_F; out_ids[(int64_t)q * k + i] = -1; }
    else {
      out_scores[(int64_t)q * k + i] = bits_score((uint32_t)(key >> 32));
      out_ids[(int64_t)q * k + i] = (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) + id_offset;
    }
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

struct ScanPlan { int cap; int grid; int warps; size_t smem_per_q; size_t cand_bytes; size_t total; };
static ScanPlan plan_scan(int64_t N, int nq, int k) {
  ScanPlan p{};
  int cap = 256;
  while (cap < 2 * k) cap <<= 1;
  p.cap = cap;
  int warps = kScanThreads / 32;
  while (warps > 1 && (size_t)warps * cap * sizeof(u64) > 160 * 1024) warps >>= 1;   // k=1024 -> 8 warps
  p.warps = warps;
  int64_t quads = (N + 3) / 4;
  int64_t g = ceil_div(quads, warps);
  p.grid = (int)(g < kNumSMs ? (g < 1 ? 1 : g) : kNumSMs);
  p.smem_per_q = (size_t)warps * cap * sizeof(u64);
  p.cand_bytes = align_up((size_t)nq * p.grid * k * sizeof(u64));
  p.total = p.cand_bytes + 256;
  return p;
}

template <int CPL, bool BF16, bool COS>
static int launch_scan(const void* index, const float* queries, int64_t N, int H, int nq, int k,
                       const ScanPlan& plan, u64* cand, cudaStream_t s) {
  int q = 0;
  // two queries per pass while the candidate buffers fit in shared memory and both query
  // fragments fit in registers (<= 64 floats)
  constexpr bool kTwoFits = CPL * (BF16 ? 8 : 4) * 2 <= 64;
  if (kTwoFits && 2 * plan.smem_per_q <= 200 * 1024) {
    TT_CUDA(cudaFuncSetAttribute(scan_topk_kernel<CPL, 2, BF16, COS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * plan.smem_per_q)));
    for (; q + 2 <= nq; q += 2) {
      scan_topk_kernel<CPL, 2, BF16, COS><<<plan.grid, plan.warps * 32, 2 * plan.smem_per_q, s>>>(index, queries, N, H, k, plan.cap, q, cand);
      TT_LAUNCH_CHECK("scan_topk_kernel");
    }
  }
  if (q < nq) {
    TT_CUDA(cudaFuncSetAttribute(scan_topk_kernel<CPL, 1, BF16, COS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_per_q));
    for (; q < nq; ++q) {
      scan_topk_kernel<CPL, 1, BF16, COS><<<plan.grid, plan.warps * 32, plan.smem_per_q, s>>>(index, queries, N, H, k, plan.cap, q, cand);
      TT_LAUNCH_CHECK("scan_topk_kernel");
    }
  }
  return TT_OK;
}

template <bool BF16, bool COS>
static int dispatch_scan(const void* index, const float* queries, int64_t N, int H, int nq, int k,
                         const ScanPlan& plan, u64* cand, cudaStream_t s) {
  const int epc = BF16 ? 8 : 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(index) & 15) == 0;
  if (aligned && H % (8 * epc) == 0) {
    const int cpl = H / (8 * epc);
    switch (cpl) {
      case 1: return launch_scan<1, BF16, COS>(index, queries, N, H, nq, k, plan, cand, s);
      case 2: return launch_scan<2, BF16, COS>(index, queries, N, H, nq, k, plan, cand, s);
      case 4: return launch_scan<4, BF16, COS>(index, queries, N, H, nq, k, plan, cand, s);
      case 8: return launch_scan<8, BF16, COS>(index, queries, N, H, nq, k, plan, cand, s);
      default: break;
    }
  }
  TT_CUDA(cudaFuncSetAttribute(scan_topk_generic_kernel<BF16, COS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_per_q));
  for (int q = 0; q < nq; ++q) {
    scan_topk_generic_kernel<BF16, COS><<<plan.grid, plan.warps * 32, plan.smem_per_q, s>>>(index, queries, N, H, k, plan.cap, q, cand);
    TT_LAUNCH_CHECK("scan_topk_generic_kernel");
  }
  return TT_OK;
}

// exact top-k over `blocks` candidate buffers per query, each `cap` keys long with its k best in front (zero keys are
// padding); used by the batched tensor-core scan
int topk_merge_blocked(const u64* cand, int blocks, int cap, int nq, int k, int64_t id_offset, float* out_scores, int64_t* out_ids,
                       cudaStream_t s) {
  BlockedKeys src{cand, (int64_t)blocks * cap, k, cap};
  merge_topk_kernel<BlockedKeys><<<nq, kMergeThreads, 0, s>>>(src, (int64_t)blocks * k, k, id_offset, out_scores, out_ids);
  TT_LAUNCH_CHECK("merge_topk_kernel");
  return TT_OK;
}

}  // namespace tt

extern "C" {

size_t tt_topk_scan_workspace(int64_t N, int H, int nq, int k) {
  (void)H;
  if (N <= 0 || nq <= 0 || k <= 0) return 256;
  return tt::plan_scan(N, nq, k).total;
}

int tt_topk_scan(const void* index, int index_bf16, const float* queries, int64_t N, int H, int nq, int k,
                 int cosine, int64_t id_offset, float* out_scores, int64_t* out_ids, void* workspace,
                 size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(index && queries && out_scores && out_ids && N > 0 && H > 0 && nq > 0, "topk_scan: bad arguments");
  TT_CHECK_ARG(k > 0 && k <= TT_TOPK_MAX && k <= N, "topk_scan: need 0 < k <= min(N, %d) (k=%d, N=%lld)", TT_TOPK_MAX, k, (long long)N);
  TT_CHECK_ARG(N < (1ll << 32), "topk_scan: a shard holds at most 2^32 rows");
  const tt::ScanPlan plan = tt::plan_scan(N, nq, k);
  if (workspace == nullptr || workspace_bytes < plan.total) { tt::set_error("topk_scan: workspace too small (%zu < %zu)", workspace_bytes, plan.total); return TT_ERR_WORKSPACE; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  tt::u64* cand = reinterpret_cast<tt::u64*>(workspace);
  int rc;
  if (index_bf16) rc = cosine ? tt::dispatch_scan<true, true>(index, queries, N, H, nq, k, plan, cand, s)
                              : tt::dispatch_scan<true, false>(index, queries, N, H, nq, k, plan, cand, s);
  else            rc = cosine ? tt::dispatch_scan<false, true>(index, queries, N, H, nq, k, plan, cand, s)
                              : tt::dispatch_scan<false, false>(index, queries, N, H, nq, k, plan, cand, s);
  if (rc) return rc;
  tt::RawKeys src{cand, (int64_t)plan.grid * k};
  tt::merge_topk_kernel<tt::RawKeys><<<nq, tt::kMergeThreads, 0, s>>>(src, (int64_t)plan.grid * k, k, id_offset, out_scores, out_ids);
  TT_LAUNCH_CHECK("merge_topk_kernel");
  return TT_OK;
}

int tt_topk_merge(const float* scores, const int64_t* ids, int R, int nq, int k, int64_t score_rank_stride,
                  int64_t id_rank_stride, float* out_scores, int64_t* out_ids, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(scores && ids && out_scores && out_ids && R > 0 && nq > 0 && k > 0 && k <= TT_TOPK_MAX, "topk_merge: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  tt::ScoreIdKeys src{scores, ids, R, nq, k, score_rank_stride > 0 ? score_rank_stride : (int64_t)nq * k,
                      id_rank_stride > 0 ? id_rank_stride : (int64_t)nq * k};
  tt::merge_topk_kernel<tt::ScoreIdKeys><<<nq, tt::kMergeThreads, 0, s>>>(src, (int64_t)R * k, k, 0, out_scores, out_ids);
  TT_LAUNCH_CHECK("merge_topk_kernel");
  return TT_OK;
}

int tt_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(src && dst && n >= 0, "cast_f32_to_bf16: bad arguments");
  if (n == 0) return TT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t blocks = tt::ceil_div(n, 256);
  if (blocks > 16 * tt::kNumSMs) blocks = 16 * tt::kNumSMs;
  tt::cast_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, (__nv_bfloat16*)dst, n);
  TT_LAUNCH_CHECK("cast_bf16_kernel");
  return TT_OK;
}

}  // extern "C"
