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
//     j, j+8, ...: a warp instruction covers 4 rows x 128 contiguous bytes), 2 or 4 row-quads in
//     flight per warp -> 16 independent LDG.128 per lane for fp32 x 256 and bf16 x 256 columns alike.
//   * the query (<= 2 per pass) lives in registers; partial dots are combined with 3 xor
//     shuffles inside the 8-lane group.
//   * top-k: every warp keeps a private candidate buffer in shared memory (CAP >= 2k keys)
//     and a running threshold.  A row is appended only if score >= threshold (ballot + prefix
//     popcount, no atomics); a full buffer is pruned by an in-warp bitonic sort.  The threshold
//     STARTS at the k-th best score of a CTA-wide sample (the first 64 rows of every warp, scored
//     unconditionally): a warp's own k best rows give a weak bound on a small shard.
//   * keys are 64-bit: (order-preserving score bits << 32) | ~row, so "larger key" ==
//     "higher score, ties -> LOWER row index" (BASELINE tie rule) and keys are unique.
//   * selections over a few hundred keys (sample threshold, block epilogue, merge of the block
//     lists, merge of the shards' lists) all work the same way: sorted runs (a warp sorts its 64
//     keys in registers with shuffles), the k-th largest of the runs' HEADS as a lower bound of the
//     k-th largest key, and a rank computation over the keys that reach it (one key per thread,
//     broadcast reads of shared memory) -- a handful of barriers instead of a CTA-wide bitonic sort
//     or radix passes over everything.  The bitonic / radix code remains as the fallback for shapes
//     that do not fit (k > 128, ties that flood the survivors, tiny shards).
//   * a second kernel (one block per query, launched programmatically behind the scan) selects
//     over the block lists; on a row-sharded index the same launch also exchanges the shard's k
//     keys with every rank over NVLink peer memory and selects the global top-k
//     (merge_exchange_merge_kernel, tt_topk_scan_p2p).  tt_topk_merge (lists from anywhere) keeps
//     the radix select.
#include <math_constants.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <type_traits>

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

// developer aid (TT_SCAN_DEBUG=1, tt_topk_scan prints them): %globaltimer stamps of CTA 0 -- start, sample scored, threshold
// known, main loop done (warp 0), packed + sorted, end
__device__ int g_scan_dbg_on = 0;
__device__ long long g_scan_dbg[8];
#define TT_SCAN_STAMP(i) do { if (dbg) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_scan_dbg[i] = t_; } } while (0)

// fp32 rows of 256+ columns (16 x 16-byte loads in flight per lane): rounds of the scan (one round = one step of every warp
// of a CTA = warps * 8 rows) are handed out to the CTAs through a global counter -- SMs do not stream at the same rate, and
// with a fixed share per CTA the kernel waited 15-19 us for the slowest one (10 M rows: 1433 -> 1397 us, 7.3 TB/s).  The
// narrower variants have half the bytes in flight and are latency-bound: the extra hop per step cost them more than the
// balance gained (bf16 x 256, 10 M rows: 766 -> 823 us with two quads per step, 730 -> 753 us with four, same box), so they
// keep the static interleaved schedule.
// One counter pair {next round, CTAs finished} per launch slot; the last CTA to finish resets its pair, so the counters are
// zero between launches (they start zero: static storage).  Launches take slots round-robin.
// Word 2: set by the CTA that drew round 0; the last CTA traps if nobody did (the counter was not zero when the launch began,
// i.e. an earlier launch on this slot died half-way: rows would be skipped silently otherwise).
constexpr int kScanSlots = 256;
__device__ unsigned g_scan_rounds[kScanSlots][4];

constexpr int kScanThreads = 512;            // upper bound; large k launches fewer warps so the buffers fit
constexpr int kSampleRows = 64;              // rows every warp scores unconditionally before the CTA fixes its starting threshold
constexpr int kMaxCapPerLane = 8;            // packed block epilogue: candidate buffers of <= 256 keys travel through registers
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// descending bitonic sort of 32 * KPL keys held KPL per lane (element e = j * 32 + lane): compare distances below 32 are
// shuffles, the others pair registers of one lane -- no shared memory, no barriers
template <int KPL>
__device__ __forceinline__ void warp_sort_desc(u64 (&v)[KPL], int lane) {
  constexpr int M = 32 * KPL;
#pragma unroll
  for (int size = 2; size <= M; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int sj = stride >> 5;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          if ((j & sj) == 0) {
            const bool desc = ((j << 5) & size) == 0;
            const u64 a = v[j], b = v[j | sj];
            if ((a < b) == desc) { v[j] = b; v[j | sj] = a; }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          const u64 other = __shfl_xor_sync(0xffffffffu, v[j], stride);
          const bool desc = ((((j << 5) | lane)) & size) == 0;
          const bool keep_max = ((lane & stride) == 0) == desc;
          v[j] = keep_max ? (v[j] > other ? v[j] : other) : (v[j] < other ? v[j] : other);
        }
      }
    }
  }
}
// sorts buf[0..cnt) of one warp (cnt <= 32 * KPL, unique keys) descending in place, zero-padded to 32 * KPL
template <int KPL>
__device__ __forceinline__ void warp_sort_buffer(u64* buf, int cnt, int lane) {
  u64 v[KPL];
#pragma unroll
  for (int j = 0; j < KPL; ++j) v[j] = (j * 32 + lane < cnt) ? buf[j * 32 + lane] : 0ull;
  warp_sort_desc<KPL>(v, lane);
#pragma unroll
  for (int j = 0; j < KPL; ++j) buf[j * 32 + lane] = v[j];
}

// The whole CTA: top-k of W DESCENDING runs (run w: lst + w * stride, s_cnt[w] keys) -> out[0..k) descending, zero-padded.
// Same idea as select_from_lists: the k-th largest of the runs' heads (first ceil(k / W) keys of each) bounds the k-th largest
// key from below; the few keys that reach it are ranked against each other (one key per thread, broadcast reads) -- four
// barriers instead of the 55 of a 1024-key bitonic sort.  `out` may alias run 0.  Returns the number of live keys written,
// or -1 with nothing written when the shape / data do not fit (fewer than k live heads, more survivors than threads).
// deal_stride > 0: the selected keys are dealt round-robin into W buffers deal_stride keys apart instead of one sorted run.
__device__ __forceinline__ int cta_select_runs(const u64* lst, int stride, const int* s_cnt, int W, int k, u64* out,
                                               u64* scratch, int* s_ctr, u64* s_thr, int deal_stride = 0) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int r = (k + W - 1) / W, nh = W * r;
  if (nh > nthr) return -1;                                  // CTA-uniform
  u64 head = 0ull;
  if (tid < nh) {
    const int w = tid / r, pos = tid - w * r;
    head = pos < s_cnt[w] ? lst[(size_t)w * stride + pos] : 0ull;
    scratch[tid] = head;
  }
  if (tid == 0) { *s_ctr = 0; *s_thr = 0ull; }
  __syncthreads();
  if (tid < nh && head != 0ull) {
    int rank = 0;
    for (int j = 0; j < nh; ++j) rank += scratch[j] > head;
    if (rank == k - 1) *s_thr = head;
  }
  __syncthreads();
  const u64 thr = *s_thr;
  if (thr == 0ull) return -1;                                // fewer than k live heads
  const int cw = s_cnt[warp];
  for (int i = lane; i < cw; i += 32) {                      // survivors are a prefix of each run
    const u64 key = lst[(size_t)warp * stride + i];
    if (key < thr) break;
    const int p = atomicAdd(s_ctr, 1);
    if (p < nthr) scratch[p] = key;
  }
  __syncthreads();
  const int S = *s_ctr;
  if (S > nthr) return -1;
  const u64 key = tid < S ? scratch[tid] : 0ull;
  int rank = 0;
  if (tid < S)
    for (int j = 0; j < S; ++j) rank += scratch[j] > key;
  __syncthreads();
  if (deal_stride > 0) {                                      // deal the selected keys out to W buffers: key of rank r -> buffer r % W, slot r / W
    if (tid < S && rank < k) out[(size_t)(rank % W) * deal_stride + rank / W] = key;
  } else {
    if (tid < S && rank < k) out[rank] = key;
    for (int i = S + tid; i < k; i += nthr) out[i] = 0ull;
  }
  __syncthreads();
  return S < k ? S : k;
}

// CPL = 16-byte chunks per lane (row = 8 lanes x CPL chunks), NQ queries per pass.
template <int CPL, int NQ, bool BF16, bool COSINE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_kernel(const void* __restrict__ index, const float* __restrict__ queries, int64_t N, int H,
                 int k, int cap, int q0, u64* __restrict__ cand, int slot_id) {
  constexpr bool DYN = CPL >= 8;                           // global round counter (see g_scan_rounds)
  typedef typename ChunkT<BF16>::type Chunk;
  constexpr int EPC = BF16 ? 8 : 4;                        // elements per chunk
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* sm = reinterpret_cast<u64*>(smem_raw);              // [NQ][warps][cap]
  pdl_trigger();                                           // the merge kernel may be set up while this one streams the index
  const bool dbg = g_scan_dbg_on && blockIdx.x == 0 && threadIdx.x == 0;
  TT_SCAN_STAMP(0);
  // dynamic rounds: local block b of `warps` steps <-> one global round; its id travels through a ring tagged with b
  __shared__ int s_next;
  __shared__ unsigned long long s_blk[16];
  unsigned* gctr = g_scan_rounds[slot_id];
  if (DYN) {
    if (threadIdx.x < 16) s_blk[threadIdx.x] = ~0ull;
    __syncthreads();
    if (threadIdx.x == 0) {                                 // the rounds of local blocks 0 and 1: their latency hides under the sample phase
      s_next = 0;
      const unsigned g0 = atomicAdd(gctr, 1u);
      if (g0 == 0u) gctr[2] = 1u;
      s_blk[0] = (unsigned long long)g0;
      s_blk[1] = (1ull << 32) | atomicAdd(gctr, 1u);
    }
  }

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
  // sample area behind the candidate buffers: [NQ][warps * kSampleRows] keys
  const int nsamp = kScanWarps * kSampleRows;
  u64* samp = sm + (size_t)NQ * kScanWarps * cap;

  const Chunk* base = reinterpret_cast<const Chunk*>(index);
  const int64_t quads = (N + 3) >> 2;
  const int64_t gw = (int64_t)blockIdx.x * kScanWarps + warp;
  const int64_t nw = (int64_t)gridDim.x * kScanWarps;

  // one step = NQD row quads of this warp (rows (it + j nw) * 4 .., j < NQD): 16 x 16-byte loads in flight per lane whatever
  // the row width (2 quads of 8 chunks for fp32 x 256 columns, 4 quads of 4 chunks for bf16 x 256 -- with two quads the
  // narrower rows had half the bytes in flight and ran latency-bound).  SAMPLE steps store every score into the sample area
  // instead of the warp's candidate buffer (slot = step * 4 NQD + quad * 4 + row-in-quad).
  constexpr int NQD = (CPL >= 8 || NQ > 1 || (BF16 && COSINE)) ? 2 : 4;   // four quads only where they fit in registers without spills
  constexpr int kSampleSteps = kSampleRows / (4 * NQD);
  auto step = [&](int64_t it, auto sample_tag, int sidx) {
    constexpr bool SAMPLE = decltype(sample_tag)::value;
    int64_t row[NQD];
    bool ok[NQD];
    Chunk a[NQD][CPL];
#pragma unroll
    for (int j = 0; j < NQD; ++j) {
      row[j] = (it + j * nw) * 4 + grp;
      ok[j] = (it + j * nw < quads) && row[j] < N;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        if (ok[j]) a[j][c] = ld_chunk(base + row[j] * chunks_per_row + sub + 8 * c); else zero_chunk(a[j][c]);
      }
    }
    float inv[NQD];
#pragma unroll
    for (int j = 0; j < NQD; ++j) {
      inv[j] = 1.0f;
      if (COSINE) {
        float sa = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) sa = sq_chunk(a[j][c], sa);
        inv[j] = 1.0f / fmaxf(sqrtf(group8_sum(sa)), 1e-8f);
      }
    }
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
#pragma unroll
      for (int j = 0; j < NQD; ++j) {
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) d = dot_chunk(a[j][c], qf[n][c], d);
        d = group8_sum(d);
        if (COSINE) d = d * (qinv[n] * inv[j]);
        if (SAMPLE) {
          if (sub == 0)
            samp[(size_t)n * nsamp + warp * kSampleRows + sidx * (4 * NQD) + j * 4 + grp] = ok[j] ? make_key(d, (uint32_t)row[j]) : 0ull;
        } else {
          tk[n].push(ok[j] && sub == 0 && d >= tk[n].thr, d, (uint32_t)row[j], lane);
        }
      }
    }
  };

  // Sample phase.  A warp sees only rows / (148 * warps) rows of the index, so a threshold built from its OWN k best rows is
  // weak on a small shard (1.25 M rows: 528 rows per warp, 19 % of them pass, and every warp sorts its buffer two or three
  // times at the same moment -- measured 50-60 us on top of the 187 us the bytes take).  Instead every warp scores its first
  // 64 rows unconditionally, the CTA sorts those warps * 64 sample keys once, and the score of the k-th best becomes the
  // starting threshold of every warp: at least k rows of this CTA reach it, so no row below it can be in the CTA's top-k.
  // The sample's k best keys continue as warp 0's candidates.  (k > warps * 64: no threshold, the scan works as before.)
  int64_t it = gw;
#pragma unroll 1
  for (int i = 0; i < kSampleSteps; ++i, it += NQD * nw) step(it, std::true_type{}, i);
  TT_SCAN_STAMP(1);
  __shared__ int s_cnt[kScanThreads / 32 + 1];
  __shared__ u64 s_scratch[kScanThreads];
  __shared__ int s_ctr;
  __shared__ u64 s_thr;
  if (!DYN && threadIdx.x == 0) s_next = kSampleSteps * kScanWarps;   // the sample rounds were taken statically
  __syncwarp();
#pragma unroll
  for (int n = 0; n < NQ; ++n) warp_sort_buffer<kSampleRows / 32>(samp + (size_t)n * nsamp + warp * kSampleRows, kSampleRows, lane);
  if (lane == 0) s_cnt[warp] = kSampleRows;
  __syncthreads();
#pragma unroll
  for (int n = 0; n < NQ; ++n) {
    u64* sp = samp + (size_t)n * nsamp;
    u64* buf0 = sm + (size_t)n * kScanWarps * cap;           // warp 0's candidate buffer
    float thr0 = -CUDART_INF_F;
    // the sample's k best keys continue as candidates, dealt out over the warps' buffers: every warp then ends the scan with
    // a few dozen keys, which its register sort (64 keys) handles -- all of them in warp 0's buffer made that warp sort 256
    int cnt0 = cta_select_runs(sp, kSampleRows, s_cnt, kScanWarps, k, buf0, s_scratch, &s_ctr, &s_thr, cap);
    if (cnt0 >= 0) {
      if (cnt0 == k) {                                       // the k-th best sample key: rank k - 1
        const u64 kth = buf0[(size_t)((k - 1) % kScanWarps) * cap + (k - 1) / kScanWarps];
        thr0 = bits_score((uint32_t)(kth >> 32));
      }
      tk[n].count = cnt0 > warp ? (cnt0 - warp + kScanWarps - 1) / kScanWarps : 0;
      tk[n].thr = thr0;
      continue;
    } else {                                                 // odd shapes (k > warps * 64, tiny shards): sort the whole sample
      bitonic_desc<true>(sp, nsamp, threadIdx.x, blockDim.x);
      const u64 kth = k <= nsamp ? sp[k - 1] : 0ull;
      if (kth != 0ull) thr0 = bits_score((uint32_t)(kth >> 32));
      const int live = k <= nsamp ? k : nsamp;               // zero keys (rows past the end) sort last and are never live
      int c = 0;
      if (warp == 0)
        for (int i = lane; i < live; i += 32) { const u64 key = sp[i]; buf0[i] = key; c += key != 0ull; }
      cnt0 = warp_sum_int(c);
    }
    if (warp == 0) tk[n].count = cnt0;
    tk[n].thr = thr0;
  }
  __syncwarp();
  TT_SCAN_STAMP(2);

  // Main loop.  A warp takes local step t from a shared-memory counter (warps of one SM do not progress at the same rate):
  // local block t / warps, warp slot t % warps.  Static variants: block b is round b of this CTA's interleaved share.  DYN:
  // the warp that takes slot 0 of block b fetches the global round of block b + 2 (one global atomic per warps * 8 rows,
  // issued before the warp's own step and published after it, so its ~1 us round trip is never waited for); everybody reads
  // the round of its own block from the ring.  Global round G = (CTA slot G % grid, round G / grid behind the sample
  // rounds) of the static interleaved schedule, so the memory access pattern does not change.
  {
    const int wshift = __ffs(kScanWarps) - 1;
    for (;;) {
      unsigned long long v = 0;
      unsigned g_next = 0, b_pub = 0;
      bool publish = false;
      if (lane == 0) {
        const int t = atomicAdd(&s_next, 1);
        const unsigned b = (unsigned)(t >> wshift);
        if (DYN) {
          publish = (t & (kScanWarps - 1)) == 0;
          if (publish) { g_next = atomicAdd(gctr, 1u); b_pub = b + 2; }
          do { v = reinterpret_cast<volatile unsigned long long*>(s_blk)[b & 15]; } while ((unsigned)(v >> 32) != b);
          v &= 0xffffffffull;
        } else {
          v = b;
        }
        v |= (unsigned long long)(unsigned)(t & (kScanWarps - 1)) << 32;
      }
      v = __shfl_sync(0xffffffffu, v, 0);
      const unsigned G = (unsigned)(v & 0xffffffffull);
      const int slot = (int)(v >> 32);
      const int64_t it0 = DYN ? (int64_t)(G % gridDim.x) * kScanWarps + (int64_t)(G / gridDim.x + kSampleSteps) * NQD * nw
                              : (int64_t)blockIdx.x * kScanWarps + (int64_t)G * NQD * nw;
      const bool done = it0 >= quads;
      if (!done) step(it0 + slot, std::false_type{}, 0);
      if (DYN && publish) reinterpret_cast<volatile unsigned long long*>(s_blk)[b_pub & 15] = ((unsigned long long)b_pub << 32) | g_next;
      if (done) break;
    }
  }
  TT_SCAN_STAMP(3);

  // block epilogue: every warp sorts its own candidates in registers (usually <= 64 keys), the CTA selects the top-k of the
  // sorted runs by head threshold + ranking (cta_select_runs); the packed bitonic sort remains for the shapes that does not fit
#pragma unroll
  for (int n = 0; n < NQ; ++n) {
    u64* region = sm + (size_t)n * kScanWarps * cap;
    const bool small = cap <= 32 * kMaxCapPerLane;          // candidate buffers of <= 256 keys travel through registers
    if (small) {
      const int c = tk[n].count;
      __syncwarp();
      // <= 64 keys (the usual case): the register sort the sample phase has already pulled into the instruction cache;
      // more: the in-warp shared-memory sort (rolled loops -- a fully unrolled 256-key register sort is ~2000 instructions
      // of cold code, ~10 us for the one warp that runs it)
      if (c <= 64) warp_sort_buffer<2>(tk[n].buf, c, lane);
      else tk[n].prune(lane);
    } else {
      tk[n].prune(lane);
    }
    __syncwarp();
    const int cnt = tk[n].count;
    __syncthreads();                                         // the previous query's selection has finished with s_cnt
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    const int sel = small ? cta_select_runs(region, cap, s_cnt, kScanWarps, k, region, s_scratch, &s_ctr, &s_thr) : -1;
    if (sel < 0) {
      int off = 0, total = 0;
      for (int w = 0; w < kScanWarps; ++w) { const int c = s_cnt[w]; off += w < warp ? c : 0; total += c; }
      int n2 = 32;
      while (n2 < total || n2 < k) n2 <<= 1;
      if (small) {                                           // pack the live keys of all warps (through registers), sort them
        u64 keep[kMaxCapPerLane];
#pragma unroll
        for (int u = 0; u < kMaxCapPerLane; ++u) keep[u] = (u * 32 + lane < cnt) ? tk[n].buf[u * 32 + lane] : 0ull;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kMaxCapPerLane; ++u)
          if (u * 32 + lane < cnt) region[off + u * 32 + lane] = keep[u];
        for (int i = total + threadIdx.x; i < n2; i += blockDim.x) region[i] = 0ull;
        __syncthreads();
        bitonic_desc<true>(region, n2, threadIdx.x, blockDim.x);
      } else {
        for (int i = cnt + lane; i < cap; i += 32) tk[n].buf[i] = 0ull;
        __syncthreads();
        bitonic_desc<true>(region, kScanWarps * cap, threadIdx.x, blockDim.x);
      }
    }
    TT_SCAN_STAMP(4);
    u64* out = cand + ((size_t)(q0 + n) * gridDim.x + blockIdx.x) * k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = region[i];
    __syncthreads();
  }
  TT_SCAN_STAMP(5);
  if (DYN && threadIdx.x == 0) {                            // last CTA out re-arms the counters for the next launch on this slot
    __threadfence();
    if (atomicAdd(gctr + 1, 1u) == gridDim.x - 1) {
      if (*reinterpret_cast<volatile unsigned*>(gctr + 2) != 1u) {
        printf("tt_b200: scan round counter of slot %d was not zero when the launch began\n", slot_id);
        __trap();
      }
      gctr[0] = 0u; gctr[1] = 0u; gctr[2] = 0u;
      __threadfence();
    }
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

struct MergeShared {
  unsigned int hist[256];
  u64 sel[TT_TOPK_MAX];
  u64 prefix, mask;
  int need, count, all;
};

// the whole CTA (kMergeThreads threads): leaves the k largest of the L keys src.get(q, 0..L) in sh.sel[0..k), descending
// (zero-padded when fewer than k keys are live); ends with a CTA-wide barrier
template <typename Keys>
__device__ __forceinline__ void select_topk_sorted(const Keys& src, int q, int64_t L, int k, MergeShared& sh) {
  const int tid = threadIdx.x;
  const bool in_regs = L <= (int64_t)kMergeRegKeys * kMergeThreads;
  u64 mine[kMergeRegKeys];
  if (in_regs) {
#pragma unroll
    for (int u = 0; u < kMergeRegKeys; ++u) {
      const int64_t i = (int64_t)u * kMergeThreads + tid;
      mine[u] = i < L ? src.get(q, i) : 0ull;               // 0 never matches a live prefix: see below
    }
  }
  if (tid == 0) { sh.prefix = 0ull; sh.mask = 0ull; sh.need = k; sh.count = 0; sh.all = 0; }
  for (int i = tid; i < TT_TOPK_MAX; i += kMergeThreads) sh.sel[i] = 0ull;
  __syncthreads();
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += kMergeThreads) sh.hist[i] = 0u;
    __syncthreads();
    const u64 prefix = sh.prefix, mask = sh.mask;
    if (in_regs) {
#pragma unroll
      for (int u = 0; u < kMergeRegKeys; ++u) {
        const u64 key = mine[u];
        if (key != 0ull && (key & mask) == prefix) atomicAdd(&sh.hist[(unsigned)((key >> shift) & 0xffull)], 1u);
      }
    } else {
      for (int64_t i = tid; i < L; i += kMergeThreads) {
        const u64 key = src.get(q, i);
        if (key != 0ull && (key & mask) == prefix) atomicAdd(&sh.hist[(unsigned)((key >> shift) & 0xffull)], 1u);
      }
    }
    __syncthreads();
    if (tid < 32) {
      // lane l owns bins [8l, 8l+8).  Walking from bin 255 down, find the bin in which the running count reaches `need`.
      unsigned c[8], lane_sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = sh.hist[8 * tid + j]; lane_sum += c[j]; }
      unsigned above = lane_sum;                               // inclusive suffix sum over lanes tid..31
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_down_sync(0xffffffffu, above, o);
        if (tid + o < 32) above += t;
      }
      const int need = sh.need;
      const unsigned excl = above - lane_sum;                   // keys in bins above this lane's
      const bool here = excl < (unsigned)need && above >= (unsigned)need;
      const unsigned who = __ballot_sync(0xffffffffu, here);
      if (who == 0u) {                                          // fewer than k live keys in total (first pass only):
        if (tid == 0) { sh.all = 1; sh.prefix = 0ull; }         // every live key is selected, the rest is padding
      } else if (here) {
        int rem = need - (int)excl, bin = 7;
        for (; bin > 0; --bin) {
          if ((int)c[bin] >= rem) break;
          rem -= (int)c[bin];
        }
        sh.need = rem;
        sh.prefix = prefix | ((u64)(8 * tid + bin) << shift);
        sh.mask = mask | (0xffull << shift);
        // the chosen bin holds exactly the keys still needed: every key >= prefix (lower digits zero) is selected and the
        // remaining digit passes cannot change that -- stop (typically after the 3rd of 8 passes)
        if ((int)c[bin] == rem) sh.all = 1;
      }
    }
    __syncthreads();
    if (sh.all) break;                                       // CTA-uniform
  }
  const u64 kth = sh.prefix;                               // exact k-th largest key (keys are unique); 0 = take all
  if (in_regs) {
#pragma unroll
    for (int u = 0; u < kMergeRegKeys; ++u) {
      const u64 key = mine[u];
      if (key >= kth && key != 0ull) {
        const int p = atomicAdd(&sh.count, 1);
        if (p < TT_TOPK_MAX) sh.sel[p] = key;
      }
    }
  } else {
    for (int64_t i = tid; i < L; i += kMergeThreads) {
      const u64 key = src.get(q, i);
      if (key >= kth && key != 0ull) {
        const int p = atomicAdd(&sh.count, 1);
        if (p < TT_TOPK_MAX) sh.sel[p] = key;
      }
    }
  }
  __syncthreads();
  int n2 = 32;
  while (n2 < k) n2 <<= 1;
  bitonic_desc<true>(sh.sel, n2, tid, kMergeThreads);
}

// Fast path for G DESCENDING lists of k keys (the scan's per-block lists, the shards' candidate lists): the first
// r = ceil(k / G) keys of every list are G r >= k distinct keys, so the k-th largest of these heads is a lower bound T of the
// k-th largest key overall; only keys >= T can be selected.  With 148 block lists of 100 that leaves a few hundred of the
// 14 800 keys: one 256-key sort for T, one pass over the keys (already in registers), one sort of the survivors -- instead of
// three or four 8-bit radix passes over everything.  Returns false (nothing decided, sh in an arbitrary state) when the shape
// or the data do not fit: fewer than k live heads, more than TT_TOPK_MAX survivors; the caller then runs the radix select.
template <typename Lists>
__device__ __forceinline__ bool select_from_lists(const Lists& src, int q, int G, int k, MergeShared& sh) {
  const int tid = threadIdx.x;
  const int r = (k + G - 1) / G;
  const int nh = G * r;
  const int64_t L = (int64_t)G * k;
  if (nh > TT_TOPK_MAX || L > (int64_t)kMergeRegKeys * kMergeThreads) return false;   // CTA-uniform
  u64 mine[kMergeRegKeys];
  {
    // key i = u * 1024 + tid lives at (list i / k, position i % k): one division, then strides of 1024 keys
    int g = tid / k, pos = tid - g * k;
    const int dg = kMergeThreads / k, dp = kMergeThreads - dg * k;
    const int iL = (int)L;
#pragma unroll
    for (int u = 0; u < kMergeRegKeys; ++u) {
      mine[u] = u * kMergeThreads + tid < iL ? src.at(q, g, pos) : 0ull;
      g += dg; pos += dp;
      if (pos >= k) { pos -= k; ++g; }
    }
  }
  // Both selections are RANK computations over a few hundred unique keys held one per thread (every thread counts the keys
  // larger than its own: broadcast reads of shared memory, ONE barrier) -- a 256-key bitonic sort by 32 warps spends 36
  // CTA-wide barriers, ~5 us, for the same answer.
  const u64 head = tid < nh ? src.at(q, tid / r, tid - (tid / r) * r) : 0ull;
  if (tid < nh) sh.sel[tid] = head;
  if (tid == 0) { sh.count = 0; sh.prefix = 0ull; }
  __syncthreads();
  if (tid < nh && head != 0ull) {
    int rank = 0;
    for (int j = 0; j < nh; ++j) rank += sh.sel[j] > head;
    if (rank == k - 1) sh.prefix = head;                     // the k-th largest head (keys are unique)
  }
  __syncthreads();
  const u64 thr = sh.prefix;
  if (thr == 0ull) return false;                             // fewer than k live heads (tiny shards); CTA-uniform
#pragma unroll
  for (int u = 0; u < kMergeRegKeys; ++u) {
    const u64 key = mine[u];
    if (key >= thr) {
      const int p = atomicAdd(&sh.count, 1);
      if (p < TT_TOPK_MAX) sh.sel[p] = key;
    }
  }
  __syncthreads();
  const int S = sh.count;                                    // >= k: the heads themselves
  if (S > kMergeThreads || S > TT_TOPK_MAX) { __syncthreads(); return false; }
  const u64 key = tid < S ? sh.sel[tid] : 0ull;
  int rank = 0;
  if (tid < S)
    for (int j = 0; j < S; ++j) rank += sh.sel[j] > key;
  __syncthreads();                                           // every survivor is in a register: permute in place
  if (tid < S) sh.sel[rank] = key;
  __syncthreads();
  return true;
}
struct RawLists {               // RawKeys seen as per_query / k sorted lists of k keys
  const u64* keys; int64_t per_query; int k;
  __device__ __forceinline__ u64 at(int q, int g, int pos) const { return keys[(int64_t)q * per_query + (int64_t)g * k + pos]; }
};
template <typename Keys>
__device__ __forceinline__ bool try_sorted_lists(const Keys&, int, int64_t, int, MergeShared&) { return false; }
__device__ __forceinline__ bool try_sorted_lists(const RawKeys& src, int q, int64_t L, int k, MergeShared& sh) {
  if (L % k != 0 || L / k > 0x7fffffff) return false;
  return select_from_lists(RawLists{src.keys, src.per_query, k}, q, (int)(L / k), k, sh);
}

__device__ __forceinline__ void write_topk(const MergeShared& sh, int q, int k, int64_t id_offset, float* __restrict__ out_scores,
                                           int64_t* __restrict__ out_ids) {
  for (int i = threadIdx.x; i < k; i += kMergeThreads) {
    const u64 key = sh.sel[i];
    if (key == 0ull) { out_scores[(int64_t)q * k + i] = -CUDART_INF_F; out_ids[(int64_t)q * k + i] = -1; }
    else {
      out_scores[(int64_t)q * k + i] = bits_score((uint32_t)(key >> 32));
      out_ids[(int64_t)q * k + i] = (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) + id_offset;
    }
  }
}

template <typename Keys>
__global__ void __launch_bounds__(kMergeThreads)
merge_topk_kernel(Keys src, int64_t L, int k, int64_t id_offset, float* __restrict__ out_scores,
                  int64_t* __restrict__ out_ids) {
  __shared__ MergeShared sh;
  pdl_wait();                                                // launched programmatically behind the scan (tt_topk_scan); a no-op otherwise
  if (!try_sorted_lists(src, blockIdx.x, L, k, sh)) select_topk_sorted(src, blockIdx.x, L, k, sh);
  write_topk(sh, blockIdx.x, k, id_offset, out_scores, out_ids);
}

// ---- row-sharded search: block merge + candidate exchange over NVLink peer memory + final merge in ONE launch ---------
// (local top-k per shard merged by all-gather).  One CTA per query:
//   1. radix-select this shard's k best keys out of the scan's per-block lists (as merge_topk_kernel does) and rebase their
//      row field to GLOBAL row numbers (id_offset + N_total < 2^32), so keys of different shards compare directly:
//      higher score first, ties -> lower global row (BASELINE tie rule);
//   2. store the k keys into slot `rank` of EVERY rank's exchange buffer (8-byte stores over NVLink), fence.sys, and bump
//      every rank's arrival counter for this source -- the protocol of p2p_allgather_kernel (p2p.cu): counters only grow,
//      round r expects r * ctas arrivals per source, rounds alternate between the two slot sets;
//   3. wait until every source's counter in the LOCAL header has reached the round's target (bounded: a peer that never
//      arrives is reported through header word 34, tt_p2p_status);
//   4. radix-select the k best of the world * k gathered keys (read through L2: the peers' stores bypass this SM's L1) and
//      write scores / global ids.
// Replaces three launches of the plain chain (block merge, tt_p2p_allgather, tt_topk_merge) and their two boundaries.
struct P2PSearchArgs {
  int world, rank, halves;
  unsigned timeout_s;
  unsigned char* base[8];
  size_t slot_bytes;
};
struct GatheredKeys {           // world slots of [nq][kpad] keys, every slot's k keys descending
  const u64* keys; int64_t slot_keys; int k, kpad;
  __device__ __forceinline__ u64 get(int q, int64_t i) const { return __ldcg(keys + (i / k) * slot_keys + (int64_t)q * kpad + (i % k)); }
  __device__ __forceinline__ u64 at(int q, int g, int pos) const { return __ldcg(keys + (int64_t)g * slot_keys + (int64_t)q * kpad + pos); }
};

__global__ void __launch_bounds__(kMergeThreads)
merge_exchange_merge_kernel(const u64* __restrict__ cand, int64_t per_query, int k, int kpad, unsigned id_offset, const P2PSearchArgs a,
                            float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  __shared__ MergeShared sh;
  __shared__ unsigned s_round;
  const int q = blockIdx.x, tid = threadIdx.x;
  unsigned* hdr = reinterpret_cast<unsigned*>(a.base[a.rank]);
  pdl_wait();                                                // the scan kernel has retired (this kernel is launched programmatically behind it)
  if (tid == 0) {
    atomicAdd(hdr + 36, 1u);                                         // CTAs started (see p2p_allgather_kernel)
    s_round = *reinterpret_cast<volatile unsigned*>(hdr + 32) + 1u;  // every CTA reads it before the last one out bumps it
  }
  RawKeys src{cand, per_query};
  if (!try_sorted_lists(src, q, per_query, k, sh)) select_topk_sorted(src, q, per_query, k, sh);
  const unsigned round = s_round;
  const size_t half_off = (a.halves == 2 && (round & 1u)) ? (size_t)a.world * a.slot_bytes : 0;
  const size_t slot0 = 256 + half_off;                               // kP2PHeaderBytes
  if (tid < kpad) {
    u64 key = tid < k ? sh.sel[tid] : 0ull;
    if (key != 0ull) key -= (u64)id_offset;                          // row field = ~row: ~(row + off) = ~row - off (no borrow: row + off < 2^32)
    for (int p = 0; p < a.world; ++p)
      reinterpret_cast<u64*>(a.base[p] + slot0 + (size_t)a.rank * a.slot_bytes)[(size_t)q * kpad + tid] = key;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < a.world) {
    unsigned* peer = reinterpret_cast<unsigned*>(a.base[tid]) + a.rank;
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(peer), "r"(1u) : "memory");
    const unsigned target = round * gridDim.x;                       // the exchange's CTA count == this grid (checked on the host)
    const unsigned* c = hdr + tid;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
      if (v >= target) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > (unsigned long long)a.timeout_s * 1000000000ull) {   // give up, tell the host, keep the context alive
        hdr[34] = 1u + tid;
        __threadfence_system();
        break;
      }
    }
  }
  __syncthreads();
  GatheredKeys g{reinterpret_cast<const u64*>(a.base[a.rank] + slot0), (int64_t)(a.slot_bytes / 8), k, kpad};
  if (!select_from_lists(g, q, a.world, k, sh)) select_topk_sorted(g, q, (int64_t)a.world * k, k, sh);
  write_topk(sh, q, k, 0, out_scores, out_ids);
  if (tid == 0) {
    __threadfence();
    const unsigned t = atomicAdd(hdr + 33, 1u);
    if (t == gridDim.x - 1) { hdr[33] = 0u; hdr[32] = round; }
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
  p.smem_per_q = (size_t)warps * (cap + kSampleRows) * sizeof(u64);     // candidate buffers + the sample area
  p.cand_bytes = align_up((size_t)nq * p.grid * k * sizeof(u64));
  p.total = p.cand_bytes + 256;
  return p;
}

static int next_scan_slot() {
  static std::atomic<unsigned> n{0};
  return (int)(n.fetch_add(1u, std::memory_order_relaxed) % (unsigned)kScanSlots);
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
      scan_topk_kernel<CPL, 2, BF16, COS><<<plan.grid, plan.warps * 32, 2 * plan.smem_per_q, s>>>(index, queries, N, H, k, plan.cap, q, cand, next_scan_slot());
      TT_LAUNCH_CHECK("scan_topk_kernel");
    }
  }
  if (q < nq) {
    TT_CUDA(cudaFuncSetAttribute(scan_topk_kernel<CPL, 1, BF16, COS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_per_q));
    for (; q < nq; ++q) {
      scan_topk_kernel<CPL, 1, BF16, COS><<<plan.grid, plan.warps * 32, plan.smem_per_q, s>>>(index, queries, N, H, k, plan.cap, q, cand, next_scan_slot());
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
  TT_CUDA(tt::launch_kernel(tt::merge_topk_kernel<tt::RawKeys>, dim3((unsigned)nq), dim3(tt::kMergeThreads), 0, s, true, src,
                            (int64_t)plan.grid * k, k, id_offset, out_scores, out_ids));
  TT_LAUNCH_CHECK("merge_topk_kernel");
  static const bool dbg_on = getenv("TT_SCAN_DEBUG") != nullptr;
  if (dbg_on) {
    static int calls = 0;
    const int one = 1;
    cudaStreamSynchronize(s);
    if (calls++ == 0) cudaMemcpyToSymbol(tt::g_scan_dbg_on, &one, sizeof(int));
    else {
      long long t[8];
      cudaMemcpyFromSymbol(t, tt::g_scan_dbg, sizeof(t));
      printf("[tt scan CTA 0, ns] N=%lld bf16=%d: sample scored %lld | threshold known %lld | loop done %lld | packed+sorted %lld | end %lld\n",
             (long long)N, index_bf16, t[1] - t[0], t[2] - t[0], t[3] - t[0], t[4] - t[0], t[5] - t[0]);
    }
  }
  return TT_OK;
}

int tt_topk_scan_p2p_ok(const tt_p2p_t* x, int nq, int k) {
  if (!x || nq <= 0 || k <= 0 || k > TT_TOPK_MAX) return 0;
  const int kpad = (k + 1) & ~1;
  return (x->world >= 1 && x->world <= 8 && x->double_buffered && x->ctas == nq && (size_t)nq * kpad * 8 <= x->slot_bytes &&
          (int64_t)x->world * k <= (int64_t)tt::kMergeRegKeys * tt::kMergeThreads) ? 1 : 0;
}

int tt_topk_scan_p2p(const void* index, int index_bf16, const float* queries, int64_t N, int H, int nq, int k,
                     int cosine, int64_t id_offset, const tt_p2p_t* x, float* out_scores, int64_t* out_ids, void* workspace,
                     size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(index && queries && out_scores && out_ids && x && N > 0 && H > 0 && nq > 0, "topk_scan_p2p: bad arguments");
  TT_CHECK_ARG(k > 0 && k <= TT_TOPK_MAX && k <= N, "topk_scan_p2p: need 0 < k <= min(N, %d) (k=%d, N=%lld)", TT_TOPK_MAX, k, (long long)N);
  TT_CHECK_ARG(id_offset >= 0 && id_offset + N < (1ll << 32), "topk_scan_p2p: global row numbers must stay below 2^32");
  if (!tt_topk_scan_p2p_ok(x, nq, k)) {
    tt::set_error("topk_scan_p2p: needs a double-buffered exchange created with ctas == nq and slots of >= nq * k * 8 bytes");
    return TT_ERR_UNSUPPORTED;
  }
  TT_CHECK_ARG(x->rank >= 0 && x->rank < x->world, "topk_scan_p2p: bad rank");
  const tt::ScanPlan plan = tt::plan_scan(N, nq, k);
  if (workspace == nullptr || workspace_bytes < plan.total) { tt::set_error("topk_scan_p2p: workspace too small (%zu < %zu)", workspace_bytes, plan.total); return TT_ERR_WORKSPACE; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  tt::u64* cand = reinterpret_cast<tt::u64*>(workspace);
  int rc;
  if (index_bf16) rc = cosine ? tt::dispatch_scan<true, true>(index, queries, N, H, nq, k, plan, cand, s)
                              : tt::dispatch_scan<true, false>(index, queries, N, H, nq, k, plan, cand, s);
  else            rc = cosine ? tt::dispatch_scan<false, true>(index, queries, N, H, nq, k, plan, cand, s)
                              : tt::dispatch_scan<false, false>(index, queries, N, H, nq, k, plan, cand, s);
  if (rc) return rc;
  tt::P2PSearchArgs a{};
  a.world = x->world; a.rank = x->rank; a.halves = 2; a.slot_bytes = x->slot_bytes;
  a.timeout_s = x->timeout_s > 0 ? (unsigned)x->timeout_s : 600u;
  for (int p = 0; p < x->world; ++p) {
    TT_CHECK_ARG(x->base[p] != nullptr, "topk_scan_p2p: peer %d not mapped", p);
    a.base[p] = static_cast<unsigned char*>(x->base[p]);
  }
  TT_CUDA(tt::launch_kernel(tt::merge_exchange_merge_kernel, dim3((unsigned)nq), dim3(tt::kMergeThreads), 0, s, true, (const tt::u64*)cand,
                            (int64_t)plan.grid * k, k, (k + 1) & ~1, (unsigned)id_offset, a, out_scores, out_ids));
  TT_LAUNCH_CHECK("merge_exchange_merge_kernel");
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
