// tc_topk.cu -- K7 batched: brute-force scan of a bf16 document index for up to 128 queries per pass on the
// tcgen05 tensor cores, fused with an exact per-query top-k.
//
// Reference: TwoTowerSearch.search scores ONE query per call (inference/search/two_tower.py:98-105); a batch of nq
// queries through that API reads the index nq times.  Here the index is read ONCE per 128 queries:
//   S[q, n] = Q[q, :] . D[n, :]     as  tcgen05.mma  M = 128 (queries), N = 128 (documents per tile), K = H
// with the query tile resident in tensor memory (TS-mode A operand) and 128-row document tiles streamed through a
// 3-stage TMA ring (64 KB per tile at H = 256).  HBM-bound until nq ~ peak_flops * 2 / (2 * peak_bytes).
//
// Exactness.  Queries are fp32 in the API; a bf16 tensor-core product would round them to 8 mantissa bits and move
// near-boundary documents in or out of the top-k.  Each query is therefore split into bf16 hi + bf16 lo
// (q = hi + lo + O(2^-17 |q|)) and both halves are multiplied with the same document tile into the same fp32
// accumulator (K = 2H): scores agree with the fp32-query scan kernel (topk_scan.cu) to ~1e-5 relative.
//
// Top-k.  An epilogue thread owns one TMEM lane = one query.  It keeps a running threshold (score of its current k-th
// best) in a register and appends (score, row) keys of documents at or above it to a per-(query, CTA) candidate buffer
// in global memory (L2-resident: CAP x 8 B); a 32-column chunk whose maximum is below the threshold costs one compare.
// A full buffer is pruned to its k best by the whole warp (bitonic sort in shared memory, as in topk_scan.cu), which
// also tightens the threshold.  Expected appends per query and CTA ~ k (1 + ln(rows_per_cta / k)); nothing of size N
// is written.  The second kernel (merge_topk_kernel, topk_scan.cu) selects the exact top-k over all CTAs' buffers.
// Keys are (order-preserving score bits << 32) | ~row: ties -> lower row index, identical to the single-query path.
#include <math_constants.h>

#include "tc_common.cuh"

namespace tt {

typedef unsigned long long u64;

// shared with topk_scan.cu
int topk_merge_blocked(const u64* cand, int blocks, int cap, int nq, int k, int64_t id_offset, float* out_scores, int64_t* out_ids,
                       cudaStream_t s);

namespace tc {

constexpr int TK_BM = 128;            // queries per pass (UMMA M)
constexpr int TK_BN = 128;            // documents per tile (UMMA N)
constexpr int TK_STAGES = 3;
constexpr int TK_EW = 8;              // epilogue warps: two per TMEM lane quarter, each owning a 64-column half of every tile
constexpr int TK_THREADS = 64 + TK_EW * 32;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int TK_CAP = 512;           // candidate keys per (query, CTA): a prune frees CAP - 32 - k slots, so the rows seen
                                      // between two prunes grow by (CAP - 32) / k (4.8x at k = 100): ~4 prunes per query and CTA
constexpr int TK_KMAX = 128;          // k <= TK_KMAX
constexpr int TK_KPL = TK_CAP / 32;   // keys per lane while a buffer is selected in registers

__device__ __forceinline__ uint32_t tk_score_bits(float f) {
  f += 0.0f;
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float tk_bits_score(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ u64 tk_make_key(float score, uint32_t row) {
  return ((u64)tk_score_bits(score) << 32) | (u64)(0xffffffffu - row);
}

// Warp-cooperative exact selection: keep the k largest of the cnt (k <= cnt <= TK_CAP) unique keys in buf[0, cnt), in any
// order, in buf[0, k); returns the k-th largest key.  The keys sit in registers (TK_KPL per lane); the k-th largest SCORE
// (high word) is found by an MSB-first bisection with one warp-wide integer reduction per bit, and only if more keys tie
// with it than fit is the low word (inverted row: larger = lower row) bisected as well.  ~0.5 us against ~6 us for a
// bitonic sort of the buffer in shared memory.
__device__ __forceinline__ u64 tk_select_topk(u64* buf, int cnt, int k, int lane) {
  u64 kv[TK_KPL];
#pragma unroll
  for (int j = 0; j < TK_KPL; ++j) { const int i = lane + 32 * j; kv[j] = i < cnt ? buf[i] : 0ull; }
  uint32_t hi = 0;                                         // largest value with count(score bits >= hi) >= k
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = hi | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < TK_KPL; ++j) c += ((uint32_t)(kv[j] >> 32) >= cand) ? 1 : 0;
    if (__reduce_add_sync(0xffffffffu, c) >= k) hi = cand;
  }
  int gt = 0, ge = 0;
#pragma unroll
  for (int j = 0; j < TK_KPL; ++j) { const uint32_t h = (uint32_t)(kv[j] >> 32); gt += h > hi ? 1 : 0; ge += h >= hi ? 1 : 0; }
  gt = __reduce_add_sync(0xffffffffu, gt);
  ge = __reduce_add_sync(0xffffffffu, ge);
  uint32_t lo = 0;                                         // ties at the k-th score: the k - gt largest low words among them
  if (ge > k) {
    const int need = k - gt;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = lo | (1u << bit);
      int c = 0;
#pragma unroll
      for (int j = 0; j < TK_KPL; ++j) c += ((uint32_t)(kv[j] >> 32) == hi && (uint32_t)kv[j] >= cand) ? 1 : 0;
      if (__reduce_add_sync(0xffffffffu, c) >= need) lo = cand;
    }
  }
  const u64 kth = ((u64)hi << 32) | (u64)lo;
  __syncwarp();                                            // every lane holds its keys in registers before buf is rewritten
  int base = 0;
#pragma unroll
  for (int j = 0; j < TK_KPL; ++j) {
    const bool keep = kv[j] >= kth;                        // zero padding: kth > 0 because cnt >= k real keys exist
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[base + __popc(m & ((1u << lane) - 1u))] = kv[j];
    base += __popc(m);
  }
  __syncwarp();
  return kth;
}

// Warp-cooperative prune of the candidate buffers of the lanes flagged in `mask` (warp-uniform): each flagged lane's buffer
// is cut to its k best and that lane's (count, threshold) are updated.  One out-of-line copy (see the epilogue).
// Values travel in and out by value ((threshold bits << 32) | count) so the caller's count / threshold stay in registers.
__device__ __noinline__ u64 tk_prune_lanes(unsigned mask, u64* gbuf, int count, float thr, int k, int lane) {
  while (mask) {
    const int L = __ffs(mask) - 1;
    mask &= mask - 1;
    const u64 gp = __shfl_sync(0xffffffffu, (u64)(uintptr_t)gbuf, L);
    const int cnt = __shfl_sync(0xffffffffu, count, L);
    __syncwarp();                                          // lane L's appends are visible to the warp
    const u64 kth = tk_select_topk(reinterpret_cast<u64*>((uintptr_t)gp), cnt, k, lane);
    if (lane == L) { count = k; thr = tk_bits_score((uint32_t)(kth >> 32)); }
  }
  return ((u64)__float_as_uint(thr) << 32) | (u64)(uint32_t)count;
}

// fp32 queries [nq, H] -> bf16 hi / lo tiles [TK_BM, H] each (rows >= nq are zero) and, for cosine scores, 1 / max(|q|, 1e-8)
__global__ void __launch_bounds__(256)
tk_split_queries_kernel(const float* __restrict__ q, int nq, int H, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                        float* __restrict__ qinv) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= TK_BM) return;
  float ss = 0.f;
  for (int e = lane; e < H; e += 32) {
    const float v = row < nq ? q[(int64_t)row * H + e] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16(v);
    hi[(int64_t)row * H + e] = h;
    lo[(int64_t)row * H + e] = __float2bfloat16(v - __bfloat162float(h));
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) qinv[row] = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
}

// 1 / max(|d_n|, 1e-8) per index row (cosine mode; computed once per index, cached by the caller)
__global__ void __launch_bounds__(256)
tk_row_inv_norm_kernel(const __nv_bfloat16* __restrict__ d, int64_t N, int H, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  float ss = 0.f;
  for (int e = lane; e < H; e += 32) { const float v = __bfloat162float(d[row * H + e]); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  if (lane == 0) out[row] = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
}

// Shared memory: 3 document stages x (128 x H x 2 B) (the query hi / lo tiles are staged through stages 1 and 2 into
// TMEM), barriers, per-warp prune scratch [4][TK_CAP] u64, per-tile document scales [2][128].
// Tensor memory: S buffers [0,128) [128,256) | Q hi [256, 256 + H/2) | Q lo [384, 384 + H/2).
__global__ void __launch_bounds__(TK_THREADS, 1)
tc_topk_kernel(const __grid_constant__ CUtensorMap tmQhi, const __grid_constant__ CUtensorMap tmQlo,
               const __grid_constant__ CUtensorMap tmD, int64_t N, int H, int k, int nq_pass, int tiles_per_cta,
               const float* __restrict__ qinv, const float* __restrict__ dinv, u64* __restrict__ cand) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const int kq = H / 64;
  const uint32_t d_bytes = (uint32_t)TK_BN * H * 2;
  uint8_t* d_tiles = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(d_tiles + TK_STAGES * d_bytes);
  uint64_t* q_bar = bars;                                  // both query tiles have landed in stages 1 and 2
  uint64_t* d_full = bars + 1;
  uint64_t* d_empty = d_full + TK_STAGES;
  uint64_t* s_full = d_empty + TK_STAGES;                  // [2]
  uint64_t* s_empty = s_full + 2;                          // [2]
  uint64_t* q_ready = s_empty + 2;                         // query tiles copied into TMEM (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  float* dscale = reinterpret_cast<float*>(tmem_slot + 4);                  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = ceil_div(N, TK_BN);
  const int64_t t_beg = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t t_end = t_beg + tiles_per_cta < ntiles ? t_beg + tiles_per_cta : ntiles;
  const int nt = (int)(t_end > t_beg ? t_end - t_beg : 0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQhi); tma_prefetch_desc(&tmQlo); tma_prefetch_desc(&tmD);
    mbar_init(q_bar, 1);
    for (int s = 0; s < TK_STAGES; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], TK_EW); }
    mbar_init(q_ready, TK_EW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;                      // columns [0,256): two S buffers of 128
  const uint32_t tmem_qh = tmem_s + 256;                   // [256, 256 + H/2): query hi tile (TMEM A operand)
  const uint32_t tmem_ql = tmem_s + 384;                   // [384, 384 + H/2): query lo tile

  if (warp == 0) {
    // query tiles borrow stages 1 (hi) and 2 (lo) until they sit in TMEM
    if (elect_one()) {
      mbar_arrive_expect_tx(q_bar, 2 * d_bytes);
      for (int kb = 0; kb < kq; ++kb) {
        tma_load_2d(d_tiles + 1 * d_bytes + kb * (TK_BM * 128), &tmQhi, q_bar, kb * 64, 0);
        tma_load_2d(d_tiles + 2 * d_bytes + kb * (TK_BM * 128), &tmQlo, q_bar, kb * 64, 0);
      }
    }
    __syncwarp();
    for (int i = 0; i < nt; ++i) {                         // whole warp, uniform control flow; one lane issues
      const int s = i % TK_STAGES;
      if (i == 1) mbar_wait(q_ready, 0);                   // stages 1 and 2 are free once the queries live in TMEM
      mbar_wait(&d_empty[s], ((i / TK_STAGES) & 1) ^ 1);
      uint8_t* dt = d_tiles + s * d_bytes;
      const int64_t row0 = (t_beg + i) * TK_BN;
      if (elect_one()) {
        mbar_arrive_expect_tx(&d_full[s], d_bytes);
        for (int kb = 0; kb < kq; ++kb) tma_load_2d(dt + kb * (TK_BN * 128), &tmD, &d_full[s], kb * 64, (int)row0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(TK_BM, TK_BN, 0, 0);
    const uint64_t dd0 = umma_desc_kmajor(smem_u32(d_tiles), 0);
    mbar_wait(q_ready, 0);
    tc_fence_after();
    const bool leader = elect_one();
    for (int i = 0; i < nt; ++i) {
      const int s = i % TK_STAGES, b = i & 1;
      mbar_wait(&d_full[s], (i / TK_STAGES) & 1);
      mbar_wait(&s_empty[b], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      if (leader) {
        const uint32_t td = tmem_s + b * TK_BN;
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {             // hi then lo: the same document tile, the same accumulator
          uint64_t dd = dd0 + (uint64_t)((s * d_bytes) >> 4);
          uint32_t ta = part ? tmem_ql : tmem_qh;
          for (int kb = 0; kb < kq; ++kb) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ts(td, ta + (uint32_t)(kk * 8), dd + (uint64_t)(kk * 2), idesc, (part | kb | kk) != 0);
            dd += TK_BN * 128 / 16;
            ta += 32;
          }
        }
        umma_commit(&d_empty[s]);
        umma_commit(&s_full[b]);
      }
      __syncwarp();
    }
  } else {
    // epilogue warps 2..9: TMEM lane quarter = warp & 3, column half = (warp - 2) / 4; a thread = one query x one 64-column
    // half of every tile, with its own candidate buffer, count and threshold (top-k of a union is inside the union of the
    // top-ks).  One warp per quarter walked the whole 128-column row serially and its dependent chain (TMEM load -> max
    // -> vote -> append) took longer per tile than the tile's HBM time; two warps per scheduler also hide each other.
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int lrow = quarter * 32 + lane;                  // query index inside the pass == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    // query tiles: smem (TMA, 128B swizzle) -> registers -> TMEM
    mbar_wait(q_bar, 0);
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
      const uint8_t* qt = d_tiles + (1 + part) * d_bytes;
      for (int kb = half; kb < kq; kb += 2) {
        uint32_t xr[32];
        const uint8_t* xrow = qt + kb * (TK_BM * 128) + lrow * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint4 v = *reinterpret_cast<const uint4*>(xrow + ((ch ^ (lrow & 7)) << 4));
          xr[4 * ch] = v.x; xr[4 * ch + 1] = v.y; xr[4 * ch + 2] = v.z; xr[4 * ch + 3] = v.w;
        }
        tmem_st_x32((part ? tmem_ql : tmem_qh) + lane_addr + (uint32_t)(kb * 32), xr);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(q_ready);

    const bool cosine = dinv != nullptr;
    const float my_qinv = cosine ? qinv[lrow] : 1.0f;
    u64* gbuf = cand + (((size_t)lrow * gridDim.x + blockIdx.x) * 2 + half) * TK_CAP;   // this thread's candidate buffer
    int count = 0;
    // Rows of a CTA arrive in increasing order, so once the buffer has been pruned to its k best a later row that only
    // TIES the k-th score can never displace it (ties -> lower row): the test is strict, and identical scores (duplicate
    // documents, the zero rows of a padded query lane) do not flood the buffer.  Padded lanes never append.
    float thr = lrow < nq_pass ? -CUDART_INF_F : CUDART_INF_F;

    // Code size matters here: the epilogue runs once per tile, and with the four 32-column chunks unrolled and the prune
    // inlined in each of them it was ~10 k instructions of straight-line code whose FETCH (stall_no_inst) set the tile
    // period (7.6 us per tile against 1.1 us of HBM time).  The chunk loop is therefore a real loop (the TMEM load of a
    // chunk sits inside it), the prune is one out-of-line function, and the per-element append block is branched over
    // whenever the chunk maximum is below the threshold -- ~40 instructions per chunk on the common path.
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const int64_t row0 = (t_beg + i) * TK_BN;
      if (cosine) {                                        // per-document scales of this tile -> smem, read back as broadcasts
        const int te = threadIdx.x - 64;
        if (te < TK_BN) { const int64_t gr = row0 + te; dscale[b * TK_BN + te] = gr < N ? __ldg(dinv + gr) : 0.f; }
        asm volatile("bar.sync 1, %0;" ::"n"(TK_EW * 32) : "memory");   // the epilogue warps only
      }
      mbar_wait(&s_full[b], (i >> 1) & 1);
      tc_fence_after();
      const int valid = (int)((N - row0) < TK_BN ? (N - row0) : TK_BN);   // ragged last tile
#pragma unroll 1
      for (int h = 2 * half; h < 2 * half + 2; ++h) {      // this thread's two 32-column chunks
        uint32_t r[32];
        tmem_ld_x32(tmem_s + lane_addr + (uint32_t)(b * TK_BN + 32 * h), r);
        tmem_ld_wait();
        if (h == 2 * half + 1) {                           // this warp has read its half: the accumulator may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[b]);
        }
        if (cosine) {
          const float4* ds4 = reinterpret_cast<const float4*>(dscale + b * TK_BN + 32 * h);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 dv = ds4[j >> 2];
            r[j + 0] = __float_as_uint(__uint_as_float(r[j + 0]) * (my_qinv * dv.x));
            r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) * (my_qinv * dv.y));
            r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) * (my_qinv * dv.z));
            r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) * (my_qinv * dv.w));
          }
        }
        if (32 * h + 32 > valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (32 * h + j >= valid) r[j] = 0xff800000u;   // -inf: never a candidate
        }
        float m8[4];                                       // maxima of the four 8-column groups, then of the chunk
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          m8[g] = __uint_as_float(r[8 * g]);
#pragma unroll
          for (int jj = 1; jj < 8; ++jj) m8[g] = fmaxf(m8[g], __uint_as_float(r[8 * g + jj]));
        }
        const float mx = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
        // a chunk adds at most 32 keys per query: make room first (warp-uniform decision, cooperative prune)
        const unsigned need = __ballot_sync(0xffffffffu, count > TK_CAP - 32);
        if (need) {
          const u64 ct = tk_prune_lanes(need, gbuf, count, thr, k, lane);
          count = (int)(uint32_t)ct; thr = __uint_as_float((uint32_t)(ct >> 32));
        }
        // Candidates are rare per element (k / rows seen) but 32 queries x 32 documents hold one in most chunks, so the
        // append code is entered per 8-column group and only where SOME query of the warp has a candidate (2-3 of the 4
        // groups mid-scan); inside, the stores are predicated per thread.
        if (__any_sync(0xffffffffu, mx > thr)) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (__any_sync(0xffffffffu, m8[g] > thr)) {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const float sv = __uint_as_float(r[8 * g + jj]);
                if (sv > thr) { gbuf[count] = tk_make_key(sv, (uint32_t)(row0 + 32 * h + 8 * g + jj)); ++count; }
              }
            }
          }
        }
        __syncwarp();
      }
    }
    // leave the k best of every live buffer in front (zero-padded when the CTA saw fewer than k rows): the merge kernel
    // reads k keys per (query, CTA) and ignores zero keys
    const unsigned over = __ballot_sync(0xffffffffu, count > k);
    if (over) count = (int)(uint32_t)tk_prune_lanes(over, gbuf, count, thr, k, lane);
    for (int i = count; i < k; ++i) gbuf[i] = 0ull;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_s, 512);
}

static size_t topk_smem(int H) {
  return 1024 + TK_STAGES * (size_t)TK_BN * H * 2 + 16 * 8 + 16 + 2 * TK_BN * 4 + 64;
}

}  // namespace tc

static bool tc_topk_supported(int H, int k) { return H % 64 == 0 && H >= 64 && H <= 256 && k >= 1 && k <= tc::TK_KMAX; }

struct TcTopkPlan { int grid, tiles_per_cta; size_t qhi, qlo, qinv, cand, total; };
static TcTopkPlan plan_tc_topk(int64_t N, int H, int nq) {
  TcTopkPlan p{};
  const int64_t ntiles = ceil_div(N, tc::TK_BN);
  p.grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  p.tiles_per_cta = (int)ceil_div(ntiles, p.grid);
  p.grid = (int)ceil_div(ntiles, p.tiles_per_cta);
  p.qhi = align_up((size_t)tc::TK_BM * H * 2);
  p.qlo = p.qhi;
  p.qinv = align_up((size_t)tc::TK_BM * 4);
  p.cand = align_up((size_t)tc::TK_BM * p.grid * 2 * tc::TK_CAP * sizeof(u64));
  p.total = p.qhi + p.qlo + p.qinv + p.cand + 1024;
  (void)nq;
  return p;
}

}  // namespace tt

extern "C" {

int tt_topk_scan_batched_ok(int H, int k) { return tt::tc_topk_supported(H, k) ? 1 : 0; }

size_t tt_topk_scan_batched_workspace(int64_t N, int H, int nq) {
  if (N <= 0 || H <= 0 || nq <= 0) return 256;
  return tt::plan_tc_topk(N, H, nq).total;
}

int tt_index_row_inv_norms(const void* index_bf16, int64_t N, int H, float* out, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(index_bf16 && out && N > 0 && H > 0, "index_row_inv_norms: bad arguments");
  tt::tc::tk_row_inv_norm_kernel<<<(unsigned)tt::ceil_div(N, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(index_bf16), N, H, out);
  TT_LAUNCH_CHECK("tk_row_inv_norm_kernel");
  return TT_OK;
}

int tt_topk_scan_batched(const void* index_bf16, const float* queries, int64_t N, int H, int nq, int k,
                         const float* row_inv_norms, int64_t id_offset, float* out_scores, int64_t* out_ids,
                         void* workspace, size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(index_bf16 && queries && out_scores && out_ids && N > 0 && nq > 0, "topk_scan_batched: bad arguments");
  TT_CHECK_ARG(tt::tc_topk_supported(H, k), "topk_scan_batched: needs H %% 64 == 0, H <= 256, k <= %d (H=%d k=%d)", tt::tc::TK_KMAX, H, k);
  TT_CHECK_ARG(k <= N && N < (1ll << 31), "topk_scan_batched: need k <= N < 2^31");
  TT_CHECK_ARG((reinterpret_cast<uintptr_t>(index_bf16) & 15) == 0, "topk_scan_batched: index must be 16-byte aligned");
  const tt::TcTopkPlan plan = tt::plan_tc_topk(N, H, nq);
  if (workspace == nullptr || workspace_bytes < plan.total) { tt::set_error("topk_scan_batched: workspace too small (%zu < %zu)", workspace_bytes, plan.total); return TT_ERR_WORKSPACE; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  tt::Workspace w(workspace, workspace_bytes);
  __nv_bfloat16* qhi = w.take<__nv_bfloat16>((size_t)tt::tc::TK_BM * H);
  __nv_bfloat16* qlo = w.take<__nv_bfloat16>((size_t)tt::tc::TK_BM * H);
  float* qinv = w.take<float>(tt::tc::TK_BM);
  tt::u64* cand = w.take<tt::u64>((size_t)tt::tc::TK_BM * plan.grid * 2 * tt::tc::TK_CAP);
  CUtensorMap tmQhi, tmQlo, tmD;
  int rc = tt::tc::make_tmap_bf16(&tmQhi, qhi, tt::tc::TK_BM, (uint64_t)H, tt::tc::TK_BM); if (rc) return rc;
  rc = tt::tc::make_tmap_bf16(&tmQlo, qlo, tt::tc::TK_BM, (uint64_t)H, tt::tc::TK_BM); if (rc) return rc;
  rc = tt::tc::make_tmap_bf16(&tmD, index_bf16, (uint64_t)N, (uint64_t)H, tt::tc::TK_BN); if (rc) return rc;
  const size_t smem = tt::tc::topk_smem(H);
  TT_CUDA(cudaFuncSetAttribute(tt::tc::tc_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int q0 = 0; q0 < nq; q0 += tt::tc::TK_BM) {          // passes of up to 128 queries: one index read each
    const int nqp = nq - q0 < tt::tc::TK_BM ? nq - q0 : tt::tc::TK_BM;
    tt::tc::tk_split_queries_kernel<<<tt::tc::TK_BM / 8, 256, 0, s>>>(queries + (int64_t)q0 * H, nqp, H, qhi, qlo, qinv);
    TT_LAUNCH_CHECK("tk_split_queries_kernel");
    TT_CUDA(tt::launch_kernel(tt::tc::tc_topk_kernel, dim3((unsigned)plan.grid), dim3(tt::tc::TK_THREADS), smem, s, false, tmQhi, tmQlo,
                              tmD, N, H, k, nqp, plan.tiles_per_cta, (const float*)qinv, row_inv_norms, cand));
    TT_LAUNCH_CHECK("tc_topk_kernel");
    rc = tt::topk_merge_blocked(cand, plan.grid * 2, tt::tc::TK_CAP, nqp, k, id_offset, out_scores + (int64_t)q0 * k,
                                out_ids + (int64_t)q0 * k, s);
    if (rc) return rc;
  }
  return TT_OK;
}

}  // extern "C"
