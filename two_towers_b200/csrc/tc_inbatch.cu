// tc_inbatch.cu -- K4 on the tensor cores (TT_PREC_BF16): in-batch softmax cross-entropy fused
// with its similarity GEMM, forward and backward, tcgen05 + TMEM + TMA.
//
// Reference: twotower/losses.py:107-116 (S = Q D^T; / temperature; cross_entropy vs arange).
//
// forward  (grid = Q row tiles x D splits): the Q tile (128 x H bf16) stays in shared memory;
//   D tiles (64 x H) stream through a 4-stage TMA ring; S = Q D^T lands in a double-buffered
//   TMEM accumulator (128 lanes x 64 fp32 columns); four epilogue warps own one TMEM lane
//   quarter each, so every thread holds ONE row of S and keeps its running (max, sum exp2)
//   in registers -- no shuffles, no shared memory, nothing B x B in HBM.
// backward (same tile loop, roles "X rows / Y columns" swapped for dQ and dD): per Y tile
//   S = X Y^T (TMEM) -> P = exp2(S c - lse c') - [positive]  (registers) -> bf16 P tile written
//   to 128B-swizzled shared memory -> O += P Y as a second tcgen05.mma whose B operand is the
//   SAME TMA-loaded Y tile read MN-major.  O (128 x H fp32) lives in TMEM for the whole loop.
//   S(t+1) is issued before O(t) so the tensor pipe overlaps the exp2 epilogue.
// one-pass form (ce_bwd_body MODE 2, the trainer's default): unit-norm rows bound every logit, so
//   E = exp(logit - bound) needs no running maximum: the loss forward and dQ come out of ONE pass
//   over S, the row normaliser is applied in the cluster tail.
// stored-E form (MODE 2 with e_store + MODE 3, while B_q x B_d bf16 fits in L2): MODE 2 also TMA-stores
//   every E tile it forms; the document gradient is then the plain product dD = E^T (X / L) -- both
//   operands MN-major, no epilogue in the loop, S formed once per step.
// Each output row has one owner CTA per split and splits are summed in split order:
// bitwise-reproducible gradients.
#include <math_constants.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "tc_common.cuh"
#include "tensor_core.cuh"

namespace tt {

// provided by inbatch_ce.cu
int split_sum(const float* partial, int nsplit, int64_t n, float* out, cudaStream_t s);
int inbatch_ce_fwd_fp32(const float* q, const float* d, int64_t Bq, int64_t Bd, int H, float inv_temp, int64_t off,
                        float loss_scale, float* loss, float* lse, float* pos_mean, void* ws, size_t ws_bytes, cudaStream_t s);
int inbatch_ce_bwd_fp32(const float* q, const float* d, const float* lse, int64_t Bq, int64_t Bd, int H, float inv_temp,
                        int64_t off, float loss_scale, const float* grad_out, float* dq, float* dd, void* ws, size_t ws_bytes,
                        cudaStream_t s);
size_t inbatch_ce_fp32_workspace(int64_t Bq, int64_t Bd, int H);

namespace tc {

constexpr int CE_BM = 128;          // X rows per CTA (UMMA M)
constexpr int CE_BN = 64;           // Y rows per tile (UMMA N of the S product, K of the O product)
constexpr int CE_THREADS = 192;     // backward with 4 epilogue warps: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (EW = 8: 320 threads)
constexpr int FWD_THREADS = 320;    // forward: warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {      // MUFU.EX2, inputs are bounded (<= 0 after max-subtraction)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / integer pipes for arguments <= 0: round-to-nearest split x = n + f (magic-number add), a polynomial
// for 2^f on [-0.5, 0.5] and the exponent patched in with an integer add (FlashAttention-4's trick: the special-function
// unit retires 16 exp2 per clock and SM -- 1024 cycles for a 128 x 128 score tile, as long as the tile's MMAs).
// Relative error 1.4e-4 (degree 3) / 5e-6 (degree 4).  MEASURED (gpurun r02f, mask 0x13 = 3 of every 8 elements): no gain
// -- forward 15.0 -> 15.3 us, backward 34.7 -> 35.1 us: with eight epilogue warps the loops are bound by the dependent
// TMEM-load -> exp -> pack -> shared-store chain and by issue slots, not by MUFU throughput, and the eight extra
// FMA-pipe instructions per element cost what the MUFU slot saved.  Kept behind the mask (0 = every element on MUFU).
constexpr unsigned kExpPolyMask = 0x0;                    // bit e set: element e (mod 8) of a row uses the polynomial
template <int DEG>
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float xf = x + 12582912.0f;                       // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (xf - 12582912.0f);
  float p;
  if (DEG == 3) {
    p = fmaf(0.05502927f, f, 0.24225698f); p = fmaf(p, f, 0.69325305f); p = fmaf(p, f, 0.99995134f);
  } else {
    p = fmaf(0.00955411f, f, 0.05587041f); p = fmaf(p, f, 0.24024697f); p = fmaf(p, f, 0.69312803f); p = fmaf(p, f, 0.99999944f);
  }
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));   // (bits(xf) - bits(magic)) << 23 == bits(xf) << 23 (mod 2^32)
}
template <int DEG>
__device__ __forceinline__ float exp2_sel(float x, int e) {   // e: compile-time element index
  return ((kExpPolyMask >> (e & 7)) & 1u) ? poly_exp2<DEG>(x) : fast_exp2(x);
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
// D tiles are 128 rows: one tcgen05.mma covers N = 128 (a 128x64x16 instruction was measured at ~66 cycles, the
// same as a 128x128x16 one, so 64-wide tiles run the tensor pipe at half rate).  The Q tile is staged through the
// third D stage into TMEM once and then read from there by every MMA (TS mode).
constexpr int FWD_BN = 128;
constexpr int FWD_STAGES = 3;

// In-kernel finalisation (optional): the last CTA of a row tile to finish merges that tile's per-split (max, sum)
// pairs into lse and a per-tile loss partial; the last row tile to finish adds the tile partials in tile order.
// Fixed orders everywhere, so the result does not depend on which CTA happens to be last.
struct FwdFinalize {
  unsigned* counters;          // [row tiles + 1], zero on entry, zero again on exit (null: finalise in a second launch)
  float* tile_sums;            // [row tiles][2]: sum(lse - pos), sum(pos)
  float* lse; float* loss; float* pos_mean;
  float loss_scale;
  long long* dbg;              // developer aid (TT_CE_DEBUG): [cta][4] %globaltimer stamps
};

__global__ void __launch_bounds__(FWD_THREADS, 1)
tc_ce_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmD, int64_t Bq,
                 int64_t Bd, int H, float inv_temp, int64_t label_offset, int tiles_per_split,
                 int64_t d_blk, int64_t d_blk_stride, int64_t d_blk_off,
                 float* __restrict__ part_ml, float* __restrict__ pos_logit, const FwdFinalize fin) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  long long* fdbg = fin.dbg ? fin.dbg + 4 * (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
#define TT_FWD_STAMP(slot) do { if (fdbg && threadIdx.x == 64) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); fdbg[slot] = t_; } } while (0)
  TT_FWD_STAMP(0);
  long long* tl = (fin.dbg && blockIdx.x == 0 && blockIdx.y == 0) ? fin.dbg + 4 * gridDim.x * gridDim.y : nullptr;
#define TT_FTL(role, tile, slot) do { if (tl && (tile) < 64) tl[((role) * 64 + (tile)) * 8 + (slot)] = clock64(); } while (0)
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);     // stays in the shared address space
  const int kq = H / 64;                                  // 64-wide K blocks
  const uint32_t d_bytes = (uint32_t)FWD_BN * H * 2;      // == Q tile bytes (both are 128 rows)
  uint8_t* d_tiles = base;
  uint8_t* q_tile = base + (FWD_STAGES - 1) * d_bytes;    // Q borrows the last stage until it sits in TMEM
  uint64_t* bars = reinterpret_cast<uint64_t*>(d_tiles + FWD_STAGES * d_bytes);
  uint64_t* q_bar = bars;
  uint64_t* d_full = bars + 1;
  uint64_t* d_empty = d_full + FWD_STAGES;
  uint64_t* s_full = d_empty + FWD_STAGES;                // [2]
  uint64_t* s_empty = s_full + 2;                         // [2]
  uint64_t* q_ready = s_empty + 2;                        // Q tile copied into TMEM (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  float* ml_s = reinterpret_cast<float*>(tmem_slot + 2);          // [128][2] (max, sum) of the second column half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t x0 = (int64_t)blockIdx.x * CE_BM;
  const int ntiles = (int)ceil_div(Bd, FWD_BN);
  const int t_beg = blockIdx.y * tiles_per_split;
  const int t_end = min(ntiles, t_beg + tiles_per_split);
  const int nt = max(0, t_end - t_beg);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmD);
    mbar_init(q_bar, 1);
    for (int s = 0; s < FWD_STAGES; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 8); }
    mbar_init(q_ready, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                             // prologue overlapped the previous kernel's tail
  const uint32_t tmem_s = *tmem_slot;                     // columns [0,256): two S buffers of 128
  const uint32_t tmem_q = tmem_s + 256;                   // columns [256, 256 + H/2): Q tile (TMEM A operand)

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_bar, d_bytes);
      for (int kb = 0; kb < kq; ++kb) tma_load_2d(q_tile + kb * (CE_BM * 128), &tmQ, q_bar, kb * 64, (int)x0);
    }
    __syncwarp();
    for (int i = 0; i < nt; ++i) {                        // whole warp, uniform control flow; one lane issues
      const int s = i % FWD_STAGES;
      if (i == FWD_STAGES - 1) mbar_wait(q_ready, 0);     // the stage Q borrowed is free once Q lives in TMEM
      mbar_wait(&d_empty[s], ((i / FWD_STAGES) & 1) ^ 1);
      uint8_t* dt = d_tiles + s * d_bytes;
      const int64_t g = (int64_t)(t_beg + i) * FWD_BN;    // logical D row -> physical row of the (gathered) buffer
      const int yc = (int)((g / d_blk) * d_blk_stride + (g % d_blk) + d_blk_off);
      if (elect_one()) {
        mbar_arrive_expect_tx(&d_full[s], d_bytes);
        for (int kb = 0; kb < kq; ++kb) tma_load_2d(dt + kb * (FWD_BN * 128), &tmD, &d_full[s], kb * 64, yc);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(CE_BM, FWD_BN, 0, 0);
    const uint64_t dd0 = umma_desc_kmajor(smem_u32(d_tiles), 0);
    mbar_wait(q_ready, 0);
    tc_fence_after();
    // This warp shares its scheduler with two busy epilogue warps: every instruction it does not issue is tensor-pipe
    // time won back.  The issuing lane is elected ONCE; descriptors advance by loop-carried adds.
    const bool leader = elect_one();
    for (int i = 0; i < nt; ++i) {
      const int s = i % FWD_STAGES, b = i & 1;
      if (lane == 0) TT_FTL(0, i, 0);
      mbar_wait(&d_full[s], (i / FWD_STAGES) & 1);
      if (lane == 0) TT_FTL(0, i, 1);
      mbar_wait(&s_empty[b], ((i >> 1) & 1) ^ 1);
      if (lane == 0) TT_FTL(0, i, 2);
      tc_fence_after();
      if (leader) {
        uint64_t dd = dd0 + (uint64_t)((s * d_bytes) >> 4);
        uint32_t ta = tmem_q;
        const uint32_t td = tmem_s + b * FWD_BN;
        for (int kb = 0; kb < kq; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(td, ta + (uint32_t)(k * 8), dd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          dd += FWD_BN * 128 / 16;
          ta += 32;
        }
        umma_commit(&d_empty[s]);
        umma_commit(&s_full[b]);
      }
      __syncwarp();
      if (lane == 0) TT_FTL(0, i, 3);
    }
  } else {
    // epilogue warps 2..9: TMEM lane quarter = warp % 4; the two warps of a quarter take the two 64-column halves of every
    // S tile (online max / sum is associative: the halves are merged once at the end).  The epilogue, not the tensor
    // pipe, bounds this kernel, so it gets two warps per scheduler.
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int lrow = quarter * 32 + lane;
    const int64_t row = x0 + lrow;
    const int64_t pcol = row + label_offset;
    const float c = inv_temp * kLog2e;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    // Q tile: smem (TMA, 128B swizzle) -> registers -> TMEM, so every S product reads A from tensor memory
    mbar_wait(q_bar, 0);
    for (int kb = half; kb < kq; kb += 2) {
      uint32_t xr[32];
      const uint8_t* xrow = q_tile + kb * (CE_BM * 128) + lrow * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 v = *reinterpret_cast<const uint4*>(xrow + ((ch ^ (lrow & 7)) << 4));
        xr[4 * ch] = v.x; xr[4 * ch + 1] = v.y; xr[4 * ch + 2] = v.z; xr[4 * ch + 3] = v.w;
      }
      tmem_st_x32(tmem_q + lane_addr + (uint32_t)(kb * 32), xr);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(q_ready);
    TT_FWD_STAMP(1);
    float m = -CUDART_INF_F, l = 0.f;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const int64_t y0 = (int64_t)(t_beg + i) * FWD_BN + 64 * half;      // first column of this warp's half tile
      if (threadIdx.x == 64) TT_FTL(1, i, 0);
      mbar_wait(&s_full[b], (i >> 1) & 1);
      if (threadIdx.x == 64) TT_FTL(1, i, 1);
      tc_fence_after();
      uint32_t r[2][32];
      const uint32_t ta = tmem_s + lane_addr + (uint32_t)(b * FWD_BN + 64 * half);
#pragma unroll
      for (int h = 0; h < 2; ++h) tmem_ld_x32(ta + 32 * h, r[h]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[b]);             // TMEM buffer may be overwritten
      if (threadIdx.x == 64) TT_FTL(1, i, 2);
      if (y0 + 64 > Bd) {                                  // ragged last tile: padded columns -> -inf
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (y0 + 32 * h + j >= Bd) r[h][j] = 0xff800000u;
      }
      const int64_t pj = pcol - y0;                        // this row's positive column inside the half tile
      if (pj >= 0 && pj < 64 && row < Bq) {
        float pv = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 32; ++j) pv = (32 * h + j == (int)pj) ? __uint_as_float(r[h][j]) : pv;
        pos_logit[row] = pv * inv_temp;
      }
      float t0 = -CUDART_INF_F, t1 = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        t0 = fmaxf(t0, fmaxf(__uint_as_float(r[0][j]), __uint_as_float(r[1][j])));
        t1 = fmaxf(t1, fmaxf(__uint_as_float(r[0][j + 1]), __uint_as_float(r[1][j + 1])));
      }
      const float mnew = fmaxf(m, fmaxf(t0, t1));
      if (threadIdx.x == 64) TT_FTL(1, i, 3);
      const bool seen = mnew > -CUDART_INF_F;              // false while this warp's half tiles have all been padding
      const float mc = seen ? mnew * c : 0.f;              // exp2(-inf * c - 0) = 0, never inf - inf
      // two phases: all exponentials first (independent MUFU ops, back to back), then the sums -- an accumulate chain
      // fed directly by MUFU results exposes the MUFU latency on every element
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 32; ++j) r[h][j] = __float_as_uint(exp2_sel<4>(fmaf(__uint_as_float(r[h][j]), c, -mc), j));
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        s0 += __uint_as_float(r[0][j]); s1 += __uint_as_float(r[0][j + 1]);
        s2 += __uint_as_float(r[1][j]); s3 += __uint_as_float(r[1][j + 1]);
      }
      l = (seen ? l * fast_exp2((m - mnew) * c) : 0.f) + ((s0 + s1) + (s2 + s3));
      m = mnew;
      if (threadIdx.x == 64) TT_FTL(1, i, 4);
    }
    TT_FWD_STAMP(2);
    // merge the two column halves of every row (fixed order: half 0, then half 1)
    if (half == 1) { ml_s[2 * lrow] = m; ml_s[2 * lrow + 1] = l; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (half == 0) {
      const float m1 = ml_s[2 * lrow], l1 = ml_s[2 * lrow + 1];
      const float M = fmaxf(m, m1);
      if (nt > 0 && M > -CUDART_INF_F)                       // a half that saw only padding has (m, l) = (-inf, 0)
        l = (m > -CUDART_INF_F ? l * fast_exp2((m - M) * c) : 0.f) + (m1 > -CUDART_INF_F ? l1 * fast_exp2((m1 - M) * c) : 0.f);
      m = M;
      if (row < Bq) {                                        // a split without tiles contributes (-inf, 0)
        part_ml[((int64_t)blockIdx.y * Bq + row) * 2 + 0] = nt > 0 ? m * inv_temp : -CUDART_INF_F;
        part_ml[((int64_t)blockIdx.y * Bq + row) * 2 + 1] = l;
      }
    }
    if (fin.counters) __threadfence();                       // partials visible before this CTA takes its ticket
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_s, 512);
  TT_FWD_STAMP(3);
  if (fin.counters == nullptr) return;

  // ---- finalisation by whichever CTA of this row tile finishes last ----
  __shared__ unsigned s_ticket;
  __shared__ float s_red[2][4];
  if (threadIdx.x == 0) { s_ticket = atomicAdd(&fin.counters[blockIdx.x], 1u); __threadfence(); }
  __syncthreads();
  if (s_ticket != gridDim.y - 1) return;
  if (warp >= 2 && warp < 6) {
    const int lrow = (warp & 3) * 32 + lane;
    const int64_t row = x0 + lrow;
    float dl = 0.f, dp = 0.f;
    if (row < Bq) {
      float M = -CUDART_INF_F, L = 0.f;
      for (int sp = 0; sp < (int)gridDim.y; ++sp) {            // split order
        const float2 ml = __ldcg(reinterpret_cast<const float2*>(part_ml) + (int64_t)sp * Bq + row);
        const float Mn = fmaxf(M, ml.x);
        L = L * expf(M - Mn) + ml.y * expf(ml.x - Mn);
        M = Mn;
      }
      const float v = M + logf(L);
      const float pl = __ldcg(pos_logit + row);
      fin.lse[row] = v;
      dl = v - pl; dp = pl;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dl += __shfl_xor_sync(0xffffffffu, dl, o); dp += __shfl_xor_sync(0xffffffffu, dp, o); }
    if (lane == 0) { s_red[0][warp & 3] = dl; s_red[1][warp & 3] = dp; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    fin.tile_sums[2 * blockIdx.x + 0] = (s_red[0][0] + s_red[0][1]) + (s_red[0][2] + s_red[0][3]);
    fin.tile_sums[2 * blockIdx.x + 1] = (s_red[1][0] + s_red[1][1]) + (s_red[1][2] + s_red[1][3]);
    fin.counters[blockIdx.x] = 0u;                           // leave the scratch zeroed for the next launch
    __threadfence();
    s_ticket = atomicAdd(&fin.counters[gridDim.x], 1u);
    __threadfence();
  }
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  if (warp == 0) {                                          // last row tile: tile partials in tile order
    float tl = 0.f, tp = 0.f;
    for (int t = lane; t < (int)gridDim.x; t += 32) { tl += __ldcg(fin.tile_sums + 2 * t); tp += __ldcg(fin.tile_sums + 2 * t + 1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { tl += __shfl_xor_sync(0xffffffffu, tl, o); tp += __shfl_xor_sync(0xffffffffu, tp, o); }
    if (lane == 0) {
      *fin.loss = tl * fin.loss_scale;
      if (fin.pos_mean) *fin.pos_mean = tp / (inv_temp * (float)Bq);
      fin.counters[gridDim.x] = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward: dQ and dD passes in ONE launch (blockIdx.z selects the pass)
// ---------------------------------------------------------------------------------------
constexpr int BWD_STAGES = 3;          // 3 x 64 KB Y stages (X staged through the last one) + 32 KB P = 224 KB

struct BwdParams {
  // index 0: dQ pass (X = queries, Y = documents, lse indexed by X row, positive at col == row + off)
  // index 1: dD pass (X = documents, Y = queries, lse indexed by Y row, positive at row == col + off)
  const float* lse[2];         // natural-log logsumexp of the logits
  int64_t Bx[2], By[2];        // gradient rows / logical rows scored against
  int64_t label_offset[2];
  int64_t y_blk[2], y_blk_stride[2], y_blk_off[2];   // logical Y row g -> physical (g / blk) * stride + g % blk + off
  int tiles_per_split[2];      // Y tiles handled by one split
  float* out[2];               // [nsplit] slices (part_stride elements apart) or the final tensor
  int64_t part_stride[2];
  // fused normalise-backward epilogue (optional, per pass): instead of fp32 slices the kernel finishes
  //   dz = (dy - y (y . dy)) * inv_norm   with y = this CTA's X tile (the unit rows themselves), dy = sum of the splits,
  // writes it as bf16 plus per-32-row column sums (-> bias gradient).  Needs <= 2 splits (CTA pair = cluster).
  __nv_bfloat16* dz[2];        // [Bx, H] bf16 (null: write out[] slices as before)
  float* dz_colsum[2];         // [ceil(Bx / 32), H]
  const float* inv_norm[2];    // [Bx]
  const __nv_bfloat16* xg[2];  // the X operand in global memory (== y), re-read by the fused tail
  int H;
  float inv_temp;
  const float* grad_out;       // nullable device scalar
  float coef;                  // loss_scale / temperature
  int dbg_pass;                // pass whose CTA (0,0) is stamped (TT_CE_DEBUG=<pass>)
  long long* dbg;              // optional timeline buffer (TT_CE_DEBUG=1): [role 0..1][tile][8] clock64 stamps of CTA (0,0,0)
  int pass_base;               // pass = blockIdx.z + pass_base (1: only the dD pass is launched)
  const __nv_bfloat16* yg[2];  // the Y operand in global memory (MODE 2 re-reads each row's positive from it)
  // MODE 2 (one-pass forward + query gradient, tc_ce_fwd_dq_kernel): logits are bounded by mfix, so the softmax needs no
  // running maximum -- E = exp(logit - mfix) is accumulated as it is produced and normalised once per row at the end.
  float mfix;                  // >= every logit (logit units, i.e. already divided by the temperature)
  float* lse_out;              // [Bx[0]] natural-log logsumexp (saved for the dD pass)
  float* loss; float* pos_mean; float loss_scale;
  unsigned* counter;           // one ticket counter, zero on entry, zero again on exit
  float* tile_sums;            // [row tiles * splits][2]: sum(lse - pos), sum(pos)
  // MODE 2 behind a peer-memory all-gather of Y (tt_inbatch_ce_fwd_dq_p2p): the kernel is launched programmatically behind
  // the exchange kernel and does NOT wait for it to retire; the TMA warp polls the exchange's per-source arrival counters
  // and consumes each rank's block of Y as it lands, its own rank's block first.
  const unsigned* gate;        // local exchange header (null: no gate): word r = arrivals from rank r, word 36 = this round's target
  int gate_rank;               // this rank: its block of Y is consumed first, straight from the exchange's SOURCE (y_own)
  const __nv_bfloat16* y_own;  // [gate_blk, H]: this rank's rows of Y where the exchange kernel reads them
  int64_t gate_blk;            // logical Y rows per source rank
  unsigned gate_timeout_s;
  // Stored-E form of the document gradient (single-process / per-rank negatives, E small enough to stay in L2): MODE 2 also
  // writes every P tile it forms (E = exp(logit - mfix) as bf16, the positives left out) to e[x_rows, e_pitch] with a TMA
  // store, the rows x_i / L_i (bf16) to xs_out and 1 - P_pos to wpos_out; MODE 3 then forms the document gradient as the plain
  // product  dD = E^T (X / L)  minus the positives' rank-one terms -- S is not recomputed and no exponential is taken twice.
  int e_store;                 // MODE 2: 1 = write E / xs_out / wpos_out
  __nv_bfloat16* xs_out;       // [Bx[0], H] bf16
  float* wpos_out;             // [Bx[0]]
  const float* wpos_in;        // MODE 3: [By[1]] 1 - P_pos of every query
};

// X rows = this CTA's 128 output rows, Y = streamed 128-row tiles.
//   COL == false: X = Q, Y = D, lse indexed by X row,  positive at col == row + off   (dQ)
//   COL == true : X = D, Y = Q, lse indexed by Y row,  positive at row == col + off   (dD)
// Shared memory: 3 Y stages x 64 KB (the X tile is staged through the last one into TMEM) + one 32 KB P tile.
// Tensor memory: O [0,256) | S [256,384) | X [384,512).
// Tensor pipe order  S(0) S(1) O(0) S(2) O(1) ...: S(i+1) runs while the warps turn S(i) into P(i); every
// tcgen05.mma has N >= 128 (a 128x64x16 instruction was measured at the same ~66 cycles as 128x128x16).
constexpr int BWD_BN = 128;

// EW = epilogue warps: 4 (one per TMEM lane quarter, a thread turns a whole 128-column S row into P) or 8 (two per quarter,
// each taking a 64-column half of every tile, like the forward kernel).  With four warps the epilogue of the dD pass
// (column lse read back from shared memory as broadcasts) takes longer than the two MMAs of a tile, so that pass -- and
// with it the launch -- ran ~25 % behind the tensor pipe; the run-once tail (accumulator dump, normalise backward) is
// also instruction-latency bound with one warp per scheduler.  Eight warps halve both.
template <int MODE, int EW>
__device__ __forceinline__ void ce_bwd_body(const CUtensorMap* tmX, const CUtensorMap* tmY, const BwdParams& p,
                                            uint8_t* base, const CUtensorMap* tmYown = nullptr, int tmem_mode = 0,
                                            const CUtensorMap* tmE = nullptr) {
  // tmem_mode: 0 = allocate and free tensor memory here; 1 = allocate, leave it to a second body of the same kernel;
  // 2 = re-use that allocation and free it (the allocation permit is relinquished after the first tcgen05.alloc)
  constexpr bool COL = MODE == 1;                           // MODE 0: dQ from saved lse, 1: dD, 2: forward + dQ in one pass,
  constexpr int PASS = (MODE == 1 || MODE == 3) ? 1 : 0;    //      3: dD from the E tiles MODE 2 stored (tmY = X / L, tmE = E)
  constexpr int NH = EW / 4;                                // warps per TMEM lane quarter
  constexpr int CW = BWD_BN / NH;                           // S columns per epilogue thread
  constexpr int HC = CW / 32;                               // 32-column register chunks per thread
  constexpr int ETH = EW * 32;                              // epilogue threads
  const int64_t Bx = p.Bx[PASS], By = p.By[PASS], label_offset = p.label_offset[PASS];
  const int tiles_per_split = p.tiles_per_split[PASS];
  const float* __restrict__ lse = p.lse[PASS];
  float* out = p.out[PASS] + (int64_t)blockIdx.y * p.part_stride[PASS];
  const int H = p.H;
  const int kq = H / 64;
  const uint32_t y_bytes = (uint32_t)BWD_BN * H * 2;      // == X tile bytes (both are 128 rows)
  constexpr uint32_t p_bytes = CE_BM * BWD_BN * 2;        // 32 KB: K-major, 2 k-blocks of [128 rows x 128 B]
  uint8_t* y_tiles = base;
  uint8_t* x_tile = base + (BWD_STAGES - 1) * y_bytes;    // X borrows the last Y stage until it sits in TMEM
  uint8_t* p_tile = y_tiles + BWD_STAGES * y_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_tile + p_bytes);
  uint64_t* x_bar = bars;
  uint64_t* y_full = bars + 1;
  uint64_t* y_empty = y_full + BWD_STAGES;
  uint64_t* s_full = y_empty + BWD_STAGES;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_empty + 1;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint64_t* x_ready = o_full + 1;                         // X tile copied into TMEM (4 epilogue warps)
  uint64_t* xfer_bar = x_ready + 1;                       // fused tail: the peer CTA's accumulator half has landed here
  uint64_t* sc_bar = xfer_bar + 1;                        // MODE 2 tail: the peers' per-row scalars have landed here
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sc_bar + 1);
  float* col_lse = reinterpret_cast<float*>(tmem_slot + 6);   // [2][128] column lse of the current / next Y tile (COL); 16-byte aligned
  float* red_s = col_lse + 2 * BWD_BN;                    // [EW][2] MODE 2: per-warp loss partials

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t x0 = (int64_t)blockIdx.x * CE_BM;
  const int ntiles = (int)ceil_div(By, BWD_BN);
  const int t_beg = blockIdx.y * tiles_per_split;
  const int t_end = min(ntiles, t_beg + tiles_per_split);
  // gated: every split takes the same slice of EVERY source rank's block and walks the blocks starting with its own
  // rank's (local data first, the peers' blocks have time to arrive); otherwise one contiguous tile range per split
  const bool rot = MODE == 2 && p.gate != nullptr;
  const int g_nblk = rot ? (int)(By / p.gate_blk) : 1, g_tpb = rot ? (int)(p.gate_blk / BWD_BN) : 1;
  const int g_tps = rot ? g_tpb / (int)gridDim.y : 1;
  // MODE 3: the loop runs over 64-row K blocks of the split's query range (E3_ST stages of [64 x H] X/L + [64 x 128] E)
  constexpr int E3_ST = 4, E3_BK = 64;
  const int k3_beg = t_beg * (BWD_BN / E3_BK);
  const int k3_end = min((int)ceil_div(By, (int64_t)E3_BK), t_end * (BWD_BN / E3_BK));
  const int nt = MODE == 3 ? max(0, k3_end - k3_beg) : rot ? g_nblk * g_tps : max(0, t_end - t_beg);
  const bool estore = MODE == 2 && p.e_store != 0;
  auto tile_at = [&](int i) -> int {
    if (!rot) return t_beg + i;
    const int b = i / g_tps;
    return ((p.gate_rank + b) % g_nblk) * g_tpb + (int)blockIdx.y * g_tps + (i - b * g_tps);
  };
  const bool fused = MODE >= 2 || p.dz[PASS] != nullptr;  // CTA-uniform (cluster-uniform): finish through the cluster tail
  float lsum = 0.f, pos_val = 0.f;                         // MODE 2: this thread's share of sum_j E_ij, its row's positive logit
  bool pos_found = false;
  long long* dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (int)blockIdx.z == p.dbg_pass) ? p.dbg : nullptr;
#define TT_STAMP(role, tile, slot) do { if (dbg) dbg[((role) * 64 + (tile)) * 8 + (slot)] = clock64(); } while (0)
#define TT_TAIL(k) do { if (dbg && threadIdx.x == 64) dbg[(64 + 40) * 8 + (k)] = clock64(); } while (0)
  // per-CTA wall-clock stamps (ns): kernel entry, X resident in TMEM, main loop done, outputs stored
  long long* cta_dbg = p.dbg ? p.dbg + 2 * 64 * 8 + 8 * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
#define TT_CTA_STAMP(slot) do { if (cta_dbg && threadIdx.x == 64) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); cta_dbg[slot] = t_; } } while (0)
  TT_CTA_STAMP(0);

  if (threadIdx.x == 0) {
    if (MODE != 3) tma_prefetch_desc(tmX);
    tma_prefetch_desc(tmY);
    if (MODE == 3 || estore) tma_prefetch_desc(tmE);
    mbar_init(x_bar, 1);
    if (MODE == 3) {                                        // full[s] = y_full + s, empty[s] = y_full + E3_ST + s (8 consecutive words)
      for (int s = 0; s < 2 * E3_ST; ++s) mbar_init(&y_full[s], 1);
    } else {
      for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
      mbar_init(s_full, 1); mbar_init(s_empty, EW);
      mbar_init(p_full, EW); mbar_init(p_empty, estore ? 2 : 1);   // stored-E: the TMA store of the P tile has read it, too
    }
    mbar_init(o_full, 1);
    mbar_init(x_ready, EW);
    mbar_init(xfer_bar, 1);
    fence_barrier_init();
    // cluster tail: the peers' accumulator rows (bulk copies) and, MODE 2, per-row scalars (st.async) complete on this
    // barrier; armed long before anybody sends
    if (fused && gridDim.y > 1) {
      const uint32_t rows = (gridDim.y - 1) * (CE_BM / gridDim.y);
      mbar_arrive_expect_tx(xfer_bar, rows * (uint32_t)H * 2u + (MODE == 2 ? rows * 8u : 0u));   // fp16 partials
    }
  }
  if (warp == 1 && tmem_mode != 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const bool gated = MODE == 2 && p.gate != nullptr;
  if (!gated) pdl_wait();                                 // prologue overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base;                      // columns [0, H)
  const uint32_t tmem_s = tmem_base + 256;                // columns [256, 384): S tile
  const uint32_t tmem_x = tmem_base + 384;                // columns [384, 384 + H/2): X tile (TMEM A operand)

  if (MODE == 3) {
    // dD = E^T (X / L): a plain TMA -> tcgen05 pipeline, both operands read MN-major (E is stored [query][document], X / L
    // [query][H]; the contraction runs over queries), no epilogue inside the loop.  Per 64-query block a CTA pulls 16 KB of E
    // and 64 H / 128 x 16 KB of X / L through L2 for 512 tensor-pipe cycles: the loop is bound by L2 bandwidth, not by the MMAs.
    const uint32_t q_bytes = (uint32_t)E3_BK * H * 2;       // X / L block: kq boxes of [64 rows x 128 B]
    constexpr uint32_t e_bytes = E3_BK * CE_BM * 2;         // E block: 2 boxes of [64 rows x 128 B] = this CTA's 128 documents
    const uint32_t st_bytes = q_bytes + e_bytes;
    uint64_t* full3 = y_full;
    uint64_t* empty3 = y_full + E3_ST;
    if (warp == 0) {
      for (int i = 0; i < nt; ++i) {
        const int s = i % E3_ST;
        mbar_wait(&empty3[s], ((i / E3_ST) & 1) ^ 1);
        uint8_t* st = y_tiles + s * st_bytes;
        const int r0 = (k3_beg + i) * E3_BK;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full3[s], st_bytes);
          for (int kb = 0; kb < kq; ++kb) tma_load_2d(st + kb * (E3_BK * 128), tmY, &full3[s], kb * 64, r0);
          tma_load_2d(st + q_bytes, tmE, &full3[s], (int)x0, r0);
          tma_load_2d(st + q_bytes + E3_BK * 128, tmE, &full3[s], (int)x0 + 64, r0);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      if (nt > 0) {
        const uint32_t idesc = umma_idesc_bf16(CE_BM, H, 1, 1);
        const uint64_t dq0 = umma_desc_mnmajor(smem_u32(y_tiles), 0, E3_BK * 128);
        const uint64_t de0 = umma_desc_mnmajor(smem_u32(y_tiles) + q_bytes, 0, E3_BK * 128);
        for (int i = 0; i < nt; ++i) {
          const int s = i % E3_ST;
          mbar_wait(&full3[s], (i / E3_ST) & 1);
          tc_fence_after();
          const uint64_t off = (uint64_t)((s * st_bytes) >> 4);
#pragma unroll
          for (int k = 0; k < E3_BK / 16; ++k) {              // 16 query rows per instruction: +2048 B in both operands
            if (elect_one()) umma_bf16(tmem_o, de0 + off + (uint64_t)(k * 128), dq0 + off + (uint64_t)(k * 128), idesc, (i | k) != 0);
          }
          if (elect_one()) umma_commit(&empty3[s]);
          __syncwarp();
        }
        if (elect_one()) umma_commit(o_full);
        __syncwarp();
      }
    } else {
      TT_CTA_STAMP(1);
      if (nt > 0) { mbar_wait(o_full, 0); tc_fence_after(); }
      TT_CTA_STAMP(2);
    }
  } else if (warp == 0) {
    unsigned g_target = 0, g_have = 0;                      // gate: this round's arrival target, bitmask of ranks seen arrived
    auto gate_wait = [&](int src) {                         // whole warp: every lane acquires, so the elected one has
      if (!gated || ((g_have >> src) & 1u)) return;
      unsigned long long t0 = 0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (unsigned spins = 1;; ++spins) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.gate + src) : "memory");
        if (v >= g_target) break;
        if ((spins & 255u) == 0) {                          // a dead peer must not hang the GPU: the exchange kernel reports it
          unsigned long long t1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > (unsigned long long)p.gate_timeout_s * 1000000000ull) break;
        }
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");   // the acquired data is read by TMA (async proxy)
      __syncwarp();
      g_have |= 1u << src;
    };
    if (gated) {                                            // X and this rank's own Y block need no gate: the exchange kernel
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g_target) : "l"(p.gate + 36) : "memory");   // triggered this launch after its own griddepcontrol.wait
      g_have = 1u << p.gate_rank;
    }
    if (elect_one()) {
      mbar_arrive_expect_tx(x_bar, y_bytes);
      for (int kb = 0; kb < kq; ++kb) tma_load_2d(x_tile + kb * (CE_BM * 128), tmX, x_bar, kb * 64, (int)x0);
    }
    __syncwarp();
    // stored-E form: this (otherwise idle) warp also sends every finished P tile -- two [128 x 64] swizzled boxes -- to
    // E[x0.., y0..] with a TMA store and gives p_empty its second arrival once the store has read the tile.  It runs two
    // tiles behind its own loads: the stage load(i + 1) needs is released by O(i - 2), i.e. after P(i - 2) anyway.
    auto store_e = [&](int t) {
      mbar_wait(p_full, t & 1);
      if (lane == 0) {
        const int yc = tile_at(t) * BWD_BN;
        tma_store_2d(tmE, p_tile, yc, (int)x0);
        tma_store_2d(tmE, p_tile + CE_BM * 128, yc + 64, (int)x0);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(p_empty);
      }
      __syncwarp();
    };
    for (int i = 0; i < nt; ++i) {                        // whole warp, uniform control flow; one lane issues
      const int s = i % BWD_STAGES;
      if (i == BWD_STAGES - 1) mbar_wait(x_ready, 0);     // the stage X borrowed is free once X lives in TMEM
      mbar_wait(&y_empty[s], ((i / BWD_STAGES) & 1) ^ 1);
      uint8_t* yt = y_tiles + s * y_bytes;
      const int64_t g = (int64_t)tile_at(i) * BWD_BN;
      const CUtensorMap* tm = tmY;
      int yc;
      if (gated && g / p.gate_blk == p.gate_rank) {       // this rank's own rows: read where the towers left them
        tm = tmYown; yc = (int)(g - (int64_t)p.gate_rank * p.gate_blk);
      } else {
        if (gated) gate_wait((int)(g / p.gate_blk));
        yc = (int)((g / p.y_blk[PASS]) * p.y_blk_stride[PASS] + (g % p.y_blk[PASS]) + p.y_blk_off[PASS]);
      }
      if (elect_one()) {
        mbar_arrive_expect_tx(&y_full[s], y_bytes);
        for (int kb = 0; kb < kq; ++kb) tma_load_2d(yt + kb * (BWD_BN * 128), tm, &y_full[s], kb * 64, yc);
      }
      __syncwarp();
      if (estore && i >= BWD_STAGES - 1) store_e(i - (BWD_STAGES - 1));
    }
    if (estore) {
      for (int i = max(0, nt - (BWD_STAGES - 1)); i < nt; ++i) store_e(i);
      if (lane == 0) tma_store_wait();
      __syncwarp();
    }
  } else if (warp == 1) {
    if (nt > 0) {
      // whole warp in uniform control flow: descriptors live in uniform registers, only the tcgen05
      // instructions are predicated on the elected lane
      const uint32_t idesc_s = umma_idesc_bf16(CE_BM, BWD_BN, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(CE_BM, H, 0, 1);     // B = Y tile read MN-major
      const uint64_t dp0 = umma_desc_kmajor(smem_u32(p_tile), 0);
      const uint64_t dyk0 = umma_desc_kmajor(smem_u32(y_tiles), 0);
      const uint64_t dym0 = umma_desc_mnmajor(smem_u32(y_tiles), 0, BWD_BN * 128);
      auto issue_s = [&](int i) {
        const int s = i % BWD_STAGES;
        if (lane == 0) TT_STAMP(0, i, 0);
        mbar_wait(&y_full[s], (i / BWD_STAGES) & 1);
        if (lane == 0) TT_STAMP(0, i, 1);
        mbar_wait(s_empty, (i & 1) ^ 1);
        if (lane == 0) TT_STAMP(0, i, 2);
        tc_fence_after();
        const uint64_t dy = dyk0 + (uint64_t)((s * y_bytes) >> 4);
        for (int kb = 0; kb < kq; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (elect_one())
              umma_bf16_ts(tmem_s, tmem_x + (uint32_t)(kb * 32 + k * 8),
                           dy + (uint64_t)(kb * (BWD_BN * 128 / 16) + k * 2), idesc_s, (kb | k) != 0);
          }
        if (elect_one()) umma_commit(s_full);
        __syncwarp();
        if (lane == 0) TT_STAMP(0, i, 3);
      };
      auto issue_o = [&](int i) {
        const int s = i % BWD_STAGES;
        if (lane == 0) TT_STAMP(0, i, 4);
        mbar_wait(p_full, i & 1);
        if (lane == 0) TT_STAMP(0, i, 5);
        tc_fence_after();
        const uint64_t dy = dym0 + (uint64_t)((s * y_bytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < BWD_BN / 16; ++kk) {         // K = 128 Y rows: 2 P k-blocks x 4 slices
          if (elect_one())
            umma_bf16(tmem_o, dp0 + (uint64_t)((kk >> 2) * (CE_BM * 128 / 16) + (kk & 3) * 2),
                      dy + (uint64_t)(kk * (2048 / 16)), idesc_o, (i | kk) != 0);
        }
        if (elect_one()) { umma_commit(p_empty); umma_commit(&y_empty[s]); }
        __syncwarp();
        if (lane == 0) TT_STAMP(0, i, 6);
      };
      mbar_wait(x_ready, 0);
      tc_fence_after();
      issue_s(0);
      for (int i = 1; i < nt; ++i) {
        issue_s(i);
        issue_o(i - 1);
      }
      issue_o(nt - 1);
      if (elect_one()) umma_commit(o_full);
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;   // half: which CW-column slice of every S tile (0 when EW == 4)
    const int col0 = half * CW;
    const int tid_e = threadIdx.x - 64;                    // 0 .. ETH-1
    const int lrow = quarter * 32 + lane;                  // row inside the tile == TMEM lane
    const int64_t row = x0 + lrow;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float c = p.inv_temp * kLog2e;
    const float row_lse = MODE == 2 ? p.mfix * kLog2e : (!COL && row < Bx) ? lse[row] * kLog2e : 0.f;
    // X tile: shared memory (TMA, 128B swizzle) -> registers -> TMEM (the TS-mode A operand of every S product)
    mbar_wait(x_bar, 0);
    for (int kb = half; kb < kq; kb += NH) {
      uint32_t xr[32];
      const uint8_t* xrow = x_tile + kb * (CE_BM * 128) + lrow * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 v = *reinterpret_cast<const uint4*>(xrow + ((ch ^ (lrow & 7)) << 4));
        xr[4 * ch] = v.x; xr[4 * ch + 1] = v.y; xr[4 * ch + 2] = v.z; xr[4 * ch + 3] = v.w;
      }
      tmem_st_x32(tmem_x + lane_addr + (uint32_t)(kb * 32), xr);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(x_ready);
    TT_CTA_STAMP(1);
    // Y tiles that can hold a positive of one of this CTA's rows (CTA-uniform band)
    const int64_t band_lo = COL ? x0 - label_offset : x0 + label_offset;
    auto load_col_lse = [&](int i) {                        // this thread's column of Y tile i (COL mode; first 128 threads)
      const int64_t gc = (int64_t)(t_beg + i) * BWD_BN + tid_e;
      return (tid_e < BWD_BN && i < nt && gc < By) ? __ldcg(lse + gc) * kLog2e : CUDART_INF_F;   // L2: the single-launch form reads what phase 1 of the SAME kernel wrote
    };
    float next_cl = COL ? load_col_lse(0) : 0.f;
    for (int i = 0; i < nt; ++i) {
      const int64_t y0 = (int64_t)tile_at(i) * BWD_BN;
      const float4* cl4 = reinterpret_cast<const float4*>(col_lse + (i & 1) * BWD_BN);
      if (COL) {                                            // column lse -> smem; read back as broadcasts
        if (tid_e < BWD_BN) col_lse[(i & 1) * BWD_BN + tid_e] = next_cl;
        asm volatile("bar.sync 1, %0;" ::"n"(ETH) : "memory");   // the epilogue warps only
        next_cl = load_col_lse(i + 1);                      // latency hides behind this tile's work
      }
      if (threadIdx.x == 64) TT_STAMP(1, i, 0);
      mbar_wait(s_full, i & 1);
      if (threadIdx.x == 64) TT_STAMP(1, i, 1);
      tc_fence_after();
      uint32_t r[HC][32];
#pragma unroll
      for (int h = 0; h < HC; ++h) tmem_ld_x32(tmem_s + lane_addr + (uint32_t)(col0 + 32 * h), r[h]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);                 // S(i+1) may overwrite the accumulator
      if (threadIdx.x == 64) TT_STAMP(1, i, 2);
      const bool in_band = y0 + BWD_BN > band_lo && y0 < band_lo + CE_BM;    // tile intersects the diagonal band
      if (MODE == 2 && in_band) {                           // this row's positive logit, before the exponentials overwrite S
        const int64_t pj = row + label_offset - y0 - col0;
        if (pj >= 0 && pj < CW && y0 + col0 + pj < By && row < Bx) {
          float pv = 0.f;
#pragma unroll
          for (int h = 0; h < HC; ++h)
#pragma unroll
            for (int j = 0; j < 32; ++j) pv = (32 * h + j == (int)pj) ? __uint_as_float(r[h][j]) : pv;
          pos_val = pv; pos_found = true;
        }
      }
      // P = exp2(S*c - lse*log2e) (ragged columns: lse = +inf in COL mode, masked below otherwise), in place
#pragma unroll
      for (int h = 0; h < HC; ++h)
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 lv = make_float4(row_lse, row_lse, row_lse, row_lse);
          if (COL) lv = cl4[(col0 + 32 * h + j) >> 2];
          r[h][j + 0] = __float_as_uint(exp2_sel<3>(fmaf(__uint_as_float(r[h][j + 0]), c, -lv.x), j + 0));
          r[h][j + 1] = __float_as_uint(exp2_sel<3>(fmaf(__uint_as_float(r[h][j + 1]), c, -lv.y), j + 1));
          r[h][j + 2] = __float_as_uint(exp2_sel<3>(fmaf(__uint_as_float(r[h][j + 2]), c, -lv.z), j + 2));
          r[h][j + 3] = __float_as_uint(exp2_sel<3>(fmaf(__uint_as_float(r[h][j + 3]), c, -lv.w), j + 3));
        }
      if (!COL && y0 + BWD_BN > By) {
#pragma unroll
        for (int h = 0; h < HC; ++h)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (y0 + col0 + 32 * h + j >= By) r[h][j] = 0u;
      }
      if (MODE == 2) {
        // unnormalised row sums.  The positive's own E is kept OUT of both the P tile and the sum: the tail re-derives it
        // from the saved logit and forms  dy = coef (O_off - L_off d_pos) / (L_off + E_pos)  -- the "P_ii - 1" factor is
        // then -L_off / L, free of cancellation when the softmax is peaked on the positive
        if (in_band) {
          const int64_t pj = row + label_offset - y0 - col0;
          if (pj >= 0 && pj < CW && y0 + col0 + pj < By) {
#pragma unroll
            for (int h = 0; h < HC; ++h)
#pragma unroll
              for (int j = 0; j < 32; ++j) r[h][j] = (32 * h + j == (int)pj) ? 0u : r[h][j];
          }
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int h = 0; h < HC; ++h)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            s0 += __uint_as_float(r[h][j]); s1 += __uint_as_float(r[h][j + 1]);
            s2 += __uint_as_float(r[h][j + 2]); s3 += __uint_as_float(r[h][j + 3]);
          }
        lsum += (s0 + s1) + (s2 + s3);
      } else if (in_band) {
        const int64_t pj = (COL ? row - label_offset : row + label_offset) - y0 - col0;   // inside this thread's slice?
        if (pj >= 0 && pj < CW && y0 + col0 + pj < By) {
#pragma unroll
          for (int h = 0; h < HC; ++h)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              r[h][j] = __float_as_uint(__uint_as_float(r[h][j]) - ((32 * h + j == (int)pj) ? 1.0f : 0.0f));
        }
      }
      if (threadIdx.x == 64) TT_STAMP(1, i, 3);
      mbar_wait(p_empty, (i & 1) ^ 1);                     // O(i-1) has finished reading the P tile
      if (threadIdx.x == 64) TT_STAMP(1, i, 4);
#pragma unroll
      for (int ch = 0; ch < CW / 8; ++ch) {                 // 16-byte chunks of this thread's slice: k-block gch/8, 128B swizzle inside
        const int h = ch >> 2, j0 = (ch & 3) * 8, gch = col0 / 8 + ch;
        const uint4 v = make_uint4(pack_bf16x2(__uint_as_float(r[h][j0 + 0]), __uint_as_float(r[h][j0 + 1])),
                                   pack_bf16x2(__uint_as_float(r[h][j0 + 2]), __uint_as_float(r[h][j0 + 3])),
                                   pack_bf16x2(__uint_as_float(r[h][j0 + 4]), __uint_as_float(r[h][j0 + 5])),
                                   pack_bf16x2(__uint_as_float(r[h][j0 + 6]), __uint_as_float(r[h][j0 + 7])));
        *reinterpret_cast<uint4*>(p_tile + (gch >> 3) * (CE_BM * 128) + lrow * 128 + (((gch & 7) ^ (lrow & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (threadIdx.x == 64) TT_STAMP(1, i, 5);
    }
    // final: O (TMEM) -> registers -> warp-private smem transpose -> 128-byte coalesced global stores
    const float scale = p.coef * (p.grad_out ? *p.grad_out : 1.0f);
    if (nt > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
    }
    TT_CTA_STAMP(2);
    float* T = reinterpret_cast<float*>(y_tiles + (warp - 2) * (32 * 36 * 4));      // [32][36] per warp; the Y stages are idle now
    const int64_t row0 = x0 + quarter * 32;
    const int nrows = (int)min((int64_t)32, Bx - row0);
    float* orow = out + row0 * H + lane;
    for (int cb = half; cb < (fused ? 0 : H / 32); cb += NH) {
      uint32_t q[32];
      if (nt > 0) {
        tmem_ld_x32(tmem_o + lane_addr + (uint32_t)(cb * 32), q);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) q[j] = 0u;
      }
      if (threadIdx.x == 64) TT_STAMP(1, 32 + cb, 0);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(&T[lane * 36 + j]) = make_float4(__uint_as_float(q[j]) * scale, __uint_as_float(q[j + 1]) * scale,
                                                                    __uint_as_float(q[j + 2]) * scale, __uint_as_float(q[j + 3]) * scale);
      __syncwarp();
      if (threadIdx.x == 64) TT_STAMP(1, 32 + cb, 1);
      float v[32];
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) v[rr] = T[rr * 36 + lane];
      if (threadIdx.x == 64) TT_STAMP(1, 32 + cb, 2);
#pragma unroll
      for (int rr = 0; rr < 32; ++rr)
        if (rr < nrows) orow[(int64_t)rr * H + cb * 32] = v[rr];
      __syncwarp();
      if (threadIdx.x == 64) TT_STAMP(1, 32 + cb, 3);
    }
    if (!fused) TT_CTA_STAMP(3);
    if (MODE == 2) {
      if (!pos_found) pos_val = -CUDART_INF_F;             // exactly one thread of the cluster holds each row's positive
      if (NH == 2) {                                       // the two column halves of a row, fixed order (col_lse is idle in this mode)
        float2* cmb = reinterpret_cast<float2*>(col_lse);
        if (half == 1) cmb[lrow] = make_float2(lsum, pos_val);
        asm volatile("bar.sync 1, %0;" ::"n"(ETH) : "memory");
        if (half == 0) { const float2 t = cmb[lrow]; lsum += t.x; pos_val = fmaxf(pos_val, t.y); cmb[lrow].x = lsum; }
        asm volatile("bar.sync 1, %0;" ::"n"(ETH) : "memory");
        if (half == 1) lsum = cmb[lrow].x;                   // both halves scale their accumulator columns by the same 1 / lsum
      }
    }
  }
  if (fused) {
    // ---- cluster tail.  The NS = gridDim.y CTAs of a row tile (one cluster; NS = 1, 2 or 4) each hold a partial
    // accumulator O_k (128 x H fp32) in tensor memory; rank r finishes rows [NFIN r, NFIN r + NFIN), NFIN = 128 / NS.
    //   A. every epilogue thread dumps its accumulator row into a local staging block (one block per TMEM lane quarter,
    //      16-byte chunks XOR-swizzled by row: conflict-free for the row-per-lane writes here and the column-per-lane
    //      reads below); quarters finished elsewhere then travel as 32-row bulk copies (cp.async.bulk shared::cta ->
    //      shared::cluster) that complete on the finishing CTA's mbarrier.  MODE 2 also sends every row's share of the
    //      softmax normaliser and the positive logit (st.async onto the same mbarrier).
    //   B. the eight warps walk the finished rows, lanes across columns: partials added in rank order (bitwise
    //      reproducible); MODE 2 forms dy = coef (O_off - L_off y_pos) / (L_off + E_pos) and lse / the loss partials;
    //      the normalise backward  dz = (dy - y (y . dy)) * inv_norm  (y = the X rows, re-read from L2) leaves as bf16
    //      rows plus per-32-row column sums (or, MODE 2 without dz, dy itself as fp32 rows).
    // This code runs ONCE per CTA, so it is bound by cold instruction fetch and by memory round trips, not by arithmetic:
    // it is specialised on NS (no run-time loops over ranks), keeps no release / acquire fences on its path (they compile
    // to gpu-scope MEMBARs), issues every global load before the exchange, and hands the loss ticket to the idle TMA warp.
    // The staging aliases the Y stages and the P tile: nobody may write into a peer before that peer has left its main
    // loop (cluster barrier 1, arrive early / wait late), and nobody may exit while a peer still reads from it (barrier 2).
    auto tail = [&](auto ns_tag) {
      constexpr int NS = decltype(ns_tag)::value;
      constexpr int NFIN = CE_BM / NS;                       // rows finished by this CTA (128, 64 or 32)
      constexpr int QPR = 4 / NS;                            // TMEM lane quarters (32-row blocks) per rank
      constexpr int RPW = NFIN / EW;                         // rows per epilogue warp (4, 8, 16; 32 with four warps)
      const uint32_t rank = NS > 1 ? cluster_ctarank() : 0u;
      const int fin0 = NFIN * (int)rank;                     // first tile row finished here
      const int nch = H / 32;                                // 32-column chunks per row
      // staging: [quarter][chunk] blocks of 32 rows x 32 columns fp16 (2 KB, 16-byte units XOR-swizzled by row pair)
      uint8_t* stage = y_tiles;                              // [4][nch] blocks: this CTA's partial
      uint8_t* recv = stage + 4 * nch * 2048;                // [NS - 1][QPR][nch] blocks: the other ranks' partials of MY rows
      uint8_t* cs_b = recv + (NS - 1) * QPR * nch * 2048;
      float* cs_s = reinterpret_cast<float*>(cs_b);          // [EW][H] column-sum partials
      uint8_t* ybuf = cs_b + EW * H * 4;                     // [NFIN][H] bf16: y = the X rows of the rows finished here
      uint8_t* pbuf = ybuf + NFIN * H * 2;                   // [NFIN][H] bf16: MODE 2, each row's positive row of Y
      float2* lp2 = reinterpret_cast<float2*>(col_lse);      // [NS][NFIN] MODE 2: (share of sum_j E_ij, positive dot product or -inf)
      const bool to_dz = p.dz[PASS] != nullptr;
      const int e = warp - 2;
      const int lc = lane * 4;                               // this lane's columns: lc .. lc + 3 and 128 + lc .. 128 + lc + 3
      const bool c0 = lc < H, c1 = 128 + lc < H;
      const int64_t gw0 = x0 + fin0 + e * RPW;               // first global row of this warp (epilogue warps)
      TT_TAIL(0);
      if (NS > 1) cluster_arrive_relaxed();                  // this CTA's tensor pipe has retired: its Y stages / P tile are free
      float inv_n = 0.f;                                     // lane i: 1 / |z| of this warp's row i
      if (warp >= 2) {
        // this warp's rows of y (and, MODE 2, of the positives) -> shared memory, asynchronously: in flight under the dump
        // and the exchange.  A rolled loop: this code runs once, every instruction of it is a cold fetch.
        {
          uint32_t pr = 0; int64_t pbase = 0;                // MODE 2: positive of row i = physical row pbase + pr
          const uint32_t blk = (uint32_t)p.y_blk[PASS];
          if (MODE == 2) {
            const uint32_t g = (uint32_t)(gw0 + label_offset);
            const uint32_t qb = g / blk;
            pr = g - qb * blk; pbase = (int64_t)qb * p.y_blk_stride[PASS] + p.y_blk_off[PASS];
          }
          const bool cl = lane * 8 < H;                      // 16 bytes = 8 bf16 per lane
#pragma unroll 1
          for (int i = 0; i < RPW; ++i) {
            const int64_t grow = gw0 + i;
            if (grow < Bx && cl) {
              if (to_dz || estore) cp_async_16(ybuf + (e * RPW + i) * H * 2 + lane * 16, p.xg[PASS] + grow * H + lane * 8);
              if (MODE == 2) cp_async_16(pbuf + (e * RPW + i) * H * 2 + lane * 16,
                                         (p.gate ? p.y_own + (grow + label_offset - (int64_t)p.gate_rank * p.gate_blk) * H
                                                 : p.yg[PASS] + (pbase + pr) * H) + lane * 8);
              if (MODE == 3) {                               // the query whose positive this document is (if any)
                const int64_t gi = grow - label_offset;
                if (gi >= 0 && gi < By)
                  cp_async_16(pbuf + (e * RPW + i) * H * 2 + lane * 16,
                              p.yg[PASS] + ((gi / blk) * p.y_blk_stride[PASS] + (gi % blk) + p.y_blk_off[PASS]) * H + lane * 8);
              }
            }
            if (MODE == 2 && ++pr == blk) { pr = 0; pbase += p.y_blk_stride[PASS]; }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (to_dz && lane < RPW && gw0 + lane < Bx) inv_n = __ldg(p.inv_norm[PASS] + gw0 + lane);
        // A. accumulator row -> fp16 (MODE 2: scaled by 1 / lsum so that every row is a convex combination of unit rows --
        // the raw E-weighted sums can be far below the fp16 normal range), chunk by chunk; a chunk that another rank
        // finishes leaves as a 2 KB bulk copy as soon as this warp has written it (the exchange runs under the dump)
        const int quarter = warp & 3, half = e >> 2;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const int dst = quarter / QPR;                       // finishing rank of this warp's 32 rows (warp-uniform)
        const int ql = quarter - dst * QPR;
        const bool remote = NS > 1 && dst != (int)rank;
        const int slot = (int)rank < dst ? (int)rank : (int)rank - 1;
        const uint32_t peer_bar = remote ? mapa_shared(smem_u32(xfer_bar), (uint32_t)dst) : 0u;
        const float rs = MODE == 2 ? (lsum > 0.f ? 1.0f / lsum : 0.f) : 1.0f;
        const int sx = (lane >> 1) & 3;
        bool waited = NS == 1;
        for (int cb = half; cb < nch; cb += NH) {            // the warps of a quarter share its 32 rows, alternating column chunks
          uint32_t q[32];
          if (nt > 0) {
            tmem_ld_x32(tmem_o + lane_addr + (uint32_t)(cb * 32), q);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) q[j] = 0u;
          }
          uint8_t* blkp = stage + (quarter * nch + cb) * 2048;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 v;
            v.x = pack_f16x2(__uint_as_float(q[8 * u + 0]) * rs, __uint_as_float(q[8 * u + 1]) * rs);
            v.y = pack_f16x2(__uint_as_float(q[8 * u + 2]) * rs, __uint_as_float(q[8 * u + 3]) * rs);
            v.z = pack_f16x2(__uint_as_float(q[8 * u + 4]) * rs, __uint_as_float(q[8 * u + 5]) * rs);
            v.w = pack_f16x2(__uint_as_float(q[8 * u + 6]) * rs, __uint_as_float(q[8 * u + 7]) * rs);
            *reinterpret_cast<uint4*>(blkp + lane * 64 + ((u ^ sx) << 4)) = v;
          }
          if (!waited) { TT_TAIL(1); cluster_wait_noacq(); waited = true; TT_TAIL(2); }   // every CTA of the cluster has left its main loop
          if (remote) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0)
              dsmem_bulk_copy(mapa_shared(smem_u32(recv + ((slot * QPR + ql) * nch + cb) * 2048), (uint32_t)dst), blkp, 2048u, peer_bar);
          }
        }
        if (!waited) cluster_wait_noacq();
        TT_CTA_STAMP(4);
        if (MODE == 2 && half == 0) {
          const int rloc = ql * 32 + lane;
          if (remote) st_async_f2(mapa_shared(smem_u32(lp2 + (int)rank * NFIN + rloc), (uint32_t)dst), lsum, pos_val, peer_bar);
          else lp2[(int)rank * NFIN + rloc] = make_float2(lsum, pos_val);
        }
        TT_TAIL(3);
        asm volatile("bar.sync 1, %0;" ::"n"(ETH) : "memory");   // this CTA's own rows (and scalars) are staged
        if (NS > 1) mbar_wait(xfer_bar, 0);                  // the peers' rows and scalars have landed
        TT_TAIL(4);
        TT_CTA_STAMP(6);
        // lane i <-> this warp's row i: per-rank weights of the fp16 partials; MODE 2: softmax normaliser, lse, loss partials
        float wk[NS], wpos = 0.f, inv_l = 0.f;
#pragma unroll
        for (int k = 0; k < NS; ++k) wk[k] = 1.0f;
        if (MODE == 3 && lane < RPW && gw0 + lane < Bx) {     // 1 - P_pos of the query this document is the positive of
          const int64_t gi = gw0 + lane - label_offset;
          if (gi >= 0 && gi < By) wpos = __ldcg(p.wpos_in + gi);
        }
        if (MODE == 2) {
          float dl = 0.f, dpz = 0.f;
          if (lane < RPW && gw0 + lane < Bx) {
            const int r = e * RPW + lane;
            float Loff = 0.f, pdot = -CUDART_INF_F;          // sum of the negatives' E (rank order), the positive's dot product
#pragma unroll
            for (int k = 0; k < NS; ++k) { const float2 t = lp2[k * NFIN + r]; wk[k] = t.x; Loff += t.x; pdot = fmaxf(pdot, t.y); }
            const float epos = fast_exp2(fmaf(pdot, p.inv_temp * kLog2e, -p.mfix * kLog2e));   // the positive's E, as the main loop formed it
            const float L = Loff + epos;
            const float invL = 1.0f / L;
#pragma unroll
            for (int k = 0; k < NS; ++k) wk[k] *= invL;      // O / L = sum_k (lsum_k / L) (O_k / lsum_k)
            wpos = Loff * invL;                              // 1 - P_pos, free of cancellation
            inv_l = invL;
            if (estore) p.wpos_out[gw0 + lane] = wpos;
            const float lse_r = p.mfix + logf(L), pl = pdot * p.inv_temp;
            p.lse_out[gw0 + lane] = lse_r;
            dl = lse_r - pl; dpz = pl;
          }
          dl = warp_sum(dl); dpz = warp_sum(dpz);
          if (lane == 0) { red_s[2 * e] = dl; red_s[2 * e + 1] = dpz; }
          asm volatile("bar.arrive 6, %0;" ::"n"(ETH + 32) : "memory");   // hand the partials to warp 0 (loss ticket)
        }
        const float scale = p.coef * (p.grad_out ? *p.grad_out : 1.0f);
        float4 cs0 = make_float4(0.f, 0.f, 0.f, 0.f), cs1 = cs0;
        // per-lane addressing, hoisted: a warp's RPW rows lie inside ONE 32-row block, so the NS partial blocks are fixed
        const int offA = c0 ? ((lane >> 3) * 2048) + ((lane & 1) << 3) : 0, offB = c1 ? ((4 + (lane >> 3)) * 2048) + ((lane & 1) << 3) : 0;
        const int ux = (lane & 7) >> 1, qrow = (e * RPW) >> 5;
        const uint8_t* srck[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k)
          srck[k] = k == (int)rank ? stage + ((int)rank * QPR + qrow) * nch * 2048 : recv + ((k < (int)rank ? k : k - 1) * QPR + qrow) * nch * 2048;
        const int yoA = c0 ? lane * 8 : 0, yoB = c1 ? 256 + lane * 8 : 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();                                        // a warp reads back only the rows it copied itself
#pragma unroll 1
        for (int i = 0; i < RPW; ++i) {                      // ONE row per iteration of a rolled loop (cold code: keep it small)
          const int r = e * RPW + i;                         // row among the NFIN rows finished here
          const int64_t grow = gw0 + i;
          const int rq = r & 31;
          const int rowoff = rq * 64 + ((ux ^ ((rq >> 1) & 3)) << 4);
          float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
#pragma unroll
          for (int k = 0; k < NS; ++k) {                     // rank order
            const uint2 h0 = *reinterpret_cast<const uint2*>(srck[k] + rowoff + offA), h1 = *reinterpret_cast<const uint2*>(srck[k] + rowoff + offB);
            const float w = MODE == 2 ? __shfl_sync(0xffffffffu, wk[k], i) : 1.0f;
            const float2 a0 = unpack_f16x2(h0.x), a1 = unpack_f16x2(h0.y), b0 = unpack_f16x2(h1.x), b1 = unpack_f16x2(h1.y);
            o0.x = fmaf(a0.x, w, o0.x); o0.y = fmaf(a0.y, w, o0.y); o0.z = fmaf(a1.x, w, o0.z); o0.w = fmaf(a1.y, w, o0.w);
            o1.x = fmaf(b0.x, w, o1.x); o1.y = fmaf(b0.y, w, o1.y); o1.z = fmaf(b1.x, w, o1.z); o1.w = fmaf(b1.y, w, o1.w);
          }
          const float wp = (MODE >= 2) ? -__shfl_sync(0xffffffffu, wpos, i) : 0.f;
          if (MODE == 2 || (MODE == 3 && wp != 0.f)) {       // MODE 3: rows without a positive have nothing staged in pbuf
            const uint2 pa = *reinterpret_cast<const uint2*>(pbuf + r * H * 2 + yoA), pb = *reinterpret_cast<const uint2*>(pbuf + r * H * 2 + yoB);
            o0.x = fmaf(wp, bf16lo_to_f32(pa.x), o0.x); o0.y = fmaf(wp, bf16hi_to_f32(pa.x), o0.y);
            o0.z = fmaf(wp, bf16lo_to_f32(pa.y), o0.z); o0.w = fmaf(wp, bf16hi_to_f32(pa.y), o0.w);
            if (c1) {
              o1.x = fmaf(wp, bf16lo_to_f32(pb.x), o1.x); o1.y = fmaf(wp, bf16hi_to_f32(pb.x), o1.y);
              o1.z = fmaf(wp, bf16lo_to_f32(pb.y), o1.z); o1.w = fmaf(wp, bf16hi_to_f32(pb.y), o1.w);
            }
          }
          const bool valid = grow < Bx;                      // warp-uniform
          if (MODE == 2 && estore && valid) {                // x_i / L_i: the B operand of the stored-E document gradient
            const uint2 ya = *reinterpret_cast<const uint2*>(ybuf + r * H * 2 + yoA), yb = *reinterpret_cast<const uint2*>(ybuf + r * H * 2 + yoB);
            const float il = __shfl_sync(0xffffffffu, inv_l, i);
            if (c0) *reinterpret_cast<uint2*>(p.xs_out + grow * H + lc) =
                make_uint2(pack_bf16x2(bf16lo_to_f32(ya.x) * il, bf16hi_to_f32(ya.x) * il), pack_bf16x2(bf16lo_to_f32(ya.y) * il, bf16hi_to_f32(ya.y) * il));
            if (c1) *reinterpret_cast<uint2*>(p.xs_out + grow * H + 128 + lc) =
                make_uint2(pack_bf16x2(bf16lo_to_f32(yb.x) * il, bf16hi_to_f32(yb.x) * il), pack_bf16x2(bf16lo_to_f32(yb.y) * il, bf16hi_to_f32(yb.y) * il));
          }
          if (to_dz) {
            const uint2 ya = *reinterpret_cast<const uint2*>(ybuf + r * H * 2 + yoA), yb = *reinterpret_cast<const uint2*>(ybuf + r * H * 2 + yoB);
            const float y0 = c0 ? bf16lo_to_f32(ya.x) : 0.f, y1 = c0 ? bf16hi_to_f32(ya.x) : 0.f, y2 = c0 ? bf16lo_to_f32(ya.y) : 0.f, y3 = c0 ? bf16hi_to_f32(ya.y) : 0.f;
            const float y4 = c1 ? bf16lo_to_f32(yb.x) : 0.f, y5 = c1 ? bf16hi_to_f32(yb.x) : 0.f, y6 = c1 ? bf16lo_to_f32(yb.y) : 0.f, y7 = c1 ? bf16hi_to_f32(yb.y) : 0.f;
            float dot = o0.x * y0;
            dot = fmaf(o0.y, y1, dot); dot = fmaf(o0.z, y2, dot); dot = fmaf(o0.w, y3, dot);
            dot = fmaf(o1.x, y4, dot); dot = fmaf(o1.y, y5, dot); dot = fmaf(o1.z, y6, dot); dot = fmaf(o1.w, y7, dot);
            dot = warp_sum(dot);
            const float si = scale * __shfl_sync(0xffffffffu, inv_n, i);
            const float4 d0 = make_float4((o0.x - y0 * dot) * si, (o0.y - y1 * dot) * si, (o0.z - y2 * dot) * si, (o0.w - y3 * dot) * si);
            const float4 d1 = make_float4((o1.x - y4 * dot) * si, (o1.y - y5 * dot) * si, (o1.z - y6 * dot) * si, (o1.w - y7 * dot) * si);
            if (valid && c0) {
              cs0.x += d0.x; cs0.y += d0.y; cs0.z += d0.z; cs0.w += d0.w;
              *reinterpret_cast<uint2*>(p.dz[PASS] + grow * H + lc) = make_uint2(pack_bf16x2(d0.x, d0.y), pack_bf16x2(d0.z, d0.w));
            }
            if (valid && c1) {
              cs1.x += d1.x; cs1.y += d1.y; cs1.z += d1.z; cs1.w += d1.w;
              *reinterpret_cast<uint2*>(p.dz[PASS] + grow * H + 128 + lc) = make_uint2(pack_bf16x2(d1.x, d1.y), pack_bf16x2(d1.z, d1.w));
            }
          } else {                                           // MODE 2 without the normalise backward: dy rows as fp32
            if (valid && c0) *reinterpret_cast<float4*>(p.out[PASS] + grow * H + lc) = make_float4(o0.x * scale, o0.y * scale, o0.z * scale, o0.w * scale);
            if (valid && c1) *reinterpret_cast<float4*>(p.out[PASS] + grow * H + 128 + lc) = make_float4(o1.x * scale, o1.y * scale, o1.z * scale, o1.w * scale);
          }
        }
        TT_TAIL(7);
        if (to_dz) {
          // per-32-row column sums: the 2 / 4 / 8 warps that share a 32-row block add their partials in warp order
          constexpr int WPB = 32 / RPW > 0 ? 32 / RPW : 1;   // warps per 32-row block
          if (WPB > 1) {
            if (c0) *reinterpret_cast<float4*>(cs_s + e * H + lc) = cs0;
            if (c1) *reinterpret_cast<float4*>(cs_s + e * H + 128 + lc) = cs1;
            asm volatile("bar.sync 1, %0;" ::"n"(ETH) : "memory");
          }
          TT_TAIL(8);
          const int64_t blk_row = x0 + fin0 + (e / WPB) * 32;
          if (e % WPB == 0 && blk_row < Bx) {
#pragma unroll
            for (int w2 = 1; w2 < WPB; ++w2) {               // fixed order: this warp's rows, then the following warps'
              const float* o = cs_s + (e + w2) * H;
              const float4 t0 = *reinterpret_cast<const float4*>(o + (c0 ? lc : 0)), t1 = *reinterpret_cast<const float4*>(o + (c1 ? 128 + lc : 0));
              cs0.x += t0.x; cs0.y += t0.y; cs0.z += t0.z; cs0.w += t0.w;
              cs1.x += t1.x; cs1.y += t1.y; cs1.z += t1.z; cs1.w += t1.w;
            }
            if (c0) *reinterpret_cast<float4*>(p.dz_colsum[PASS] + (blk_row >> 5) * H + lc) = cs0;
            if (c1) *reinterpret_cast<float4*>(p.dz_colsum[PASS] + (blk_row >> 5) * H + 128 + lc) = cs1;
          }
        }
        TT_TAIL(9);
      } else {
        if (NS > 1) cluster_wait_noacq();
        if (MODE == 2 && warp == 0) {
          // loss (the TMA warp is idle by now): this CTA's partial -> its slot; the CTA that takes the last ticket adds all
          // slots in slot order.  One acq_rel atomic is the only gpu-scope fence of the kernel, and it is off the epilogue
          // warps' path.
          asm volatile("bar.sync 6, %0;" ::"n"(ETH + 32) : "memory");
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int k = 0; k < EW; ++k) { a += red_s[2 * k]; b += red_s[2 * k + 1]; }
          const unsigned nslots = gridDim.x * (unsigned)NS;
          unsigned ticket = 0;
          if (lane == 0) {
            const unsigned slot = blockIdx.x * (unsigned)NS + rank;
            __stcg(p.tile_sums + 2 * slot, a); __stcg(p.tile_sums + 2 * slot + 1, b);
            ticket = atom_add_acq_rel_gpu(p.counter, 1u);
          }
          ticket = __shfl_sync(0xffffffffu, ticket, 0);
          if (ticket == nslots - 1) {
            float tl = 0.f, tp = 0.f;
            for (unsigned t = lane; t < nslots; t += 32) { tl += __ldcg(p.tile_sums + 2 * t); tp += __ldcg(p.tile_sums + 2 * t + 1); }
            tl = warp_sum(tl); tp = warp_sum(tp);
            if (lane == 0) {
              *p.loss = tl * p.loss_scale;
              if (p.pos_mean) *p.pos_mean = tp / (p.inv_temp * (float)Bx);
              *p.counter = 0u;                                   // re-armed for the next launch
            }
          }
        }
      }
      if (NS > 1) { cluster_arrive_relaxed(); cluster_wait_noacq(); }   // no CTA leaves while a peer's bulk copy may still read its staging
      TT_TAIL(10);
      TT_CTA_STAMP(3);
    };
    if (gridDim.y == 4) tail(std::integral_constant<int, 4>{});
    else if (gridDim.y == 2) tail(std::integral_constant<int, 2>{});
    else tail(std::integral_constant<int, 1>{});
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1 && tmem_mode != 1) tmem_dealloc(tmem_base, 512);
  if (threadIdx.x == 0) {                                  // the single-launch form runs a second body over the same barrier words
    for (uint64_t* b = x_bar; b <= sc_bar; ++b) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
  }
  __syncthreads();
}

template <int EW>
__global__ void __launch_bounds__(64 + EW * 32, 1)
tc_ce_bwd_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmY0,
                 const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmY1, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const int pass = blockIdx.z + p.pass_base;
  if ((int64_t)blockIdx.x * CE_BM >= p.Bx[pass] || (p.out[pass] == nullptr && p.dz[pass] == nullptr)) return;   // cluster-uniform: nothing to do for this pass
  if (pass == 0) ce_bwd_body<0, EW>(&tmX0, &tmY0, p, base);
  else           ce_bwd_body<1, EW>(&tmX1, &tmY1, p, base);
}

// Both launches of the one-pass loss as ONE (square single-GPU case): phase 1 = forward + query gradient, a grid-wide barrier
// (the document gradient needs the lse of every query), phase 2 = document gradient.  The grid is at most one CTA per SM and
// every CTA is resident before the barrier can be reached (checked on the host with cudaOccupancyMaxActiveClusters), so
// the barrier cannot deadlock; it is a self-resetting sense-reversal barrier in the call site's sync scratch
// (word 1: arrivals, word 2: generation).  Saves the kernel boundary between the two launches (~7 us with 224 KB of
// shared memory and all of tensor memory per CTA: nothing of the second launch can start before the first has left).
__global__ void __launch_bounds__(64 + 8 * 32, 1)
tc_ce_onepass_kernel(const __grid_constant__ CUtensorMap tmXq, const __grid_constant__ CUtensorMap tmYq,
                     const __grid_constant__ CUtensorMap tmXd, const __grid_constant__ CUtensorMap tmYd, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  unsigned gen = 0;
  if (threadIdx.x == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(p.counter + 2) : "memory");
  ce_bwd_body<2, 8>(&tmXq, &tmYq, p, base, &tmYq, 1);
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    __threadfence();                                       // this CTA's lse / dz rows before its arrival
    if (atomicAdd(p.counter + 1, 1u) == total - 1) {
      p.counter[1] = 0u;
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.counter + 2), "r"(gen + 1) : "memory");
    } else {
      const long long t0 = clock64();
      for (unsigned spins = 1;; ++spins) {
        unsigned g2;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g2) : "l"(p.counter + 2) : "memory");
        if (g2 != gen) break;
        if ((spins & 1023u) == 0 && clock64() - t0 > 4000000000ll) {
          printf("tt_b200: grid barrier of the one-pass loss timed out (block %d,%d)\n", blockIdx.x, blockIdx.y);
          __trap();
        }
      }
    }
  }
  __syncthreads();
  ce_bwd_body<1, 8>(&tmXd, &tmYd, p, base, nullptr, 2);
}

// forward + query gradient in one pass (MODE 2): grid = (row tiles, splits), one cluster per row tile
__global__ void __launch_bounds__(64 + 8 * 32, 1)
tc_ce_fwd_dq_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                    const __grid_constant__ CUtensorMap tmYown, const __grid_constant__ CUtensorMap tmE, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  ce_bwd_body<2, 8>(&tmX, &tmY, p, base, &tmYown, 0, &tmE);
}

// document gradient from the stored E tiles (MODE 3): grid = (document row tiles, splits over the queries), one cluster per
// row tile; tmXs = X / L [queries, H] (64-row boxes), tmE = E [queries, documents] (64-row boxes)
__global__ void __launch_bounds__(64 + 8 * 32, 1)
tc_ce_dd_stored_kernel(const __grid_constant__ CUtensorMap tmXs, const __grid_constant__ CUtensorMap tmE, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  ce_bwd_body<3, 8>(&tmXs, &tmXs, p, base, nullptr, 0, &tmE);
}

static size_t fwd_smem(int H) {
  const size_t need = 1024 + FWD_STAGES * (size_t)FWD_BN * H * 2 + 20 * 8 + 16 + 128 * 2 * 4;
  // every CTA allocates all 512 TMEM columns: two CTAs on one SM would serialise on the allocator, so small H still
  // asks for more than half an SM's shared memory
  return need > 120 * 1024 ? need : 120 * 1024;
}
static size_t bwd_smem(int H) {
  const size_t need = 1024 + BWD_STAGES * (size_t)BWD_BN * H * 2 + (size_t)CE_BM * BWD_BN * 2 + 26 * 8 + 16 + 2 * BWD_BN * 4 + 64;
  return need > 120 * 1024 ? need : 120 * 1024;          // one CTA per SM (every CTA allocates all 512 TMEM columns)
}

static int pick_split(int64_t xtiles, int64_t By, int bn = CE_BN) {
  const int64_t yt = ceil_div(By, bn);
  int64_t s = kNumSMs / (xtiles > 0 ? xtiles : 1);                     // one wave
  if (s > yt) s = yt;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  const int64_t per = ceil_div(yt, s);
  return (int)ceil_div(yt, per);
}

}  // namespace tc

static bool tc_ce_supported(int H) { return H % 64 == 0 && H >= 64 && H <= 256; }

// splits: forward uses its own; backward uses ONE split count for both passes (they share a launch)
static int fwd_splits(int64_t Bq, int64_t Bd) { return tc::pick_split(ceil_div(Bq, tc::CE_BM), Bd, tc::FWD_BN); }
static int bwd_splits2(int64_t Bx0, int64_t By0, int64_t Bx1, int64_t By1) {
  const int64_t xt = ceil_div(Bx0, tc::CE_BM) + ceil_div(Bx1, tc::CE_BM);
  const int a = tc::pick_split(xt, By0, tc::BWD_BN), b = tc::pick_split(xt, By1, tc::BWD_BN);
  return a < b ? a : b;
}
static int bwd_splits(int64_t Bq, int64_t Bd) { return bwd_splits2(Bq, Bd, Bd, Bq); }
// cluster tail: the splits of a row tile form one cluster that finishes 128 / ns rows each -> ns in {1, 2, 4}
static int cluster_splits(int64_t xtiles, int64_t By) {
  const int64_t yt = ceil_div(By, tc::BWD_BN);
  int ns = 1;
  while (ns < 4 && xtiles * (ns * 2) <= kNumSMs && ns * 2 <= yt) ns *= 2;
  return ns;
}

struct TcCePlan { int ns_f, ns_b; size_t qb, db, ml, pos, partial, total; };
static TcCePlan plan_tc_ce(int64_t Bq, int64_t Bd, int H) {
  TcCePlan p{};
  p.ns_f = fwd_splits(Bq, Bd);
  p.ns_b = bwd_splits(Bq, Bd);
  p.qb = align_up((size_t)Bq * H * 2);
  p.db = align_up((size_t)Bd * H * 2);
  p.ml = align_up((size_t)p.ns_f * Bq * 2 * 4);
  p.pos = align_up((size_t)Bq * 4);
  p.partial = align_up(p.ns_b > 1 ? (size_t)p.ns_b * (Bq + Bd) * H * 4 : 0);
  p.total = p.qb + p.db + p.ml + p.pos + p.partial + 1024;
  return p;
}

size_t tc_inbatch_workspace(int64_t Bq, int64_t Bd, int H) {
  size_t f = inbatch_ce_fp32_workspace(Bq, Bd, H);
  if (!tc_ce_supported(H)) return f;
  size_t t = plan_tc_ce(Bq, Bd, H).total;
  return t > f ? t : f;
}

namespace tc {
int cast3_public(const float* a, __nv_bfloat16* ab, int64_t na, const float* b, __nv_bfloat16* bb, int64_t nb, cudaStream_t s);
}

// forward on bf16 operands; D may be a block-interleaved (all-gathered) buffer: logical row g lives at physical
// row (g / d_blk) * d_blk_stride + g % d_blk + d_blk_off of a buffer with d_buf_rows rows.
int tc_inbatch_fwd_ex(const __nv_bfloat16* qa, int64_t Bq, const __nv_bfloat16* da, int64_t Bd, int64_t d_buf_rows,
                      int64_t d_blk, int64_t d_blk_stride, int64_t d_blk_off, int H, float inv_temp, int64_t label_offset,
                      float loss_scale, float* loss, float* lse, float* pos_mean, float* part_ml, float* pos, void* sync_scratch,
                      cudaStream_t s) {
  CUtensorMap tmQ, tmD;
  int rc = tc::make_tmap_bf16(&tmQ, qa, (uint64_t)Bq, (uint64_t)H, tc::CE_BM); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmD, da, (uint64_t)d_buf_rows, (uint64_t)H, tc::FWD_BN); if (rc) return rc;
  const int ns = fwd_splits(Bq, Bd);
  const int yt = (int)ceil_div(Bd, tc::FWD_BN);
  const int per = (int)ceil_div(yt, ns);
  const size_t smem = tc::fwd_smem(H);
  TT_CUDA(cudaFuncSetAttribute(tc::tc_ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(Bq, tc::CE_BM), (unsigned)ns);
  tc::FwdFinalize fin{};
  if (sync_scratch) {
    fin.counters = static_cast<unsigned*>(sync_scratch);
    fin.tile_sums = reinterpret_cast<float*>(static_cast<char*>(sync_scratch) + align_up((size_t)(grid.x + 1) * 4, 16));
    fin.lse = lse; fin.loss = loss; fin.pos_mean = pos_mean; fin.loss_scale = loss_scale;
  }
  static const bool dbg_on = getenv("TT_CE_DEBUG") != nullptr;
  long long* dbg_dev = nullptr;
  const size_t ncta = (size_t)grid.x * grid.y;
  const size_t fdbg_n = ncta * 4 + 2 * 64 * 8;
  if (dbg_on) { cudaMalloc(&dbg_dev, fdbg_n * sizeof(long long)); cudaMemset(dbg_dev, 0, fdbg_n * sizeof(long long)); fin.dbg = dbg_dev; }
  TT_CUDA(launch_kernel(tc::tc_ce_fwd_kernel, grid, dim3(tc::FWD_THREADS), smem, s, true, tmQ, tmD, Bq, Bd, H, inv_temp, label_offset, per,
                        d_blk, d_blk_stride, d_blk_off, part_ml, pos, fin));
  TT_LAUNCH_CHECK("tc_ce_fwd_kernel");
  if (dbg_on) {
    std::vector<long long> h(fdbg_n);
    cudaStreamSynchronize(s);
    cudaMemcpy(h.data(), dbg_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    long long g0 = 0;
    for (size_t i = 0; i < ncta; ++i) if (h[4 * i] && (!g0 || h[4 * i] < g0)) g0 = h[4 * i];
    printf("[tt ce_fwd per-CTA ns] grid %u x %u, %d tiles per CTA: start q_ready loop_done end\n", grid.x, grid.y, per);
    for (size_t i = 0; i < ncta; i += 9) printf("  cta %3zu: %6lld %6lld %6lld %6lld\n", i, h[4 * i] - g0, h[4 * i + 1] - g0, h[4 * i + 2] - g0, h[4 * i + 3] - g0);
    const long long* t = h.data() + ncta * 4;
    const long long t0 = t[0];
    printf("[tt ce_fwd timeline, cycles] tile: MMA{begin,d_full,s_empty,issued} EPI{begin,s_full,loaded,maxed,done}\n");
    for (int i = 0; i < (per < 12 ? per : 12); ++i) {
      printf("  %2d: MMA", i);
      for (int k = 0; k < 4; ++k) printf(" %6lld", t[(0 * 64 + i) * 8 + k] - t0);
      printf("   EPI");
      for (int k = 0; k < 5; ++k) printf(" %6lld", t[(1 * 64 + i) * 8 + k] - t0);
      printf("\n");
    }
  }
  if (sync_scratch) return TT_OK;
  return inbatch_finalize(part_ml, pos, ns, Bq, inv_temp, loss_scale, lse, loss, pos_mean, nullptr, s);
}

size_t tc_inbatch_fwd_sync_bytes(int64_t Bq) {
  const size_t tiles = (size_t)ceil_div(Bq, tc::CE_BM);
  return align_up((tiles + 1) * 4, 16) + tiles * 2 * 4;
}

int tc_inbatch_fwd(const float* q, const float* d, const __nv_bfloat16* q_bf16, const __nv_bfloat16* d_bf16, int64_t Bq,
                   int64_t Bd, int H, float inv_temp, int64_t label_offset, float loss_scale, float* loss, float* lse,
                   float* pos_mean, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!tc_ce_supported(H))      // shapes the tensor-core kernel does not tile: same library, fp32 FFMA kernels
    return inbatch_ce_fwd_fp32(q, d, Bq, Bd, H, inv_temp, label_offset, loss_scale, loss, lse, pos_mean, ws, ws_bytes, s);
  const TcCePlan plan = plan_tc_ce(Bq, Bd, H);
  if (ws == nullptr || ws_bytes < plan.total) { set_error("tc_inbatch_fwd: workspace too small (%zu < %zu)", ws_bytes, plan.total); return TT_ERR_WORKSPACE; }
  Workspace w(ws, ws_bytes);
  __nv_bfloat16* qb = w.take<__nv_bfloat16>((size_t)Bq * H);
  __nv_bfloat16* db = w.take<__nv_bfloat16>((size_t)Bd * H);
  float* part_ml = w.take<float>(plan.ml / 4);
  float* pos = w.take<float>(Bq);
  if (!q_bf16 || !d_bf16) {
    int rc = tc::cast3_public(q_bf16 ? nullptr : q, qb, q_bf16 ? 0 : Bq * H, d_bf16 ? nullptr : d, db, d_bf16 ? 0 : Bd * H, s);
    if (rc) return rc;
  }
  return tc_inbatch_fwd_ex(q_bf16 ? q_bf16 : qb, Bq, d_bf16 ? d_bf16 : db, Bd, Bd, Bd > 0 ? Bd : 1, 0, 0, H, inv_temp,
                           label_offset, loss_scale, loss, lse, pos_mean, part_ml, pos, nullptr, s);
}

size_t tc_inbatch_fwd_ex_workspace(int64_t Bq, int64_t Bd) {
  return align_up((size_t)fwd_splits(Bq, Bd) * Bq * 2 * 4) + align_up((size_t)Bq * 4) + 256;
}

// ---- backward: both gradient passes in ONE launch ---------------------------------------------------------------
struct CePass {                  // rows X receive gradients from all (logical) rows of Y
  const __nv_bfloat16* x; int64_t Bx;
  const __nv_bfloat16* y; int64_t By, y_buf_rows, y_blk, y_blk_stride, y_blk_off;
  const float* lse; int64_t label_offset;
  float* out; int64_t part_stride;
  __nv_bfloat16* dz; float* dz_colsum; const float* inv_norm;     // fused normalise backward (all or none)
};

static int launch_tc_bwd(const CePass& pq, const CePass& pd, int H, float inv_temp, int nsplit, const float* grad_out,
                         float coef, cudaStream_t s) {
  CUtensorMap tmX0, tmY0, tmX1, tmY1;
  int rc = tc::make_tmap_bf16(&tmX0, pq.x, (uint64_t)pq.Bx, (uint64_t)H, tc::CE_BM); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmY0, pq.y, (uint64_t)pq.y_buf_rows, (uint64_t)H, tc::BWD_BN); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmX1, pd.x, (uint64_t)pd.Bx, (uint64_t)H, tc::CE_BM); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmY1, pd.y, (uint64_t)pd.y_buf_rows, (uint64_t)H, tc::BWD_BN); if (rc) return rc;
  tc::BwdParams p{};
  const CePass* ps[2] = {&pq, &pd};
  for (int k = 0; k < 2; ++k) {
    p.lse[k] = ps[k]->lse; p.Bx[k] = ps[k]->Bx; p.By[k] = ps[k]->By; p.label_offset[k] = ps[k]->label_offset;
    p.y_blk[k] = ps[k]->y_blk > 0 ? ps[k]->y_blk : 1; p.y_blk_stride[k] = ps[k]->y_blk_stride; p.y_blk_off[k] = ps[k]->y_blk_off;
    p.tiles_per_split[k] = (int)ceil_div(ceil_div(ps[k]->By, tc::BWD_BN), nsplit);
    p.out[k] = ps[k]->out; p.part_stride[k] = ps[k]->part_stride;
    p.dz[k] = ps[k]->dz; p.dz_colsum[k] = ps[k]->dz_colsum; p.inv_norm[k] = ps[k]->inv_norm; p.xg[k] = ps[k]->x; p.yg[k] = ps[k]->y;
  }
  const bool fused = pq.dz != nullptr || pd.dz != nullptr;
  if (fused) {
    if (nsplit != 1 && nsplit != 2 && nsplit != 4) { set_error("tc_inbatch_bwd: the fused normalise backward needs 1, 2 or 4 splits (got %d)", nsplit); return TT_ERR_UNSUPPORTED; }
    for (int k = 0; k < 2; ++k) {
      const bool any = ps[k]->dz || ps[k]->out;
      if (any && (!ps[k]->dz || !ps[k]->dz_colsum || !ps[k]->inv_norm)) { set_error("tc_inbatch_bwd: fused passes need dz, dz_colsum and inv_norm"); return TT_ERR_INVALID; }
      p.out[k] = nullptr;
    }
  }
  p.H = H; p.inv_temp = inv_temp; p.grad_out = grad_out; p.coef = coef;
  const size_t smem = tc::bwd_smem(H);
  // epilogue warps: 8 by default (two per TMEM lane quarter); TT_CE_EW=4 selects the one-warp-per-quarter variant
  static const int ew = [] { const char* e = getenv("TT_CE_EW"); return (e && atoi(e) == 4) ? 4 : 8; }();
  TT_CUDA(cudaFuncSetAttribute(tc::tc_ce_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TT_CUDA(cudaFuncSetAttribute(tc::tc_ce_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const bool dbg_on = getenv("TT_CE_DEBUG") != nullptr;
  long long* dbg_dev = nullptr;
  const bool has_q = pq.out || pq.dz, has_d = pd.out || pd.dz;
  if (!has_q && !has_d) return TT_OK;
  const int64_t x0 = has_q ? ceil_div(pq.Bx, tc::CE_BM) : 0, x1 = has_d ? ceil_div(pd.Bx, tc::CE_BM) : 0;
  p.pass_base = has_q ? 0 : 1;                           // only the launched passes get CTAs
  const int nz = (has_q && has_d) ? 2 : 1;
  dim3 grid((unsigned)(x0 > x1 ? x0 : x1), (unsigned)nsplit, (unsigned)nz);
  const size_t ncta = (size_t)grid.x * grid.y * grid.z, dbg_n = 2 * 64 * 8 + 8 * ncta;
  if (dbg_on) { p.dbg_pass = atoi(getenv("TT_CE_DEBUG")) == 1 ? 1 : 0; cudaMalloc(&dbg_dev, dbg_n * sizeof(long long)); cudaMemset(dbg_dev, 0, dbg_n * sizeof(long long)); p.dbg = dbg_dev; }
  // fused + 2 splits: the two CTAs of a row tile form a cluster and exchange accumulator halves through shared memory
  if (ew == 8)
    TT_CUDA(launch_kernel_cluster(tc::tc_ce_bwd_kernel<8>, grid, dim3(64 + 8 * 32), smem, s, true, fused ? (unsigned)nsplit : 1u,
                                  tmX0, tmY0, tmX1, tmY1, p));
  else
    TT_CUDA(launch_kernel_cluster(tc::tc_ce_bwd_kernel<4>, grid, dim3(tc::CE_THREADS), smem, s, true, fused ? (unsigned)nsplit : 1u,
                                  tmX0, tmY0, tmX1, tmY1, p));
  TT_LAUNCH_CHECK("tc_ce_bwd_kernel");
  if (dbg_on) {                                              // developer aid: per-tile timeline of CTA (0,0,0)
    std::vector<long long> hostv(dbg_n);
    long long* host = hostv.data();
    cudaStreamSynchronize(s);
    cudaMemcpy(host, dbg_dev, dbg_n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    long long t0 = host[0];
    for (int k = 0; k < 8; ++k) if (host[64 * 8 + k] && host[64 * 8 + k] < t0) t0 = host[64 * 8 + k];
    const int ntl = p.tiles_per_split[0] < 64 ? p.tiles_per_split[0] : 64;
    printf("[tt ce_bwd timeline, cycles since first stamp] tile: MMA{s_issue_begin,y_full,s_empty,s_issued,o_wait,p_full,o_issued} EPI{begin,s_full,loaded,computed,p_empty,p_written}\n");
    for (int t = 0; t < ntl; ++t) {
      printf("  %2d: MMA", t);
      for (int k = 0; k < 7; ++k) printf(" %6lld", host[(0 * 64 + t) * 8 + k] ? host[(0 * 64 + t) * 8 + k] - t0 : -1);
      printf("   EPI");
      for (int k = 0; k < 6; ++k) printf(" %6lld", host[(1 * 64 + t) * 8 + k] ? host[(1 * 64 + t) * 8 + k] - t0 : -1);
      printf("\n");
    }
    printf("[tt ce_bwd tail, cycles] loop_done sync1 dumped preloaded sync2 summed normalised rows_done colsum_bar colsum_done end:");
    for (int k = 0; k <= 10; ++k) printf(" %lld", host[(64 + 40) * 8 + k] ? host[(64 + 40) * 8 + k] - t0 : -1);
    printf("\n");
    printf("[tt ce_bwd O store, cycles] cb: tmem_loaded staged read stored\n");
    for (int cb = 0; cb < H / 32; ++cb) {
      printf("  %2d:", cb);
      for (int k = 0; k < 4; ++k) printf(" %6lld", host[(1 * 64 + 32 + cb) * 8 + k] - t0);
      printf("\n");
    }
    const long long* c = host + 2 * 64 * 8;
    long long g0 = 0;
    for (size_t i = 0; i < ncta; ++i) if (c[8 * i] && (!g0 || c[8 * i] < g0)) g0 = c[8 * i];
    printf("[tt ce_bwd per-CTA, ns since first CTA start] cta: start x_ready loop_done stored | fused: sync1 dumped sync2\n");
    for (size_t i = 0; i < ncta; ++i)
      printf("  cta %3zu: %6lld %6lld %6lld %6lld | %6lld %6lld %6lld\n", i, c[8 * i] - g0, c[8 * i + 1] - g0, c[8 * i + 2] - g0, c[8 * i + 3] - g0,
             c[8 * i + 4] ? c[8 * i + 4] - g0 : -1, c[8 * i + 5] ? c[8 * i + 5] - g0 : -1, c[8 * i + 6] ? c[8 * i + 6] - g0 : -1);
  }
  return TT_OK;
}

// ---- one-pass step: forward + query gradient (MODE 2), then the document gradient as its own launch --------------
size_t tc_inbatch_onepass_sync_bytes(int64_t Bq) {
  return 16 + (size_t)ceil_div(Bq, tc::CE_BM) * 4 * 2 * 4;
}
int tc_inbatch_onepass_ok(int64_t Bq, int64_t Bd, int H, float logit_bound) {
  // E = exp(logit - bound) must stay a normal fp32 / bf16 number for logit >= -bound
  return (tc_ce_supported(H) && Bq > 0 && Bd > 0 && logit_bound > 0.f && 2.0f * logit_bound * tc::kLog2e < 120.0f) ? 1 : 0;
}
int tc_inbatch_dd_nparts(int64_t x_rows, int64_t y_rows) { return cluster_splits(ceil_div(x_rows, tc::CE_BM), y_rows); }

// ---- stored-E form: what tt_inbatch_ce_fwd_dq leaves for tt_inbatch_ce_dd_stored ------------------------------------
struct CeStash { __nv_bfloat16* e; int64_t pitch; __nv_bfloat16* xs; float* wpos; size_t bytes; };
static CeStash stash_at(void* base, int64_t Bq, int64_t Bd, int H) {
  CeStash st{};
  st.pitch = (Bd + 63) / 64 * 64;
  char* b = static_cast<char*>(base);
  const size_t eb = align_up((size_t)Bq * st.pitch * 2), xb = align_up((size_t)Bq * H * 2), wb = align_up((size_t)Bq * 4);
  st.e = reinterpret_cast<__nv_bfloat16*>(b);
  st.xs = reinterpret_cast<__nv_bfloat16*>(b + eb);
  st.wpos = reinterpret_cast<float*>(b + eb + xb);
  st.bytes = eb + xb + wb;
  return st;
}
size_t tc_inbatch_stash_bytes(int64_t Bq, int64_t Bd, int H) { return stash_at(nullptr, Bq, Bd, H).bytes; }
// worth it while E stays in L2 between the two launches (126 MB, shared with everything else the step touches)
int tc_inbatch_stash_ok(int64_t Bq, int64_t Bd, int H) {
  static const size_t max_bytes = [] { const char* e = getenv("TT_CE_STASH_MAX_MB"); return (size_t)(e ? atoi(e) : 48) << 20; }();
  return (tc_ce_supported(H) && H >= 128 && Bq > 0 && Bd > 0 && (size_t)Bq * ((Bd + 63) / 64 * 64) * 2 <= max_bytes) ? 1 : 0;
}

int tc_inbatch_fwd_dq(const tt_ce_pass_t* t, int H, float inv_temp, float logit_bound, float loss_scale, const float* grad_out,
                      float* loss, float* lse_out, float* pos_mean, void* sync_scratch, const tt_p2p_t* y_exchange, const void* y_own,
                      cudaStream_t s, void* stash) {
  if (!tc_inbatch_onepass_ok(t->x_rows, t->y_rows, H, logit_bound)) {
    set_error("tc_inbatch_fwd_dq: needs H %% 64 == 0, H <= 256 and 2 * logit_bound * log2(e) < 120 (got H=%d bound=%g)", H, (double)logit_bound);
    return TT_ERR_UNSUPPORTED;
  }
  const int64_t Bx = t->x_rows, By = t->y_rows;
  const int64_t y_blk = t->y_blk > 0 ? t->y_blk : 1;
  if (y_blk < By && y_blk % tc::BWD_BN != 0) { set_error("tc_inbatch_fwd_dq: y_blk must be a multiple of %d", tc::BWD_BN); return TT_ERR_UNSUPPORTED; }
  const bool to_dz = t->dz_bf16 != nullptr;
  if (to_dz && (!t->dz_colsum || !t->inv_norm)) { set_error("tc_inbatch_fwd_dq: dz needs dz_colsum and inv_norm"); return TT_ERR_INVALID; }
  if (!to_dz && !t->out_parts) { set_error("tc_inbatch_fwd_dq: needs dz_bf16 or out_parts (dq)"); return TT_ERR_INVALID; }
  CUtensorMap tmX, tmY, tmYown, tmE;
  int rc = tc::make_tmap_bf16(&tmX, t->x_bf16, (uint64_t)Bx, (uint64_t)H, tc::CE_BM); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmY, t->y_bf16, (uint64_t)t->y_buf_rows, (uint64_t)H, tc::BWD_BN); if (rc) return rc;
  tmYown = tmY;
  tmE = tmX;
  const int64_t xt = ceil_div(Bx, tc::CE_BM);
  const int ns = cluster_splits(xt, By);
  tc::BwdParams p{};
  if (stash) {
    if (y_exchange || !tc_inbatch_stash_ok(Bx, By, H)) { set_error("tc_inbatch_fwd_dq: stash not supported for these shapes (see tt_inbatch_ce_stash_ok)"); return TT_ERR_UNSUPPORTED; }
    const CeStash st = stash_at(stash, Bx, By, H);
    rc = tc::make_tmap_bf16(&tmE, st.e, (uint64_t)Bx, (uint64_t)By, tc::CE_BM, (uint64_t)st.pitch); if (rc) return rc;
    p.e_store = 1; p.xs_out = st.xs; p.wpos_out = st.wpos;
  }
  p.lse[0] = nullptr; p.Bx[0] = Bx; p.By[0] = By; p.label_offset[0] = t->label_offset;
  p.y_blk[0] = y_blk; p.y_blk_stride[0] = t->y_blk_stride; p.y_blk_off[0] = t->y_blk_off;
  p.tiles_per_split[0] = (int)ceil_div(ceil_div(By, tc::BWD_BN), ns);
  p.out[0] = to_dz ? nullptr : t->out_parts; p.part_stride[0] = 0;
  p.dz[0] = (__nv_bfloat16*)t->dz_bf16; p.dz_colsum[0] = t->dz_colsum; p.inv_norm[0] = t->inv_norm;
  p.xg[0] = (const __nv_bfloat16*)t->x_bf16; p.yg[0] = (const __nv_bfloat16*)t->y_bf16;
  p.H = H; p.inv_temp = inv_temp; p.grad_out = grad_out; p.coef = loss_scale * inv_temp;
  p.mfix = logit_bound; p.lse_out = lse_out; p.loss = loss; p.pos_mean = pos_mean; p.loss_scale = loss_scale;
  p.counter = static_cast<unsigned*>(sync_scratch);
  p.tile_sums = reinterpret_cast<float*>(static_cast<char*>(sync_scratch) + 16);
  if (y_exchange) {
    // Y is the gathered buffer of a peer-memory exchange launched just before this call on the same stream: consume it
    // rank by rank as it lands instead of waiting for the exchange kernel to retire
    const tt_p2p_t* x = y_exchange;
    const int64_t blk = x->world > 0 ? By / x->world : 0;
    const bool ok = x->world >= 1 && x->world <= 8 && x->rank >= 0 && x->rank < x->world && By % x->world == 0 &&
                    blk % ((int64_t)tc::BWD_BN * ns) == 0 && y_blk >= By && t->y_blk_off == 0 &&
                    t->y_bf16 == static_cast<const char*>(x->base[x->rank]) + 256 && (size_t)blk * H * 2 <= x->slot_bytes &&
                    (size_t)blk * H * 2 == x->slot_bytes;
    if (!ok) { set_error("tc_inbatch_fwd_dq: y_bf16 must be the gathered slots of y_exchange (world x rows x H bf16, rows %% %d == 0)", tc::BWD_BN * ns); return TT_ERR_INVALID; }
    if (!y_own || t->label_offset != (int64_t)x->rank * blk || Bx != blk) {
      set_error("tc_inbatch_fwd_dq: the p2p form needs y_own, x_rows == y_rows / world and label_offset == rank * x_rows"); return TT_ERR_INVALID;
    }
    p.gate = static_cast<const unsigned*>(x->base[x->rank]);
    p.gate_rank = x->rank; p.gate_blk = blk; p.gate_timeout_s = x->timeout_s > 0 ? (unsigned)x->timeout_s : 600u;
    p.y_own = static_cast<const __nv_bfloat16*>(y_own);
    rc = tc::make_tmap_bf16(&tmYown, y_own, (uint64_t)blk, (uint64_t)H, tc::BWD_BN); if (rc) return rc;
  }
  const size_t smem = tc::bwd_smem(H);
  TT_CUDA(cudaFuncSetAttribute(tc::tc_ce_fwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const bool dbg_on = getenv("TT_CE_DEBUG") != nullptr;
  dim3 grid((unsigned)xt, (unsigned)ns, 1);
  long long* dbg_dev = nullptr;
  const size_t ncta = (size_t)grid.x * grid.y, dbg_n = 2 * 64 * 8 + 8 * ncta;
  if (dbg_on) { cudaMalloc(&dbg_dev, dbg_n * sizeof(long long)); cudaMemset(dbg_dev, 0, dbg_n * sizeof(long long)); p.dbg = dbg_dev; p.dbg_pass = 0; }
  TT_CUDA(launch_kernel_cluster(tc::tc_ce_fwd_dq_kernel, grid, dim3(64 + 8 * 32), smem, s, true, (unsigned)ns, tmX, tmY, tmYown, tmE, p));
  TT_LAUNCH_CHECK("tc_ce_fwd_dq_kernel");
  if (dbg_on) {
    std::vector<long long> h(dbg_n);
    cudaStreamSynchronize(s);
    cudaMemcpy(h.data(), dbg_dev, dbg_n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    const long long* c = h.data() + 2 * 64 * 8;
    long long g0 = 0;
    for (size_t i = 0; i < ncta; ++i) if (c[8 * i] && (!g0 || c[8 * i] < g0)) g0 = c[8 * i];
    printf("[tt ce_fwd_dq per-CTA, ns since first CTA start] grid %u x %u, %d tiles per CTA: start x_ready loop_done | sync1 dumped sync2 stored\n",
           grid.x, grid.y, p.tiles_per_split[0]);
    for (size_t i = 0; i < ncta; i += 7)
      printf("  cta %3zu: %6lld %6lld %6lld | %6lld %6lld %6lld %6lld\n", i, c[8 * i] - g0, c[8 * i + 1] - g0, c[8 * i + 2] - g0,
             c[8 * i + 4] - g0, c[8 * i + 5] - g0, c[8 * i + 6] - g0, c[8 * i + 3] - g0);
    const long long t0 = h[0];
    printf("[tt ce_fwd_dq tail, cycles] loop_done sync1 dumped preloaded sync2 summed normalised rows_done colsum_bar colsum_done end:");
    for (int k = 0; k <= 10; ++k) printf(" %lld", h[(64 + 40) * 8 + k] ? h[(64 + 40) * 8 + k] - t0 : -1);
    printf("\n");
    const int ntl = p.tiles_per_split[0] < 12 ? p.tiles_per_split[0] : 12;
    printf("[tt ce_fwd_dq timeline, cycles] tile: MMA{s_issue_begin,y_full,s_empty,s_issued,o_wait,p_full,o_issued} EPI{begin,s_full,loaded,computed,p_empty,p_written}\n");
    for (int i = 0; i < ntl; ++i) {
      printf("  %2d: MMA", i);
      for (int k = 0; k < 7; ++k) printf(" %6lld", h[(0 * 64 + i) * 8 + k] ? h[(0 * 64 + i) * 8 + k] - t0 : -1);
      printf("   EPI");
      for (int k = 0; k < 6; ++k) printf(" %6lld", h[(1 * 64 + i) * 8 + k] ? h[(1 * 64 + i) * 8 + k] - t0 : -1);
      printf("\n");
    }
  }
  return TT_OK;
}

// both launches as one (see tc_ce_onepass_kernel); returns TT_ERR_UNSUPPORTED when the shapes / the device do not allow it
int tc_inbatch_onepass_single(const tt_ce_pass_t* tq, const tt_ce_pass_t* td, int H, float inv_temp, float logit_bound, float loss_scale,
                              const float* grad_out, float* loss, float* lse_out, float* pos_mean, void* sync_scratch, cudaStream_t s) {
  if (!tc_inbatch_onepass_ok(tq->x_rows, tq->y_rows, H, logit_bound)) { set_error("tc_inbatch_onepass_single: unsupported shape / bound"); return TT_ERR_UNSUPPORTED; }
  const int64_t B = tq->x_rows;
  if (tq->y_rows != B || td->x_rows != B || td->y_rows != B || tq->label_offset != 0 || td->label_offset != 0 ||
      !tq->dz_bf16 || !td->dz_bf16 || !tq->dz_colsum || !td->dz_colsum || !tq->inv_norm || !td->inv_norm || td->lse != lse_out) {
    set_error("tc_inbatch_onepass_single: needs the square single-process case (Bq == Bd, offset 0), both passes in dz form, d_pass->lse == lse");
    return TT_ERR_UNSUPPORTED;
  }
  const int64_t xt = ceil_div(B, tc::CE_BM);
  const int ns = cluster_splits(xt, B);
  if (xt * ns > kNumSMs) { set_error("tc_inbatch_onepass_single: grid larger than the SM count"); return TT_ERR_UNSUPPORTED; }
  CUtensorMap tmXq, tmYq, tmXd, tmYd;
  int rc = tc::make_tmap_bf16(&tmXq, tq->x_bf16, (uint64_t)B, (uint64_t)H, tc::CE_BM); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmYq, tq->y_bf16, (uint64_t)tq->y_buf_rows, (uint64_t)H, tc::BWD_BN); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmXd, td->x_bf16, (uint64_t)B, (uint64_t)H, tc::CE_BM); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmYd, td->y_bf16, (uint64_t)td->y_buf_rows, (uint64_t)H, tc::BWD_BN); if (rc) return rc;
  tc::BwdParams p{};
  const tt_ce_pass_t* ps[2] = {tq, td};
  for (int k = 0; k < 2; ++k) {
    p.lse[k] = k == 0 ? nullptr : lse_out; p.Bx[k] = B; p.By[k] = B; p.label_offset[k] = 0;
    p.y_blk[k] = ps[k]->y_blk > 0 ? ps[k]->y_blk : 1; p.y_blk_stride[k] = ps[k]->y_blk_stride; p.y_blk_off[k] = ps[k]->y_blk_off;
    p.tiles_per_split[k] = (int)ceil_div(ceil_div(B, tc::BWD_BN), ns);
    p.out[k] = nullptr; p.part_stride[k] = 0;
    p.dz[k] = (__nv_bfloat16*)ps[k]->dz_bf16; p.dz_colsum[k] = ps[k]->dz_colsum; p.inv_norm[k] = ps[k]->inv_norm;
    p.xg[k] = (const __nv_bfloat16*)ps[k]->x_bf16; p.yg[k] = (const __nv_bfloat16*)ps[k]->y_bf16;
  }
  p.H = H; p.inv_temp = inv_temp; p.grad_out = grad_out; p.coef = loss_scale * inv_temp;
  p.mfix = logit_bound; p.lse_out = lse_out; p.loss = loss; p.pos_mean = pos_mean; p.loss_scale = loss_scale;
  p.counter = static_cast<unsigned*>(sync_scratch);
  p.tile_sums = reinterpret_cast<float*>(static_cast<char*>(sync_scratch) + 16);
  const size_t smem = tc::bwd_smem(H);
  TT_CUDA(cudaFuncSetAttribute(tc::tc_ce_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)xt, (unsigned)ns, 1);
  // every cluster of the grid must be resident at the same time (grid barrier): ask the occupancy calculator once per shape
  static int ok_key = -1, ok_val = 0;
  const int key = (int)(xt * 8 + ns) * 1024 + H;
  if (ok_key != key) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(64 + 8 * 32); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = (unsigned)ns; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = ns > 1 ? 1 : 0;
    int nclusters = 0;
    if (ns > 1) {
      TT_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, tc::tc_ce_onepass_kernel, &cfg));
    } else {
      int per_sm = 0;
      TT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tc::tc_ce_onepass_kernel, 64 + 8 * 32, smem));
      nclusters = per_sm * kNumSMs;
    }
    ok_key = key; ok_val = nclusters >= (int)xt ? 1 : 0;
  }
  if (!ok_val) { set_error("tc_inbatch_onepass_single: the grid's clusters cannot all be resident at once"); return TT_ERR_UNSUPPORTED; }
  TT_CUDA(launch_kernel_cluster(tc::tc_ce_onepass_kernel, grid, dim3(64 + 8 * 32), smem, s, true, (unsigned)ns, tmXq, tmYq, tmXd, tmYd, p));
  TT_LAUNCH_CHECK("tc_ce_onepass_kernel");
  return TT_OK;
}

// document gradient alone: every field of tt_ce_pass_t is honoured; dz -> cluster tail, out_parts -> nparts slices
int tc_inbatch_dd(const tt_ce_pass_t* t, int H, float inv_temp, float loss_scale, const float* grad_out, cudaStream_t s) {
  if (!tc_ce_supported(H)) { set_error("tc_inbatch_dd: needs H %% 64 == 0, H <= 256"); return TT_ERR_UNSUPPORTED; }
  CePass pd{};
  pd.x = (const __nv_bfloat16*)t->x_bf16; pd.Bx = t->x_rows; pd.y = (const __nv_bfloat16*)t->y_bf16; pd.By = t->y_rows;
  pd.y_buf_rows = t->y_buf_rows; pd.y_blk = t->y_blk; pd.y_blk_stride = t->y_blk_stride; pd.y_blk_off = t->y_blk_off;
  pd.lse = t->lse; pd.label_offset = t->label_offset; pd.out = t->out_parts; pd.part_stride = t->part_stride;
  pd.dz = (__nv_bfloat16*)t->dz_bf16; pd.dz_colsum = t->dz_colsum; pd.inv_norm = t->inv_norm;
  if (pd.y_blk < pd.By && pd.y_blk % tc::BWD_BN != 0) { set_error("tc_inbatch_dd: y_blk must be a multiple of %d", tc::BWD_BN); return TT_ERR_UNSUPPORTED; }
  CePass pq = pd;                                           // tensor maps need valid operands; the pass itself is not launched
  pq.out = nullptr; pq.dz = nullptr; pq.dz_colsum = nullptr; pq.inv_norm = nullptr;
  return launch_tc_bwd(pq, pd, H, inv_temp, tc_inbatch_dd_nparts(pd.Bx, pd.By), grad_out, loss_scale * inv_temp, s);
}

// document gradient from the stash tt_inbatch_ce_fwd_dq(stash = ...) left: x = documents, y = the queries of that call
int tc_inbatch_dd_stored(const tt_ce_pass_t* t, int H, float inv_temp, float loss_scale, const float* grad_out, const void* stash,
                         cudaStream_t s) {
  const int64_t Bx = t->x_rows, By = t->y_rows;            // documents, queries
  if (!tc_inbatch_stash_ok(By, Bx, H)) { set_error("tc_inbatch_dd_stored: unsupported shapes (see tt_inbatch_ce_stash_ok)"); return TT_ERR_UNSUPPORTED; }
  const bool to_dz = t->dz_bf16 != nullptr;
  if (to_dz && (!t->dz_colsum || !t->inv_norm)) { set_error("tc_inbatch_dd_stored: dz needs dz_colsum and inv_norm"); return TT_ERR_INVALID; }
  if (!to_dz && !t->out_parts) { set_error("tc_inbatch_dd_stored: needs dz_bf16 or out_parts (dd)"); return TT_ERR_INVALID; }
  const CeStash st = stash_at(const_cast<void*>(stash), By, Bx, H);
  CUtensorMap tmXs, tmE;
  int rc = tc::make_tmap_bf16(&tmXs, st.xs, (uint64_t)By, (uint64_t)H, 64); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmE, st.e, (uint64_t)By, (uint64_t)Bx, 64, (uint64_t)st.pitch); if (rc) return rc;
  const int64_t xt = ceil_div(Bx, tc::CE_BM);
  const int ns = cluster_splits(xt, By);
  tc::BwdParams p{};
  p.Bx[1] = Bx; p.By[1] = By; p.label_offset[1] = t->label_offset;
  p.y_blk[1] = t->y_blk > 0 ? t->y_blk : 1; p.y_blk_stride[1] = t->y_blk_stride; p.y_blk_off[1] = t->y_blk_off;
  p.tiles_per_split[1] = (int)ceil_div(ceil_div(By, tc::BWD_BN), ns);
  p.out[1] = to_dz ? nullptr : t->out_parts; p.part_stride[1] = 0;
  p.dz[1] = (__nv_bfloat16*)t->dz_bf16; p.dz_colsum[1] = t->dz_colsum; p.inv_norm[1] = t->inv_norm;
  p.xg[1] = (const __nv_bfloat16*)t->x_bf16; p.yg[1] = (const __nv_bfloat16*)t->y_bf16;
  p.H = H; p.inv_temp = inv_temp; p.grad_out = grad_out; p.coef = loss_scale * inv_temp;
  p.wpos_in = st.wpos;
  const size_t smem = tc::bwd_smem(H);
  TT_CUDA(cudaFuncSetAttribute(tc::tc_ce_dd_stored_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const bool dbg_on = getenv("TT_CE_DEBUG") != nullptr;
  dim3 grid((unsigned)xt, (unsigned)ns, 1);
  long long* dbg_dev = nullptr;
  const size_t ncta = (size_t)grid.x * grid.y, dbg_n = 2 * 64 * 8 + 8 * ncta;
  if (dbg_on) { cudaMalloc(&dbg_dev, dbg_n * sizeof(long long)); cudaMemset(dbg_dev, 0, dbg_n * sizeof(long long)); p.dbg = dbg_dev; p.dbg_pass = 0; }
  TT_CUDA(launch_kernel_cluster(tc::tc_ce_dd_stored_kernel, grid, dim3(64 + 8 * 32), smem, s, true, (unsigned)ns, tmXs, tmE, p));
  TT_LAUNCH_CHECK("tc_ce_dd_stored_kernel");
  if (dbg_on) {
    std::vector<long long> h(dbg_n);
    cudaStreamSynchronize(s);
    cudaMemcpy(h.data(), dbg_dev, dbg_n * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    const long long* c = h.data() + 2 * 64 * 8;
    long long g0 = 0;
    for (size_t i = 0; i < ncta; ++i) if (c[8 * i] && (!g0 || c[8 * i] < g0)) g0 = c[8 * i];
    printf("[tt ce_dd_stored per-CTA, ns since first CTA start] grid %u x %u: start loop_begin loop_done | dumped sync2 stored\n", grid.x, grid.y);
    for (size_t i = 0; i < ncta; i += 7)
      printf("  cta %3zu: %6lld %6lld %6lld | %6lld %6lld %6lld\n", i, c[8 * i] - g0, c[8 * i + 1] - g0, c[8 * i + 2] - g0,
             c[8 * i + 4] - g0, c[8 * i + 6] - g0, c[8 * i + 3] - g0);
  }
  return TT_OK;
}

static CePass plain_pass(const __nv_bfloat16* x, int64_t Bx, const __nv_bfloat16* y, int64_t By, const float* lse, int64_t off,
                         float* out, int64_t stride) {
  CePass c{};
  c.x = x; c.Bx = Bx; c.y = y; c.By = By; c.y_buf_rows = By; c.y_blk = By > 0 ? By : 1; c.y_blk_stride = 0; c.y_blk_off = 0;
  c.lse = lse; c.label_offset = off; c.out = out; c.part_stride = stride;
  return c;
}

int tc_inbatch_bwd_nparts(int64_t Bq, int64_t Bd, int H) { return tc_ce_supported(H) ? bwd_splits(Bq, Bd) : 1; }
int tc_inbatch_bwd_nparts2(int64_t Bx0, int64_t By0, int64_t Bx1, int64_t By1, int H) {
  return tc_ce_supported(H) ? bwd_splits2(Bx0, By0, Bx1, By1) : 1;
}

// Partial-slice variant used by the fused trainer: the consumer (normalise-backward) sums the slices.
int tc_inbatch_bwd_parts(const __nv_bfloat16* q_bf16, const __nv_bfloat16* d_bf16, const float* lse, int64_t Bq, int64_t Bd,
                         int H, float inv_temp, int64_t label_offset, float loss_scale, const float* grad_out,
                         float* dq_parts, int64_t stride_q, float* dd_parts, int64_t stride_d, cudaStream_t s) {
  if (!tc_ce_supported(H) || !q_bf16 || !d_bf16) { set_error("tc_inbatch_bwd_parts: needs bf16 operands and H %% 64 == 0, H <= 256"); return TT_ERR_UNSUPPORTED; }
  const CePass pq = plain_pass(q_bf16, Bq, d_bf16, Bd, lse, label_offset, dq_parts, stride_q);
  const CePass pd = plain_pass(d_bf16, Bd, q_bf16, Bq, lse, label_offset, dd_parts, stride_d);
  return launch_tc_bwd(pq, pd, H, inv_temp, bwd_splits(Bq, Bd), grad_out, loss_scale * inv_temp, s);
}

// General two-pass form (data-parallel training with global negatives): every field of tt_ce_pass_t is honoured.
int tc_inbatch_bwd_parts_ex(const tt_ce_pass_t* q_pass, const tt_ce_pass_t* d_pass, int H, float inv_temp, float loss_scale,
                            const float* grad_out, int nparts, cudaStream_t s) {
  if (!tc_ce_supported(H)) { set_error("tc_inbatch_bwd_parts_ex: needs H %% 64 == 0, H <= 256"); return TT_ERR_UNSUPPORTED; }
  auto conv = [](const tt_ce_pass_t* t) {
    CePass c{};
    c.x = (const __nv_bfloat16*)t->x_bf16; c.Bx = t->x_rows; c.y = (const __nv_bfloat16*)t->y_bf16; c.By = t->y_rows;
    c.y_buf_rows = t->y_buf_rows; c.y_blk = t->y_blk; c.y_blk_stride = t->y_blk_stride; c.y_blk_off = t->y_blk_off;
    c.lse = t->lse; c.label_offset = t->label_offset; c.out = t->out_parts; c.part_stride = t->part_stride;
    c.dz = (__nv_bfloat16*)t->dz_bf16; c.dz_colsum = t->dz_colsum; c.inv_norm = t->inv_norm;
    return c;
  };
  const CePass pq = conv(q_pass), pd = conv(d_pass);
  const int want = bwd_splits2(pq.Bx, pq.By, pd.Bx, pd.By);
  if (nparts != want) { set_error("tc_inbatch_bwd_parts_ex: nparts %d != %d (query tt_inbatch_ce_bwd_nparts_ex)", nparts, want); return TT_ERR_INVALID; }
  // a tile may not straddle two blocks of an interleaved buffer; a single block (y_blk >= rows) has no such constraint
  if ((pq.y_blk < pq.By && pq.y_blk % tc::BWD_BN != 0) || (pd.y_blk < pd.By && pd.y_blk % tc::BWD_BN != 0)) {
    set_error("tc_inbatch_bwd_parts_ex: y_blk must be a multiple of %d", tc::BWD_BN); return TT_ERR_UNSUPPORTED;
  }
  return launch_tc_bwd(pq, pd, H, inv_temp, nparts, grad_out, loss_scale * inv_temp, s);
}

int tc_inbatch_bwd(const float* q, const float* d, const __nv_bfloat16* q_bf16, const __nv_bfloat16* d_bf16,
                   const float* lse, int64_t Bq, int64_t Bd, int H, float inv_temp, int64_t label_offset,
                   float loss_scale, const float* grad_out, float* dq, float* dd, void* ws, size_t ws_bytes,
                   cudaStream_t s) {
  if (!tc_ce_supported(H))
    return inbatch_ce_bwd_fp32(q, d, lse, Bq, Bd, H, inv_temp, label_offset, loss_scale, grad_out, dq, dd, ws, ws_bytes, s);
  const TcCePlan plan = plan_tc_ce(Bq, Bd, H);
  if (ws == nullptr || ws_bytes < plan.total) { set_error("tc_inbatch_bwd: workspace too small (%zu < %zu)", ws_bytes, plan.total); return TT_ERR_WORKSPACE; }
  Workspace w(ws, ws_bytes);
  __nv_bfloat16* qb = w.take<__nv_bfloat16>((size_t)Bq * H);
  __nv_bfloat16* db = w.take<__nv_bfloat16>((size_t)Bd * H);
  (void)w.take<float>(plan.ml / 4);
  (void)w.take<float>(Bq);
  float* partial = plan.partial ? w.take<float>(plan.partial / 4) : nullptr;
  int rc;
  if (!q_bf16 || !d_bf16) {
    rc = tc::cast3_public(q_bf16 ? nullptr : q, qb, q_bf16 ? 0 : Bq * H, d_bf16 ? nullptr : d, db, d_bf16 ? 0 : Bd * H, s);
    if (rc) return rc;
  }
  const __nv_bfloat16* qa = q_bf16 ? q_bf16 : qb;
  const __nv_bfloat16* da = d_bf16 ? d_bf16 : db;
  const float coef = loss_scale * inv_temp;
  const int ns = plan.ns_b;
  float* pq_buf = ns > 1 ? partial : dq;                                   // [ns][Bq][H]
  float* pd_buf = ns > 1 ? partial + (size_t)ns * Bq * H : dd;             // [ns][Bd][H]
  const CePass pq = plain_pass(qa, Bq, da, Bd, lse, label_offset, dq ? pq_buf : nullptr, Bq * H);
  const CePass pd = plain_pass(da, Bd, qa, Bq, lse, label_offset, dd ? pd_buf : nullptr, Bd * H);
  rc = launch_tc_bwd(pq, pd, H, inv_temp, ns, grad_out, coef, s);
  if (rc) return rc;
  if (ns > 1) {
    if (dq) { rc = split_sum(pq_buf, ns, Bq * H, dq, s); if (rc) return rc; }
    if (dd) { rc = split_sum(pd_buf, ns, Bd * H, dd, s); if (rc) return rc; }
  }
  return TT_OK;
}

}  // namespace tt
