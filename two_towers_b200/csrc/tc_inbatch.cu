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
// Each output row has one owner CTA per split and splits are summed in split order:
// bitwise-reproducible gradients.
#include <math_constants.h>

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
constexpr int CE_THREADS = 192;     // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr float kLog2e = 1.4426950408889634f;

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
constexpr int FWD_STAGES = 4;

__global__ void __launch_bounds__(CE_THREADS, 1)
tc_ce_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmD, int64_t Bq,
                 int64_t Bd, int H, float inv_temp, int64_t label_offset, int tiles_per_split,
                 float* __restrict__ part_ml, float* __restrict__ pos_logit) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  const int kq = H / 64;                                  // 64-wide K blocks
  const uint32_t q_bytes = (uint32_t)CE_BM * H * 2, d_bytes = (uint32_t)CE_BN * H * 2;
  uint8_t* q_tile = base;
  uint8_t* d_tiles = base + q_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(d_tiles + FWD_STAGES * d_bytes);
  uint64_t* q_bar = bars;
  uint64_t* d_full = bars + 1;
  uint64_t* d_empty = d_full + FWD_STAGES;
  uint64_t* s_full = d_empty + FWD_STAGES;                // [2]
  uint64_t* s_empty = s_full + 2;                         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t x0 = (int64_t)blockIdx.x * CE_BM;
  const int ntiles = (int)ceil_div(Bd, CE_BN);
  const int t_beg = blockIdx.y * tiles_per_split;
  const int t_end = min(ntiles, t_beg + tiles_per_split);
  const int nt = max(0, t_end - t_beg);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmD);
    mbar_init(q_bar, 1);
    for (int s = 0; s < FWD_STAGES; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_bar, q_bytes);
      for (int kb = 0; kb < kq; ++kb) tma_load_2d(q_tile + kb * (CE_BM * 128), &tmQ, q_bar, kb * 64, (int)x0);
      for (int i = 0; i < nt; ++i) {
        const int s = i % FWD_STAGES;
        mbar_wait(&d_empty[s], ((i / FWD_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&d_full[s], d_bytes);
        uint8_t* dt = d_tiles + s * d_bytes;
        for (int kb = 0; kb < kq; ++kb) tma_load_2d(dt + kb * (CE_BN * 128), &tmD, &d_full[s], kb * 64, (t_beg + i) * CE_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(CE_BM, CE_BN, 0, 0);
      mbar_wait(q_bar, 0);
      for (int i = 0; i < nt; ++i) {
        const int s = i % FWD_STAGES, b = i & 1;
        mbar_wait(&d_full[s], (i / FWD_STAGES) & 1);
        mbar_wait(&s_empty[b], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(q_tile), da = smem_u32(d_tiles + s * d_bytes);
        for (int kb = 0; kb < kq; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_s + b * CE_BN, umma_desc_kmajor(qa + kb * (CE_BM * 128), k),
                      umma_desc_kmajor(da + kb * (CE_BN * 128), k), idesc, (kb | k) != 0);
        umma_commit(&d_empty[s]);
        umma_commit(&s_full[b]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int64_t row = x0 + quarter * 32 + lane;
    const int64_t pcol = row + label_offset;
    const float c = inv_temp * kLog2e;
    float m = -CUDART_INF_F, l = 0.f;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const int64_t y0 = (int64_t)(t_beg + i) * CE_BN;
      mbar_wait(&s_full[b], (i >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t ta = tmem_s + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * CE_BN);
      tmem_ld_x32(ta, r0);
      tmem_ld_x32(ta + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[b]);             // TMEM buffer may be overwritten
      float tmax = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float v0 = (y0 + j < Bd) ? __uint_as_float(r0[j]) : -CUDART_INF_F;
        float v1 = (y0 + 32 + j < Bd) ? __uint_as_float(r1[j]) : -CUDART_INF_F;
        r0[j] = __float_as_uint(v0); r1[j] = __float_as_uint(v1);
        tmax = fmaxf(tmax, fmaxf(v0, v1));
        if (y0 + j == pcol && row < Bq) pos_logit[row] = v0 * inv_temp;
        if (y0 + 32 + j == pcol && row < Bq) pos_logit[row] = v1 * inv_temp;
      }
      const float mnew = fmaxf(m, tmax);
      const float mc = mnew * c;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        sum += exp2f(fmaf(__uint_as_float(r0[j]), c, -mc));
        sum += exp2f(fmaf(__uint_as_float(r1[j]), c, -mc));
      }
      l = l * exp2f((m - mnew) * c) + sum;
      m = mnew;
    }
    if (row < Bq && nt > 0) {
      part_ml[((int64_t)blockIdx.y * Bq + row) * 2 + 0] = m * inv_temp;
      part_ml[((int64_t)blockIdx.y * Bq + row) * 2 + 1] = l;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_s, 128);
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
constexpr int BWD_STAGES = 3;

template <bool COL_LSE>
__global__ void __launch_bounds__(CE_THREADS, 1)
tc_ce_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                 const float* __restrict__ lse, int64_t Bx, int64_t By, int H, float inv_temp, int64_t label_offset,
                 int tiles_per_split, const float* __restrict__ grad_out, float coef, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  const int kq = H / 64;
  const uint32_t x_bytes = (uint32_t)CE_BM * H * 2, y_bytes = (uint32_t)CE_BN * H * 2;
  constexpr uint32_t p_bytes = CE_BM * CE_BN * 2;         // 16 KB, K-major 128 rows x 64
  uint8_t* x_tile = base;
  uint8_t* y_tiles = x_tile + x_bytes;
  uint8_t* p_tiles = y_tiles + BWD_STAGES * y_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_tiles + 2 * p_bytes);
  uint64_t* x_bar = bars;
  uint64_t* y_full = bars + 1;
  uint64_t* y_empty = y_full + BWD_STAGES;
  uint64_t* s_full = y_empty + BWD_STAGES;                // [2]
  uint64_t* s_empty = s_full + 2;                         // [2]
  uint64_t* p_full = s_empty + 2;                         // [2]
  uint64_t* p_empty = p_full + 2;                         // [2]
  uint64_t* o_full = p_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t x0 = (int64_t)blockIdx.x * CE_BM;
  const int ntiles = (int)ceil_div(By, CE_BN);
  const int t_beg = blockIdx.y * tiles_per_split;
  const int t_end = min(ntiles, t_beg + tiles_per_split);
  const int nt = max(0, t_end - t_beg);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmY);
    mbar_init(x_bar, 1);
    for (int s = 0; s < BWD_STAGES; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], 4);
      mbar_init(&p_full[b], 4); mbar_init(&p_empty[b], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base;                      // columns [0, H)
  const uint32_t tmem_s = tmem_base + 256;                // columns [256, 384): two S buffers

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(x_bar, x_bytes);
      for (int kb = 0; kb < kq; ++kb) tma_load_2d(x_tile + kb * (CE_BM * 128), &tmX, x_bar, kb * 64, (int)x0);
      for (int i = 0; i < nt; ++i) {
        const int s = i % BWD_STAGES;
        mbar_wait(&y_empty[s], ((i / BWD_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&y_full[s], y_bytes);
        uint8_t* yt = y_tiles + s * y_bytes;
        for (int kb = 0; kb < kq; ++kb) tma_load_2d(yt + kb * (CE_BN * 128), &tmY, &y_full[s], kb * 64, (t_beg + i) * CE_BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nt > 0) {
      const uint32_t idesc_s = umma_idesc_bf16(CE_BM, CE_BN, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(CE_BM, H, 0, 1);     // B = Y tile read MN-major
      const uint32_t xa = smem_u32(x_tile);
      auto issue_s = [&](int i) {
        const int s = i % BWD_STAGES, b = i & 1;
        mbar_wait(&y_full[s], (i / BWD_STAGES) & 1);
        mbar_wait(&s_empty[b], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t ya = smem_u32(y_tiles + s * y_bytes);
        for (int kb = 0; kb < kq; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_s + b * CE_BN, umma_desc_kmajor(xa + kb * (CE_BM * 128), k),
                      umma_desc_kmajor(ya + kb * (CE_BN * 128), k), idesc_s, (kb | k) != 0);
        umma_commit(&s_full[b]);
      };
      mbar_wait(x_bar, 0);
      issue_s(0);
      for (int i = 0; i < nt; ++i) {
        if (i + 1 < nt) issue_s(i + 1);
        const int s = i % BWD_STAGES, b = i & 1;
        mbar_wait(&p_full[b], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(p_tiles + b * p_bytes), ya = smem_u32(y_tiles + s * y_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_o, umma_desc_kmajor(pa, k), umma_desc_mnmajor(ya, k, CE_BN * 128), idesc_o, (i | k) != 0);
        umma_commit(&p_empty[b]);
        umma_commit(&y_empty[s]);
      }
      umma_commit(o_full);
    }
  } else {
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;                  // row inside the tile == TMEM lane
    const int64_t row = x0 + lrow;
    const float c = inv_temp * kLog2e;
    const float row_lse = (!COL_LSE && row < Bx) ? lse[row] * kLog2e : 0.f;
    for (int i = 0; i < nt; ++i) {
      const int b = i & 1;
      const int64_t y0 = (int64_t)(t_beg + i) * CE_BN;
      mbar_wait(&s_full[b], (i >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t ta = tmem_s + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * CE_BN);
      tmem_ld_x32(ta, r0);
      tmem_ld_x32(ta + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[b]);
      // P = exp2(S*c - lse*log2e) - [positive]; packed to bf16 pairs in place
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        float pv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int jj = j + u;
          const int64_t col = y0 + jj;
          const float sv = __uint_as_float(jj < 32 ? r0[jj] : r1[jj - 32]);
          float p = 0.f;
          if (col < By && row < Bx) {
            const float lv = COL_LSE ? __ldg(lse + col) * kLog2e : row_lse;
            p = exp2f(fmaf(sv, c, -lv));
            const bool pos = COL_LSE ? (row == col + label_offset) : (col == row + label_offset);
            if (pos) p -= 1.0f;
          }
          pv[u] = p;
        }
        pk[j >> 1] = pack_bf16x2(pv[0], pv[1]);
      }
      mbar_wait(&p_empty[b], ((i >> 1) & 1) ^ 1);          // O-GEMM(i-2) has finished reading this P buffer
      uint8_t* prow = p_tiles + b * p_bytes + lrow * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {                      // 8 x 16-byte chunks, 128B swizzle: chunk ^= row & 7
        uint4 v = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        *reinterpret_cast<uint4*>(prow + ((ch ^ (lrow & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
    }
    // final: O (TMEM) -> global, scaled by grad * loss_scale / temperature
    const float scale = coef * (grad_out ? *grad_out : 1.0f);
    float* o = out + (int64_t)blockIdx.y * Bx * H;
    if (nt > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
    }
    for (int cb = 0; cb < H / 32; ++cb) {
      uint32_t r[32];
      if (nt > 0) {
        tmem_ld_x32(tmem_o + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32), r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (row < Bx) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]) * scale, __uint_as_float(r[j + 1]) * scale,
                                 __uint_as_float(r[j + 2]) * scale, __uint_as_float(r[j + 3]) * scale);
          *reinterpret_cast<float4*>(o + row * H + cb * 32 + j) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static size_t fwd_smem(int H) { return 1024 + (size_t)CE_BM * H * 2 + FWD_STAGES * (size_t)CE_BN * H * 2 + 16 * 8 + 16; }
static size_t bwd_smem(int H) {
  return 1024 + (size_t)CE_BM * H * 2 + BWD_STAGES * (size_t)CE_BN * H * 2 + 2 * (size_t)CE_BM * CE_BN * 2 + 20 * 8 + 16;
}

static int pick_split(int64_t Bx, int64_t By) {
  const int64_t xt = ceil_div(Bx, CE_BM), yt = ceil_div(By, CE_BN);
  int64_t s = kNumSMs / xt;                 // one wave
  if (s > yt) s = yt;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  const int64_t per = ceil_div(yt, s);
  return (int)ceil_div(yt, per);
}

}  // namespace tc

static bool tc_ce_supported(int H) { return H % 64 == 0 && H >= 64 && H <= 256; }

struct TcCePlan { int ns_f, ns_q, ns_d; size_t qb, db, ml, pos, partial, total; };
static TcCePlan plan_tc_ce(int64_t Bq, int64_t Bd, int H) {
  TcCePlan p{};
  p.ns_f = tc::pick_split(Bq, Bd);
  p.ns_q = tc::pick_split(Bq, Bd);
  p.ns_d = tc::pick_split(Bd, Bq);
  p.qb = align_up((size_t)Bq * H * 2);
  p.db = align_up((size_t)Bd * H * 2);
  p.ml = align_up((size_t)p.ns_f * Bq * 2 * 4);
  p.pos = align_up((size_t)Bq * 4);
  size_t a = p.ns_q > 1 ? (size_t)p.ns_q * Bq * H * 4 : 0, b = p.ns_d > 1 ? (size_t)p.ns_d * Bd * H * 4 : 0;
  p.partial = align_up(a > b ? a : b);
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
  int rc;
  if (!q_bf16 || !d_bf16) {
    rc = tc::cast3_public(q_bf16 ? nullptr : q, qb, q_bf16 ? 0 : Bq * H, d_bf16 ? nullptr : d, db, d_bf16 ? 0 : Bd * H, s);
    if (rc) return rc;
  }
  const __nv_bfloat16* qa = q_bf16 ? q_bf16 : qb;
  const __nv_bfloat16* da = d_bf16 ? d_bf16 : db;
  CUtensorMap tmQ, tmD;
  rc = tc::make_tmap_bf16(&tmQ, qa, (uint64_t)Bq, (uint64_t)H, tc::CE_BM); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmD, da, (uint64_t)Bd, (uint64_t)H, tc::CE_BN); if (rc) return rc;
  const int yt = (int)ceil_div(Bd, tc::CE_BN);
  const int per = (int)ceil_div(yt, plan.ns_f);
  const size_t smem = tc::fwd_smem(H);
  TT_CUDA(cudaFuncSetAttribute(tc::tc_ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(Bq, tc::CE_BM), (unsigned)plan.ns_f);
  tc::tc_ce_fwd_kernel<<<grid, tc::CE_THREADS, smem, s>>>(tmQ, tmD, Bq, Bd, H, inv_temp, label_offset, per, part_ml, pos);
  TT_LAUNCH_CHECK("tc_ce_fwd_kernel");
  return inbatch_finalize(part_ml, pos, plan.ns_f, Bq, inv_temp, loss_scale, lse, loss, pos_mean, nullptr, s);
}

template <bool COL>
static int launch_tc_bwd(const __nv_bfloat16* X, const __nv_bfloat16* Y, const float* lse, int64_t Bx, int64_t By, int H,
                         float inv_temp, int64_t off, int nsplit, const float* grad_out, float coef, float* out,
                         float* partial, cudaStream_t s) {
  CUtensorMap tmX, tmY;
  int rc = tc::make_tmap_bf16(&tmX, X, (uint64_t)Bx, (uint64_t)H, tc::CE_BM); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmY, Y, (uint64_t)By, (uint64_t)H, tc::CE_BN); if (rc) return rc;
  const int yt = (int)ceil_div(By, tc::CE_BN);
  const int per = (int)ceil_div(yt, nsplit);
  const size_t smem = tc::bwd_smem(H);
  TT_CUDA(cudaFuncSetAttribute(tc::tc_ce_bwd_kernel<COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(Bx, tc::CE_BM), (unsigned)nsplit);
  float* dst = nsplit > 1 ? partial : out;
  tc::tc_ce_bwd_kernel<COL><<<grid, tc::CE_THREADS, smem, s>>>(tmX, tmY, lse, Bx, By, H, inv_temp, off, per, grad_out, coef, dst);
  TT_LAUNCH_CHECK("tc_ce_bwd_kernel");
  if (nsplit > 1) return split_sum(partial, nsplit, Bx * H, out, s);
  return TT_OK;
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
  if (dq) {
    rc = launch_tc_bwd<false>(qa, da, lse, Bq, Bd, H, inv_temp, label_offset, plan.ns_q, grad_out, coef, dq, partial, s);
    if (rc) return rc;
  }
  if (dd) {
    rc = launch_tc_bwd<true>(da, qa, lse, Bd, Bq, H, inv_temp, label_offset, plan.ns_d, grad_out, coef, dd, partial, s);
    if (rc) return rc;
  }
  return TT_OK;
}

}  // namespace tt
