// tc_mlp.cu -- the whole tower MLP forward as ONE tcgen05 kernel (TT_PREC_BF16, K3):
//
//     y = normalize( relu(x W1^T + b1) W2^T + b2 )        twotower/encoders.py:38-42,77
//
// One CTA owns 128 rows.  Both weight matrices live in shared memory, the hidden activation never
// leaves the SM:
//   TMA   : x tile [128,E] + W1 [H,E]  ->  GEMM 1 (tcgen05.mma, fp32 accumulator in TMEM columns [0,H))
//   warps : TMEM -> +b1 -> ReLU -> bf16 -> 128B-swizzled smem tile (the A operand of GEMM 2) and -> global h1
//   TMA   : W2 [H,H] (k-block 0 re-uses the x/W1 bytes once GEMM 1 has retired)  ->  GEMM 2 (TMEM columns [256,256+H))
//   warps : TMEM -> +b2 -> row sum of squares -> z (fp32, saved for backward), y (bf16 and optional fp32)
// Every epilogue thread owns one TMEM lane == one row, so the row L2 norm is a private register reduction.
// Compared with the unfused path (GEMM, GEMM, normalise) this is 1 launch instead of 3 and the
// [R,H] hidden / pre-normalise tensors are written once and never re-read in the forward.
#include <math_constants.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tensor_core.cuh"

namespace tt {
namespace tc {

constexpr int MLP_BM = 128;
constexpr int MLP_THREADS = 320;        // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)

struct MlpFwdParams {
  int64_t R;
  int E, H;
  const float* b1;
  const float* b2;
  __nv_bfloat16* h1b;     // [R,H]
  float* z;               // [R,H] nullable
  float* y;               // [R,H] nullable
  __nv_bfloat16* yb;      // [R,H] nullable
  float* inv_norm;        // [R] nullable: 1 / max(|z|, 1e-12)
  int V;                  // > 0: x = P * table is formed in-kernel (GEMM 0) from the pooling matrix and the bf16 table
  // ids != null (pool mode): the pooling matrix is BUILT here from the token ids (one integer histogram per row in shared
  // memory, twotower/encoders.py:62-72 mask + mean) and stored to P [R,V] for the backward -- the separate histogram launch disappears
  const void* ids;        // [R, L] int32 / int64, 0 = padding
  int id_bytes, L;
  float* inv_len;         // [R] nullable: 1 / (number of non-pad tokens + 1e-9)
  long long* bad_id;      // mapped host word that records an out-of-range token id (nullable)
  long long* dbg;         // developer aid (TT_MLP_DEBUG): CTA 0, thread 64: %globaltimer at each phase boundary
};

__global__ void __launch_bounds__(MLP_THREADS, 1)
tc_mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmH1,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmT, const MlpFwdParams p) {
  // pool mode (p.V > 0): tmX maps the pooling matrix P [R,V] and tmT the bf16 table [V,E].  GEMM 0 forms
  // x = P table (accumulator in TMEM columns [256, 256+E)), the warps round it to bf16 straight into the swizzled
  // x tile GEMM 1 reads: the pooled activations never touch HBM.  P and the table borrow the hidden-tile bytes.
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
#define TT_MLP_STAMP(slot) do { if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.dbg[slot] = t_; } } while (0)
  TT_MLP_STAMP(0);
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const int E = p.E, H = p.H;
  const int kE = E / 64, kH = H / 64;
  const uint32_t x_bytes = (uint32_t)kE * MLP_BM * 128;            // x tile, kE k-blocks of [128 rows x 128 B]
  const uint32_t w1_bytes = (uint32_t)kE * H * 128;                // W1, kE k-blocks of [H rows x 128 B]
  const uint32_t w2_blk = (uint32_t)H * 128;                       // one W2 k-block [H rows x 128 B]
  const uint32_t a_bytes = x_bytes + w1_bytes;                     // region A (>= w2_blk, checked on the host)
  uint8_t* x_tile = base;
  uint8_t* w1_tile = base + x_bytes;
  uint8_t* w2_rest = base + a_bytes;                               // W2 k-blocks 1..kH-1
  uint8_t* h1_tile = w2_rest + (uint32_t)(kH - 1) * w2_blk;        // [128 x H] bf16, kH k-blocks of 16 KB
  uint8_t* after = h1_tile + (uint32_t)kH * MLP_BM * 128;
  float* bias_s = reinterpret_cast<float*>(after);                 // b1[H], b2[H]
  uint64_t* bars = reinterpret_cast<uint64_t*>(after + 2 * H * sizeof(float));
  uint64_t* bar_x = bars;        // x + W1 landed
  uint64_t* bar_w2a = bars + 1;  // W2 k-block 0 landed
  uint64_t* bar_w2b = bars + 2;  // W2 k-blocks 1.. landed
  uint64_t* bar_g1 = bars + 3;   // GEMM 1 retired (region A reusable)
  uint64_t* bar_acc1 = bars + 4; // accumulator 1 ready
  uint64_t* bar_h1 = bars + 5;   // hidden tile written (4 warps)
  uint64_t* bar_acc2 = bars + 6; // accumulator 2 ready
  uint64_t* bar_p = bars + 7;    // pool mode: P tile + table landed
  uint64_t* bar_acc0 = bars + 8; // pool mode: accumulator 0 (x) ready
  uint64_t* bar_x0 = bars + 9;   // pool mode: x tile written (4 warps)
  uint64_t* bar_pw = bars + 10;  // pool mode from ids: P tile written (the 4 warps of the first column half)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* ss_s = reinterpret_cast<float*>(bars + 13);               // [2][128] per-half row sums of squares
  const int kV = p.V / 64;                                         // pool mode: k-blocks of GEMM 0
  uint8_t* p_tile = h1_tile;                                       // [128 x V] bf16, kV k-blocks of 16 KB
  uint8_t* t_tile = h1_tile + (uint32_t)kV * MLP_BM * 128;         // table, MN-major: E/64 boxes of [V rows x 128 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * MLP_BM;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    mbar_init(bar_x, 1); mbar_init(bar_w2a, 1); mbar_init(bar_w2b, 1); mbar_init(bar_g1, 1);
    mbar_init(bar_acc1, 1); mbar_init(bar_h1, 8); mbar_init(bar_acc2, 1);
    mbar_init(bar_p, 1); mbar_init(bar_acc0, 1); mbar_init(bar_x0, 8); mbar_init(bar_pw, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  pdl_wait();                                                      // barrier init / TMEM alloc overlapped the previous kernel
  if (warp >= 2) {                                                 // biases -> smem (read as broadcasts later)
    for (int i = threadIdx.x - 64; i < H; i += 256) { bias_s[i] = p.b1[i]; bias_s[H + i] = p.b2[i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a1 = tmem_base, tmem_a2 = tmem_base + 256;

  if (warp == 0) {
    if (elect_one()) {
      if (kV > 0) {
        mbar_arrive_expect_tx(bar_p, (p.ids ? 0u : (uint32_t)kV * MLP_BM * 128) + (uint32_t)p.V * E * 2);
        if (!p.ids)
          for (int kb = 0; kb < kV; ++kb) tma_load_2d(p_tile + kb * (MLP_BM * 128), &tmX, bar_p, kb * 64, (int)m0);
        for (int nb = 0; nb < kE; ++nb) tma_load_2d(t_tile + (uint32_t)nb * p.V * 128, &tmT, bar_p, nb * 64, 0);
        mbar_arrive_expect_tx(bar_x, w1_bytes);
      } else {
        mbar_arrive_expect_tx(bar_x, a_bytes);
        for (int kb = 0; kb < kE; ++kb) tma_load_2d(x_tile + kb * (MLP_BM * 128), &tmX, bar_x, kb * 64, (int)m0);
      }
      for (int kb = 0; kb < kE; ++kb) tma_load_2d(w1_tile + (uint32_t)kb * w2_blk, &tmW1, bar_x, kb * 64, 0);
      if (kH > 1) {
        mbar_arrive_expect_tx(bar_w2b, (uint32_t)(kH - 1) * w2_blk);
        for (int kb = 1; kb < kH; ++kb) tma_load_2d(w2_rest + (uint32_t)(kb - 1) * w2_blk, &tmW2, bar_w2b, kb * 64, 0);
      }
    }
    __syncwarp();
    mbar_wait(bar_g1, 0);                                        // GEMM 1 has finished reading x / W1
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w2a, w2_blk);
      tma_load_2d(base, &tmW2, bar_w2a, 0, 0);
    }
    __syncwarp();
  } else if (warp == 1) {
    // whole warp in uniform control flow; only the tcgen05 instructions are predicated on the elected lane
    const uint32_t idesc = umma_idesc_bf16(MLP_BM, H, 0, 0);
    const uint64_t dx = umma_desc_kmajor(smem_u32(x_tile), 0);
    const uint64_t dw1 = umma_desc_kmajor(smem_u32(w1_tile), 0);
    if (kV > 0) {                                                  // GEMM 0: x = P table  (A K-major, B = table read MN-major)
      const uint32_t idesc0 = umma_idesc_bf16(MLP_BM, E, 0, 1);
      const uint64_t dp = umma_desc_kmajor(smem_u32(p_tile), 0);
      const uint64_t dt = umma_desc_mnmajor(smem_u32(t_tile), 0, (uint32_t)p.V * 128);
      mbar_wait(bar_p, 0);
      if (p.ids) {                                                 // P built by the warps: also the saved pooling matrix of the backward
        mbar_wait(bar_pw, 0);
        if (lane == 0) {
          for (int kb = 0; kb < kV; ++kb) tma_store_2d(&tmX, p_tile + kb * (MLP_BM * 128), kb * 64, (int)m0);
          tma_store_commit();
        }
        __syncwarp();
      }
      tc_fence_after();
      for (int kb = 0; kb < kV; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (elect_one())
            umma_bf16(tmem_a2, dp + (uint64_t)(kb * (MLP_BM * 128 / 16) + k * 2), dt + (uint64_t)((kb * 4 + k) * 128), idesc0, (kb | k) != 0);
        }
      if (elect_one()) umma_commit(bar_acc0);
      __syncwarp();
      mbar_wait(bar_x0, 0);                                        // x tile written by the warps
    }
    mbar_wait(bar_x, 0);
    tc_fence_after();
    for (int kb = 0; kb < kE; ++kb)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (elect_one())
          umma_bf16(tmem_a1, dx + (uint64_t)(kb * (MLP_BM * 128 / 16) + k * 2),
                    dw1 + (uint64_t)(kb * (w2_blk >> 4) + k * 2), idesc, (kb | k) != 0);
      }
    if (kV > 0 && p.ids && lane == 0) tma_store_wait_read();    // the P tile (== hidden tile bytes) has been read out
    __syncwarp();
    if (elect_one()) { umma_commit(bar_g1); umma_commit(bar_acc1); }
    __syncwarp();
    const uint64_t dh = umma_desc_kmajor(smem_u32(h1_tile), 0);
    const uint64_t dw2a = umma_desc_kmajor(smem_u32(base), 0);
    const uint64_t dw2b = umma_desc_kmajor(smem_u32(w2_rest), 0);
    mbar_wait(bar_h1, 0);
    if (kH > 1) mbar_wait(bar_w2b, 0);
    tc_fence_after();
    // k-blocks 1.. of W2 have been resident since the start; k-block 0 (loaded late into the x/W1 bytes) goes last so
    // its TMA latency hides behind the other MMAs
    for (int it = 0; it < kH; ++it) {
      const int kb = (it + 1) % kH;
      if (kb == 0) { mbar_wait(bar_w2a, 0); tc_fence_after(); }
      const uint64_t dw = kb == 0 ? dw2a : dw2b + (uint64_t)((kb - 1) * (w2_blk >> 4));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (elect_one())
          umma_bf16(tmem_a2, dh + (uint64_t)(kb * (MLP_BM * 128 / 16) + k * 2), dw + (uint64_t)(k * 2), idesc, (it | k) != 0);
      }
    }
    if (elect_one()) umma_commit(bar_acc2);
    __syncwarp();
  } else {
    // epilogue warps 2..9: TMEM lane quarter = warp % 4, and the two warps of a quarter split the columns in halves
    // (two warps per scheduler hide each other's instruction latency; the epilogues are instruction-bound)
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int lrow = quarter * 32 + lane;
    const int64_t row = m0 + lrow;
    const bool row_ok = row < p.R;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int nce = E / 32, nch = H / 32;                          // 32-column chunks (both even)
    TT_MLP_STAMP(1);
    if (kV > 0 && p.ids && half == 0) {
      // ---- pooling matrix from the token ids: P[r,v] = count_r(v) / (len_r + 1e-9), bf16, straight into the swizzled
      // A tile of GEMM 0.  One THREAD per row (the 128 threads of the first column half): a private byte histogram in
      // shared memory (counts <= L <= 64; it borrows the x-tile bytes, nothing lives there before GEMM 0 has retired), no
      // atomics, no barriers, all 128 rows of the tile in flight at once.  Rows are rotated by 8 r bytes against each
      // other so that equal ids of different rows fall into different banks.
      const int V = p.V, L = p.L;
      const int r = lrow;
      uint8_t* h = x_tile + r * V;
      const int rot = (V & (V - 1)) == 0 ? (8 * r) & (V - 1) : 0;
      for (int i = 0; i < V; i += 16) *reinterpret_cast<uint4*>(h + i) = make_uint4(0u, 0u, 0u, 0u);
      int cnt = 0;
      if (row_ok) {
        auto count = [&](long long id) {
          if (id > 0 && id < V) { int pos = (int)id + rot; if (pos >= V) pos -= V; h[pos] = (uint8_t)(h[pos] + 1); ++cnt; }
          // nn.Embedding raises IndexError for an id outside [0, V) (embeddings.py:33-40): treated as padding, recorded for the host
          else if (id != 0 && p.bad_id) *reinterpret_cast<volatile long long*>(p.bad_id) = id * 2 + 1;
        };
        if (p.id_bytes == 4) {
          const int* q = static_cast<const int*>(p.ids) + row * L;
#pragma unroll 16
          for (int t = 0; t < L; ++t) count(__ldg(q + t));
        } else {
          const long long* q = static_cast<const long long*>(p.ids) + row * L;
#pragma unroll 16
          for (int t = 0; t < L; ++t) count(__ldg(q + t));
        }
      }
      const float il = 1.0f / ((float)cnt + 1e-9f);                 // encoders.py:72
      if (p.inv_len && row_ok) p.inv_len[row] = il;
      for (int c = 0; c < V / 8; ++c) {                             // 8 counts -> one 16-byte chunk of the swizzled tile
        int pos = 8 * c + rot; if (pos >= V) pos -= V;
        const uint2 w = *reinterpret_cast<const uint2*>(h + pos);
        const uint4 v = make_uint4(pack_bf16x2((float)(w.x & 255u) * il, (float)((w.x >> 8) & 255u) * il),
                                   pack_bf16x2((float)((w.x >> 16) & 255u) * il, (float)(w.x >> 24) * il),
                                   pack_bf16x2((float)(w.y & 255u) * il, (float)((w.y >> 8) & 255u) * il),
                                   pack_bf16x2((float)((w.y >> 16) & 255u) * il, (float)(w.y >> 24) * il));
        *reinterpret_cast<uint4*>(p_tile + (uint32_t)(c >> 3) * (MLP_BM * 128) + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pw);
    }
    if (kV > 0) {
      // ---- epilogue 0 (pool mode): x = acc0 -> bf16 -> swizzled x tile (A operand of GEMM 1) -----------------------
      mbar_wait(bar_acc0, 0);
      tc_fence_after();
      TT_MLP_STAMP(2);
      for (int c = half * (nce / 2); c < (half + 1) * (nce / 2); ++c) {
        uint32_t r[32];
        tmem_ld_x32(tmem_a2 + lane_addr + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        uint8_t* xrow = x_tile + (uint32_t)(c >> 1) * (MLP_BM * 128) + lrow * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 v = make_uint4(pack_bf16x2(__uint_as_float(r[8 * u]), __uint_as_float(r[8 * u + 1])),
                                     pack_bf16x2(__uint_as_float(r[8 * u + 2]), __uint_as_float(r[8 * u + 3])),
                                     pack_bf16x2(__uint_as_float(r[8 * u + 4]), __uint_as_float(r[8 * u + 5])),
                                     pack_bf16x2(__uint_as_float(r[8 * u + 6]), __uint_as_float(r[8 * u + 7])));
          *reinterpret_cast<uint4*>(xrow + ((((c & 1) * 4 + u) ^ (lrow & 7)) << 4)) = v;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_x0);
    }
    TT_MLP_STAMP(3);
    // ---- epilogue 1: hidden = relu(acc1 + b1) -> bf16 -> smem A tile + global h1 ------------------------------
    mbar_wait(bar_acc1, 0);
    tc_fence_after();
    TT_MLP_STAMP(4);
    for (int c = half * (nch / 2); c < (half + 1) * (nch / 2); ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem_a1 + lane_addr + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float a = fmaxf(__uint_as_float(r[j]) + bias_s[c * 32 + j], 0.f);
        const float b = fmaxf(__uint_as_float(r[j + 1]) + bias_s[c * 32 + j + 1], 0.f);
        pk[j >> 1] = pack_bf16x2(a, b);
      }
      uint8_t* hrow = h1_tile + (uint32_t)(c >> 1) * (MLP_BM * 128) + lrow * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ch = (c & 1) * 4 + u;
        const uint4 v = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        *reinterpret_cast<uint4*>(hrow + ((ch ^ (lrow & 7)) << 4)) = v;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_h1);
    asm volatile("bar.sync 2, 256;" ::: "memory");               // both column halves of every row are in the tile
    if (half == 0 && lane == 0) {
      // the same swizzled tile GEMM 2 reads is also the source of the global h1 write: one TMA store per k-block of
      // this quarter's 32 rows (rows past R are clipped) instead of 1 KB-strided per-thread stores
      for (int kb = 0; kb < kH; ++kb)
        tma_store_2d(&tmH1, h1_tile + (uint32_t)kb * (MLP_BM * 128) + quarter * 4096, kb * 64, (int)m0 + quarter * 32);
      tma_store_commit();
    }
    TT_MLP_STAMP(5);
    // ---- epilogue 2: z = acc2 + b2; y = z / max(|z|, 1e-12) --------------------------------------------------
    mbar_wait(bar_acc2, 0);
    tc_fence_after();
    TT_MLP_STAMP(6);
    float ss = 0.f;
    for (int c = half * (nch / 2); c < (half + 1) * (nch / 2); ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem_a2 + lane_addr + (uint32_t)(c * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float v = __uint_as_float(r[j]) + bias_s[H + c * 32 + j];
        ss = fmaf(v, v, ss);
      }
    }
    ss_s[half * MLP_BM + lrow] = ss;
    if (half == 0 && lane == 0) tma_store_wait_read();           // the hidden tile has been read out by its TMA stores:
    asm volatile("bar.sync 2, 256;" ::: "memory");               // it now stages the bf16 y tile
    const float inv = 1.0f / fmaxf(sqrtf(ss_s[lrow] + ss_s[MLP_BM + lrow]), 1e-12f);
    if (half == 0 && p.inv_norm && row_ok) p.inv_norm[row] = inv;
    for (int c = half * (nch / 2); c < (half + 1) * (nch / 2); ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem_a2 + lane_addr + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      float zv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) zv[j] = __uint_as_float(r[j]) + bias_s[H + c * 32 + j];
      if (p.yb) {
        uint8_t* yrow = h1_tile + (uint32_t)(c >> 1) * (MLP_BM * 128) + lrow * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 v = make_uint4(pack_bf16x2(zv[8 * u] * inv, zv[8 * u + 1] * inv), pack_bf16x2(zv[8 * u + 2] * inv, zv[8 * u + 3] * inv),
                                     pack_bf16x2(zv[8 * u + 4] * inv, zv[8 * u + 5] * inv), pack_bf16x2(zv[8 * u + 6] * inv, zv[8 * u + 7] * inv));
          *reinterpret_cast<uint4*>(yrow + ((((c & 1) * 4 + u) ^ (lrow & 7)) << 4)) = v;
        }
      }
      if (row_ok) {
        if (p.z) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(p.z + row * H + c * 32 + j) = make_float4(zv[j], zv[j + 1], zv[j + 2], zv[j + 3]);
        }
        if (p.y) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(p.y + row * H + c * 32 + j) =
                make_float4(zv[j] * inv, zv[j + 1] * inv, zv[j + 2] * inv, zv[j + 3] * inv);
        }
      }
    }
    if (p.yb) {
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (half == 0 && lane == 0) {
        for (int kb = 0; kb < kH; ++kb)
          tma_store_2d(&tmY, h1_tile + (uint32_t)kb * (MLP_BM * 128) + quarter * 4096, kb * 64, (int)m0 + quarter * 32);
        tma_store_commit();
      }
    }
    TT_MLP_STAMP(7);
    if (half == 0 && lane == 0) tma_store_wait();                // smem stays valid until every bulk store has completed
    TT_MLP_STAMP(8);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static size_t mlp_fused_smem(int E, int H) {
  const int kE = E / 64, kH = H / 64;
  return 1024 + (size_t)kE * MLP_BM * 128 + (size_t)kE * H * 128 + (size_t)(kH - 1) * H * 128 + (size_t)kH * MLP_BM * 128 +
         2 * (size_t)H * 4 + 13 * 8 + 16 + 2 * MLP_BM * 4;
}

}  // namespace tc

// shapes the fused kernel tiles: k-blocks of 64, W2 k-block 0 must fit in the x/W1 bytes, 227 KB of shared memory
bool tc_mlp_fused_supported(int E, int H) {
  if (E % 64 != 0 || H % 64 != 0 || E < 64 || H < 64 || H > 256) return false;
  const int kE = E / 64;
  if ((size_t)kE * tc::MLP_BM * 128 + (size_t)kE * H * 128 < (size_t)H * 128) return false;
  return tc::mlp_fused_smem(E, H) <= 227 * 1024;
}

// x = P table inside the kernel: V a multiple of 64, P tile + table fit the hidden-tile bytes, E <= 256 accumulator columns
bool tc_mlp_fwd_pool_supported(int E, int H, int64_t V) {
  if (!tc_mlp_fused_supported(E, H) || V < 64 || V % 64 != 0 || V > 256) return false;   // TMA box: <= 256 rows
  return (size_t)(V / 64) * tc::MLP_BM * 128 + (size_t)V * E * 2 <= (size_t)(H / 64) * tc::MLP_BM * 128;
}

int tc_mlp_fwd_fused(const __nv_bfloat16* xb, const __nv_bfloat16* w1b, const float* b1, const __nv_bfloat16* w2b,
                     const float* b2, int64_t R, int E, int H, __nv_bfloat16* h1b, float* z, float* y,
                     __nv_bfloat16* yb, float* inv_norm, const __nv_bfloat16* pool, int64_t V, const __nv_bfloat16* table_bf16,
                     cudaStream_t s, const void* ids, int id_bytes, int L, float* inv_len) {
  CUtensorMap tmX, tmW1, tmW2, tmH1, tmY, tmT;
  int rc;
  if (pool) {
    if (!tc_mlp_fwd_pool_supported(E, H, V) || !table_bf16) { set_error("tc_mlp_fwd: in-kernel x = P table needs V %% 64 == 0 (<= 256), the bf16 table and a tile that fits"); return TT_ERR_UNSUPPORTED; }
    rc = tc::make_tmap_bf16(&tmX, pool, (uint64_t)R, (uint64_t)V, tc::MLP_BM); if (rc) return rc;
    rc = tc::make_tmap_bf16(&tmT, table_bf16, (uint64_t)V, (uint64_t)E, (uint32_t)V); if (rc) return rc;
  } else {
    rc = tc::make_tmap_bf16(&tmX, xb, (uint64_t)R, (uint64_t)E, tc::MLP_BM); if (rc) return rc;
    tmT = tmX;
  }
  rc = tc::make_tmap_bf16(&tmW1, w1b, (uint64_t)H, (uint64_t)E, (uint32_t)H); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmW2, w2b, (uint64_t)H, (uint64_t)H, (uint32_t)H); if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmH1, h1b, (uint64_t)R, (uint64_t)H, 32); if (rc) return rc;          // store maps: 32-row boxes (one warp)
  rc = tc::make_tmap_bf16(&tmY, yb ? yb : h1b, (uint64_t)R, (uint64_t)H, 32); if (rc) return rc;
  tc::MlpFwdParams p{};
  p.R = R; p.E = E; p.H = H; p.b1 = b1; p.b2 = b2; p.h1b = h1b; p.z = z; p.y = y; p.yb = yb; p.inv_norm = inv_norm;
  p.V = pool ? (int)V : 0;
  if (ids) {
    // a byte histogram per row borrows the x-tile bytes: 128 rows x V bytes must fit E/64 k-blocks of 16 KB; counts <= 255
    if (!pool || L < 1 || L > 255 || (id_bytes != 4 && id_bytes != 8) || (size_t)tc::MLP_BM * V > (size_t)(E / 64) * tc::MLP_BM * 128) {
      set_error("tc_mlp_fwd: building P from ids needs the pool mode, L <= 255 and V <= 128 E / 64"); return TT_ERR_UNSUPPORTED;
    }
    p.ids = ids; p.id_bytes = id_bytes; p.L = L; p.inv_len = inv_len; p.bad_id = bad_id_word();
  }
  const size_t smem = tc::mlp_fused_smem(E, H);
  TT_CUDA(cudaFuncSetAttribute(tc::tc_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const bool dbg_on = getenv("TT_MLP_DEBUG") != nullptr;
  long long* dbg_dev = nullptr;
  if (dbg_on) { cudaMalloc(&dbg_dev, 16 * sizeof(long long)); cudaMemset(dbg_dev, 0, 16 * sizeof(long long)); p.dbg = dbg_dev; }
  TT_CUDA(launch_kernel(tc::tc_mlp_fwd_kernel, dim3((unsigned)ceil_div(R, tc::MLP_BM)), dim3(tc::MLP_THREADS), smem, s, true, tmX, tmW1, tmW2, tmH1, tmY, tmT, p));
  TT_LAUNCH_CHECK("tc_mlp_fwd_kernel");
  if (dbg_on) {
    long long h[16];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    printf("[tt tc_mlp_fwd CTA 0, ns] start 0 | warps ready %lld | acc0 %lld | x written %lld | acc1 %lld | h1 written %lld | acc2 %lld | y staged %lld | stores done %lld\n",
           h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0]);
  }
  return TT_OK;
}

}  // namespace tt
