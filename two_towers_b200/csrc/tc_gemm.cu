// tc_gemm.cu -- bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM,
// operands staged by TMA into 128B-swizzled shared memory) and the tower MLP built on it
// (TT_PREC_BF16 path of K3).
//
//   C[M,N] = epilogue( A * B )   A: [M,K] (K-major) or stored [K,M] (MN-major)
//                                B: [N,K] (K-major, nn.Linear weight layout) or stored [K,N] (MN-major)
//   epilogue: + bias[n] -> relu -> * (mask[m,n] > 0); fp32 and/or bf16 output; or raw split-K partials.
//
// One CTA = one 128 x BN output tile (x one K split).  Warp roles (192 threads):
//   warp 0  TMA producer   : 4-stage ring of {A 128x64, B BNx64} bf16 tiles, mbarrier expect_tx
//   warp 1  MMA issuer     : one lane issues 4 x tcgen05.mma (K=16) per stage, tcgen05.commit
//                            frees the stage and finally signals the accumulator
//   warps 2-5 epilogue     : tcgen05.ld 32 lanes x 32 columns -> registers -> global
// Both operand majors are supported because the backward GEMMs (dW = dY^T X, dX = dY W) consume
// the same row-major activations with the roles of the axes swapped; no transposes are
// materialised.
#include <stdlib.h>
#include <vector>

#include "tc_common.cuh"
#include "sgemm.cuh"
#include "tensor_core.cuh"
#include "reduce.cuh"

namespace tt {
namespace tc {

// ---------------------------------------------------------------------------------------
// host: tensor map encode through the runtime's driver entry point
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || p == nullptr) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint64_t pitch) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return TT_ERR_CUDA; }
  if (pitch == 0) pitch = cols;                          // elements between rows (>= cols; columns past `cols` read as zero)
  if (pitch < cols || (pitch * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA operand needs 16-byte aligned base and row pitch (cols=%llu pitch=%llu)", (unsigned long long)cols, (unsigned long long)pitch);
    return TT_ERR_UNSUPPORTED;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows); return TT_ERR_CUDA; }
  return TT_OK;
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;
  int kblocks_per_split;      // 64-wide K blocks handled by one blockIdx.z
  float* C;                   // fp32 output [M,ldc] (nullable) or split partials [z][M][N]
  __nv_bfloat16* Cb;          // bf16 output [M,ldc] (nullable)
  int ldc;
  const float* bias;
  int act;
  const __nv_bfloat16* mask;  // nullable bf16 [M,ldmask]: out = mask > 0 ? acc : 0 (only without bias / act)
  int ldmask;
  int splits;
  int tiles_m, tiles_n, nblocks;   // block decomposition of this problem inside a (possibly shared) launch
  float* colsum_part;         // nullable [ceil(M/32), N]: per-32-row column sums of the epilogue output (bias grads)
  long long* dbg;             // developer aid (TT_GEMM_DEBUG=1): [cta][4] %globaltimer stamps
};

constexpr int kStages = 3;                 // 3 x 32 KB: two CTAs fit one SM and overlap each other's prologue / epilogue
constexpr int kGemmThreads = 192;         // TMA warp, MMA warp, 4 epilogue warps (one per TMEM lane quarter)
constexpr int BM = 128, BK = 64;
constexpr uint32_t kATile = BM * BK * 2;                         // 16 KB

// Up to two independent GEMM problems share one launch (e.g. {dW2 = dz^T h1, da1 = dz W2} in the tower
// backward): CTAs [0, p0.nblocks) work on problem 0, the rest on problem 1.
template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0, const GemmParams p0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, const GemmParams p1) {
  pdl_trigger();
  long long* cta_dbg = p0.dbg ? p0.dbg + 4 * blockIdx.x : nullptr;
  long long* fine = (p0.dbg && (blockIdx.x == 0 || (int)blockIdx.x == p0.nblocks)) ? p0.dbg + 4 * (gridDim.x + (blockIdx.x ? 64 : 0)) : nullptr;
#define TT_FINE(c, slot) do { if (fine && threadIdx.x == 64) fine[(c) * 8 + (slot)] = clock64(); } while (0)
#define TT_GEMM_STAMP(slot) do { if (cta_dbg && threadIdx.x == 64) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); cta_dbg[slot] = t_; } } while (0)
  TT_GEMM_STAMP(0);
  const bool second = (int)blockIdx.x >= p0.nblocks;
  const GemmParams& p = second ? p1 : p0;
  const CUtensorMap& tmA = second ? tmA1 : tmA0;
  const CUtensorMap& tmB = second ? tmB1 : tmB0;
  const int lin = (int)blockIdx.x - (second ? p0.nblocks : 0);
  const int bx = lin % p.tiles_n, by = (lin / p.tiles_n) % p.tiles_m, bz = lin / (p.tiles_n * p.tiles_m);
  constexpr uint32_t kBTile = BN * BK * 2;
  constexpr uint32_t kStageBytes = kATile + kBTile;
  extern __shared__ __align__(1024) uint8_t smem[];
  // carve: [stages x (A,B)] [full barriers][empty barriers][tmem_full][tmem slot]
  uint8_t* tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);       // stays in the shared address space
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  uint8_t* stage_tiles = tiles;   // 4 x [32][36] fp32 epilogue staging: reuses pipeline stage 0 (idle once the accumulator is complete)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = by * BM, n0 = bx * BN;
  const int total_kb = (p.K + BK - 1) / BK;
  const int kb_beg = bz * p.kblocks_per_split;
  const int kb_end = min(total_kb, kb_beg + p.kblocks_per_split);
  const int nkb = max(0, kb_end - kb_beg);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  pdl_wait();                                             // everything above overlapped the previous kernel's tail
  TT_GEMM_STAMP(1);

  if (warp == 0) {
    for (int i = 0; i < nkb; ++i) {                       // whole warp, uniform control flow; one lane issues
      const int s = i % kStages, k0 = (kb_beg + i) * BK;
      mbar_wait(&empty_bar[s], ((i / kStages) & 1) ^ 1);
      uint8_t* a = tiles + s * kStageBytes;
      uint8_t* b = a + kATile;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
        if (!p.a_mn) tma_load_2d(a, &tmA, &full_bar[s], k0, m0);
        else { tma_load_2d(a, &tmA, &full_bar[s], m0, k0); tma_load_2d(a + 8192, &tmA, &full_bar[s], m0 + 64, k0); }
        if (!p.b_mn) tma_load_2d(b, &tmB, &full_bar[s], k0, n0);
        else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(b + j * 8192, &tmB, &full_bar[s], n0 + 64 * j, k0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // This warp shares its scheduler with epilogue warps (of this CTA and of the co-resident one): the issuing lane is
    // elected once and the stage-0 descriptors are built once; a later stage only shifts their start-address field.
    const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
    const bool leader = elect_one();
    const uint32_t a0 = smem_u32(tiles), b0 = a0 + kATile;
    const uint64_t da_base = p.a_mn ? umma_desc_mnmajor(a0, 0, 8192u) : umma_desc_kmajor(a0, 0);
    const uint64_t db_base = p.b_mn ? umma_desc_mnmajor(b0, 0, 8192u) : umma_desc_kmajor(b0, 0);
    const uint64_t sa = p.a_mn ? 128u : 2u, sb = p.b_mn ? 128u : 2u;       // (bytes per k16 step) >> 4
    for (int i = 0; i < nkb; ++i) {
      const int s = i % kStages;
      mbar_wait(&full_bar[s], (i / kStages) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t off = (uint64_t)((uint32_t)s * kStageBytes >> 4);   // (address >> 4) lives in the low 14 bits
        const uint64_t da0 = da_base + off, db0 = db_base + off;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_acc, da0 + sa * k, db0 + sb * k, idesc, (i | k) != 0);
        umma_commit(&empty_bar[s]);                       // stage reusable once these MMAs have read it
      }
      __syncwarp();
    }
    if (leader) umma_commit(acc_bar);                     // accumulator complete
    __syncwarp();
  } else {
    // epilogue warps 2..5 -> TMEM lane quarters (warp % 4).  Each thread holds one accumulator row;
    // a 32x32 chunk is transposed through a warp-private shared-memory tile so that global stores
    // (and the bias / mask loads) are coalesced: lane == column, 128 contiguous bytes per row.
    const int quarter = warp & 3;
    float (*T)[36] = reinterpret_cast<float (*)[36]>(stage_tiles + (warp - 2) * (32 * 36 * 4));
    const bool split = p.splits > 1;
    float* outp = split ? p.C + (size_t)bz * p.M * p.N : p.C;
    const int ldo = split ? p.N : p.ldc;
    const int row0 = m0 + quarter * 32;
    // everything the loop needs, in registers (p is a run-time choice between two parameter blocks)
    const int N = p.N;
    const float* bias = split ? nullptr : p.bias;
    const bool has_bias = bias != nullptr, relu = !split && p.act == 1;
    __nv_bfloat16* Cb = split ? nullptr : p.Cb;
    const int ldcb = p.ldc;
    float* csp = p.colsum_part;
    const bool vec_ok = (N & 3) == 0 && (ldo & 3) == 0 && (ldcb & 3) == 0 && (reinterpret_cast<uintptr_t>(outp) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(Cb) & 7) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(csp) & 15) == 0;
    // ReLU-gate mask for this thread's accumulator row (32 bf16 = 64 contiguous bytes, two full sectors), fetched
    // one chunk ahead: the first before the accumulator is even complete, the next while a chunk is processed
    const bool use_mask = !split && p.mask != nullptr;
    uint4 mk_next[4];
    auto load_mask = [&](int c, uint4* mk) {
      if (use_mask && c < BN / 32) {
        const int grow = row0 + lane;
        const bool ok = grow < p.M && (n0 + c * 32 + 32 <= p.N) && ((p.ldmask & 7) == 0);
        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + (size_t)(ok ? grow : 0) * p.ldmask + n0 + c * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) mk[u] = ok ? __ldg(mp + u) : make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        if (!ok && grow < p.M) {                               // ragged / unaligned tail: element-wise
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const int c0 = n0 + c * 32 + j;
            const uint32_t lo = (c0 < p.N) ? (uint32_t)__bfloat16_as_ushort(p.mask[(size_t)grow * p.ldmask + c0]) : 0x3f80u;
            const uint32_t hi = (c0 + 1 < p.N) ? (uint32_t)__bfloat16_as_ushort(p.mask[(size_t)grow * p.ldmask + c0 + 1]) : 0x3f80u;
            w[j >> 1] = lo | (hi << 16);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) mk[u] = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
        }
      }
    };
    load_mask(0, mk_next);
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    TT_GEMM_STAMP(2);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      TT_FINE(c, 0);
      uint4 mk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) mk[u] = mk_next[u];
      load_mask(c + 1, mk_next);
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld_x32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        TT_FINE(c, 1);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (use_mask) {
        const uint32_t* mw = reinterpret_cast<const uint32_t*>(mk);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          // bf16 > 0  <=>  sign bit clear and magnitude non-zero
          const uint32_t w = mw[j >> 1];                      // bf16 > 0  <=>  its bits, as a signed integer, are > 0
          r[j] = (int)(w << 16) > 0 ? r[j] : 0u;
          r[j + 1] = (int)(w & 0xffff0000u) > 0 ? r[j + 1] : 0u;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(&T[lane][j]) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      __syncwarp();
      TT_FINE(c, 2);
      const int nrows = min(32, p.M - row0);
      if (vec_ok) {
        // fast path: lane -> 4 consecutive columns of row (4 k + lane / 8); one instruction stores 4 rows x 128 B.
        // Straight-line code: with a single epilogue warp per scheduler, instruction-level parallelism is all there is.
        const int cq = (lane & 7) * 4, rq = lane >> 3;
        const int col0 = n0 + c * 32 + cq;
        const bool cok = col0 < N;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias && cok) bv = *reinterpret_cast<const float4*>(bias + col0);
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        float* o32 = outp ? outp + (size_t)(row0 + rq) * ldo + col0 : nullptr;
        __nv_bfloat16* o16 = Cb ? Cb + (size_t)(row0 + rq) * ldcb + col0 : nullptr;
        if (nrows == 32 && n0 + c * 32 + 32 <= N) {         // whole chunk in range (warp-uniform): no predicates at all
          float4 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const float4*>(&T[4 * k + rq][cq]);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            v[k].x += bv.x; v[k].y += bv.y; v[k].z += bv.z; v[k].w += bv.w;
            if (relu) { v[k].x = fmaxf(v[k].x, 0.f); v[k].y = fmaxf(v[k].y, 0.f); v[k].z = fmaxf(v[k].z, 0.f); v[k].w = fmaxf(v[k].w, 0.f); }
            cs.x += v[k].x; cs.y += v[k].y; cs.z += v[k].z; cs.w += v[k].w;
          }
          if (o32) {
#pragma unroll
            for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(o32 + (size_t)(4 * k) * ldo) = v[k];
          }
          if (o16) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              *reinterpret_cast<uint2*>(o16 + (size_t)(4 * k) * ldcb) = make_uint2(pack_bf16x2(v[k].x, v[k].y), pack_bf16x2(v[k].z, v[k].w));
          }
        } else
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 v = *reinterpret_cast<const float4*>(&T[4 * k + rq][cq]);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          const bool ok = cok && (4 * k + rq < nrows);
          if (ok) {
            cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
            if (o32) *reinterpret_cast<float4*>(o32 + (size_t)(4 * k) * ldo) = v;
            if (o16) *reinterpret_cast<uint2*>(o16 + (size_t)(4 * k) * ldcb) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
          }
        }
        if (csp) {                                         // 32-row column sums: fold the 4 row groups (lanes +8, +16)
#pragma unroll
          for (int d = 8; d <= 16; d <<= 1) {
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, d); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, d);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, d); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, d);
          }
          if (rq == 0 && cok && nrows > 0) *reinterpret_cast<float4*>(csp + (size_t)(row0 >> 5) * N + col0) = cs;
        }
        __syncwarp();
        TT_FINE(c, 7);
        continue;
      }
      const int col = n0 + c * 32 + lane;
      const bool col_ok = col < p.N;
      const float bias_v = (!split && p.bias && col_ok) ? p.bias[col] : 0.f;
      // generic path (unaligned leading dimensions): 8 rows at a time, 128-byte coalesced scalar stores
      float csum = 0.f;
#pragma unroll 1
      for (int rb = 0; rb < 32; rb += 8) {
        if (rb >= nrows) break;
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = T[rb + u][lane];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (!split) {
            float x = v[u] + bias_v;
            if (p.act == 1) x = fmaxf(x, 0.f);
            v[u] = x;
          }
          if (rb + u < nrows) {
            csum += v[u];
            if (col_ok) {
              if (outp) outp[(size_t)(row0 + rb + u) * ldo + col] = v[u];
              if (!split && p.Cb) p.Cb[(size_t)(row0 + rb + u) * p.ldc + col] = __float2bfloat16(v[u]);
            }
          }
        }
      }
      if (p.colsum_part && col_ok && nrows > 0) p.colsum_part[(size_t)(row0 >> 5) * p.N + col] = csum;
      __syncwarp();
      TT_FINE(c, 7);
    }
  }
  tc_fence_before();
  __syncthreads();
  TT_GEMM_STAMP(3);
  if (warp == 1) tmem_dealloc(tmem_acc, BN);
}

template <int BN>
static constexpr size_t gemm_smem_bytes() {
  return 1024 + kStages * (kATile + BN * BK * 2) + (2 * kStages + 1) * 8 + 16;
}

struct TcGemm {
  int M, N, K;
  const __nv_bfloat16* A; int a_mn;      // a_mn ? stored [K,M] : stored [M,K]
  const __nv_bfloat16* B; int b_mn;      // b_mn ? stored [K,N] : stored [N,K]
  int lda = 0, ldb = 0;                  // row pitch of the stored matrices in elements (0: contiguous); multiple of 8
  float* C = nullptr; __nv_bfloat16* Cb = nullptr; int ldc = 0;
  const float* bias = nullptr; int act = 0; const __nv_bfloat16* mask = nullptr; int ldmask = 0;
  int splits = 1; float* partial = nullptr;
  bool defer_reduce = false;             // split-K: leave the partials, the caller reduces them later
  float* colsum_part = nullptr;          // [ceil(M/32), N] column-sum partials of the output
};

template <int BN>
static int fill_problem(const TcGemm& g, CUtensorMap* tmA, CUtensorMap* tmB, GemmParams* pp) {
  int rc;
  if (!g.a_mn) rc = make_tmap_bf16(tmA, g.A, (uint64_t)g.M, (uint64_t)g.K, BM, (uint64_t)g.lda);
  else         rc = make_tmap_bf16(tmA, g.A, (uint64_t)g.K, (uint64_t)g.M, 64, (uint64_t)g.lda);
  if (rc) return rc;
  if (!g.b_mn) rc = make_tmap_bf16(tmB, g.B, (uint64_t)g.N, (uint64_t)g.K, BN, (uint64_t)g.ldb);
  else         rc = make_tmap_bf16(tmB, g.B, (uint64_t)g.K, (uint64_t)g.N, 64, (uint64_t)g.ldb);
  if (rc) return rc;
  GemmParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.a_mn = g.a_mn; p.b_mn = g.b_mn;
  const int total_kb = (g.K + BK - 1) / BK;
  p.kblocks_per_split = (total_kb + g.splits - 1) / g.splits;
  p.splits = g.splits;
  p.C = g.splits > 1 ? g.partial : g.C;
  p.Cb = g.Cb; p.ldc = g.ldc; p.bias = g.bias; p.act = g.act; p.mask = g.mask; p.ldmask = g.ldmask;
  p.colsum_part = g.colsum_part;
  p.tiles_n = (int)ceil_div(g.N, BN); p.tiles_m = (int)ceil_div(g.M, BM);
  p.nblocks = p.tiles_n * p.tiles_m * g.splits;
  *pp = p;
  return TT_OK;
}

static int finish_problem(const TcGemm& g, cudaStream_t s) {
  if (g.splits > 1 && !g.defer_reduce) {
    // fixed-order reduction + epilogue (shared with the fp32 path)
    SgemmArgs a{};
    a.M = g.M; a.N = g.N; a.K = g.K; a.C = g.C; a.ldc = g.ldc; a.bias = g.bias; a.act = g.act;
    a.splits = g.splits; a.partial = g.partial;
    return splitk_reduce(a, s);
  }
  return TT_OK;
}

// g1 may be null (single problem).  Both problems must use the same tile width BN.
template <int BN>
static int launch_gemm_pair(const TcGemm& g0, const TcGemm* g1, cudaStream_t s) {
  CUtensorMap tmA0, tmB0, tmA1, tmB1;
  GemmParams p0{}, p1{};
  int rc = fill_problem<BN>(g0, &tmA0, &tmB0, &p0); if (rc) return rc;
  if (g1) { rc = fill_problem<BN>(*g1, &tmA1, &tmB1, &p1); if (rc) return rc; }
  else { tmA1 = tmA0; tmB1 = tmB0; p1 = p0; p1.nblocks = 0; }
  constexpr size_t smem = gemm_smem_bytes<BN>();
  TT_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TT_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  static const bool dbg_on = getenv("TT_GEMM_DEBUG") != nullptr;
  const int ncta = p0.nblocks + p1.nblocks;
  long long* dbg_dev = nullptr;
  const size_t dbg_n = (size_t)ncta * 4 + 128;
  if (dbg_on) { cudaMalloc(&dbg_dev, dbg_n * sizeof(long long)); cudaMemset(dbg_dev, 0, dbg_n * sizeof(long long)); p0.dbg = dbg_dev; }
  TT_CUDA(launch_kernel(tc_gemm_kernel<BN>, dim3((unsigned)(p0.nblocks + p1.nblocks)), dim3(kGemmThreads), smem, s, true, tmA0, tmB0, p0,
                        tmA1, tmB1, p1));
  TT_LAUNCH_CHECK("tc_gemm_kernel");
  if (dbg_on) {
    std::vector<long long> h(dbg_n);
    cudaStreamSynchronize(s);
    cudaMemcpy(h.data(), dbg_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbg_dev);
    long long g0 = 0;
    for (int i = 0; i < ncta; ++i) if (h[4 * i] && (!g0 || h[4 * i] < g0)) g0 = h[4 * i];
    printf("[tt tc_gemm<%d> per-CTA ns] problems {M %d N %d K %d splits %d : %d ctas} {M %d N %d K %d splits %d : %d ctas}\n", BN, p0.M, p0.N, p0.K,
           p0.splits, p0.nblocks, p1.M, p1.N, p1.K, p1.splits, p1.nblocks);
    for (int i = 0; i < ncta; i += (ncta > 64 ? 7 : 1))
      printf("  cta %3d: start %6lld  ready %6lld  acc %6lld  end %6lld\n", i, h[4 * i] - g0, h[4 * i + 1] - g0, h[4 * i + 2] - g0, h[4 * i + 3] - g0);
    long long mx = 0; for (int i = 0; i < ncta; ++i) if (h[4 * i + 3] - g0 > mx) mx = h[4 * i + 3] - g0;
    printf("  last end %lld ns\n", mx);
    for (int pr = 0; pr < 2; ++pr) {
      const long long* f = h.data() + (size_t)ncta * 4 + pr * 64;
      printf("  problem %d first CTA, epilogue cycles per 32-col chunk {begin, tmem, staged, lds0, lds1, lds2, lds3, end}:\n", pr);
      for (int c = 0; c < BN / 32; ++c) { printf("    c%d:", c); for (int k = 0; k < 8; ++k) printf(" %6lld", f[c * 8 + k] ? f[c * 8 + k] - f[0] : -1); printf("\n"); }
    }
  }
  rc = finish_problem(g0, s); if (rc) return rc;
  if (g1) { rc = finish_problem(*g1, s); if (rc) return rc; }
  return TT_OK;
}

static int check_problem(const TcGemm& g) {
  TT_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, "tc_gemm: bad shape");
  TT_CHECK_ARG(g.splits == 1 || (g.partial && g.C && !g.Cb && !g.mask), "tc_gemm: split-K needs partial + fp32 output only");
  TT_CHECK_ARG(!g.mask || (!g.bias && g.act == 0), "tc_gemm: the gate mask is applied to the raw accumulator (no bias / act)");
  return TT_OK;
}

int tc_gemm(const TcGemm& g, cudaStream_t s) {
  int rc = check_problem(g); if (rc) return rc;
  if (g.N <= 64) return launch_gemm_pair<64>(g, nullptr, s);
  return launch_gemm_pair<128>(g, nullptr, s);            // 128-wide tiles: twice the CTAs of 256-wide ones for N = 256
}

// two independent problems in one launch (same tile-width class)
int tc_gemm2(const TcGemm& g0, const TcGemm& g1, cudaStream_t s) {
  int rc = check_problem(g0); if (rc) return rc;
  rc = check_problem(g1); if (rc) return rc;
  const bool n0 = g0.N <= 64, n1 = g1.N <= 64;
  if (n0 != n1) { rc = tc_gemm(g0, s); if (rc) return rc; return tc_gemm(g1, s); }
  if (n0) return launch_gemm_pair<64>(g0, &g1, s);
  return launch_gemm_pair<128>(g0, &g1, s);
}

static int pick_splits(int M, int N, int K) {
  const int bn = N <= 64 ? 64 : 128;
  const int64_t tiles = ceil_div(M, BM) * ceil_div(N, bn);
  const int total_kb = (K + BK - 1) / BK;
  if (tiles >= kNumSMs / 2 || total_kb <= 4) return 1;
  int64_t want = ceil_div(kNumSMs, tiles);
  int64_t maxs = total_kb / 8;              // >= 8 k-blocks (512 of K) per split: halves the partial-slice traffic
  if (want > maxs) want = maxs;
  if (want > 64) want = 64;
  return (int)(want < 1 ? 1 : want);
}

// ---------------------------------------------------------------------------------------
// small elementwise helpers of the bf16 path
// ---------------------------------------------------------------------------------------
__global__ void cast2_kernel(const float* __restrict__ a, __nv_bfloat16* __restrict__ ab, int64_t na,
                             const float* __restrict__ b, __nv_bfloat16* __restrict__ bb, int64_t nb,
                             const float* __restrict__ c, __nv_bfloat16* __restrict__ cb, int64_t nc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < na; i += stride) ab[i] = __float2bfloat16(a[i]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += stride) bb[i] = __float2bfloat16(b[i]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nc; i += stride) cb[i] = __float2bfloat16(c[i]);
}

static int cast3(const float* a, __nv_bfloat16* ab, int64_t na, const float* b, __nv_bfloat16* bb, int64_t nb,
                 const float* c, __nv_bfloat16* cb, int64_t nc, cudaStream_t s) {
  int64_t n = na > nb ? na : nb;
  if (nc > n) n = nc;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  cast2_kernel<<<(unsigned)blocks, 256, 0, s>>>(a, ab, na, b, bb, nb, c, cb, nc);
  TT_LAUNCH_CHECK("cast2_kernel");
  return TT_OK;
}

int cast3_public(const float* a, __nv_bfloat16* ab, int64_t na, const float* b, __nv_bfloat16* bb, int64_t nb, cudaStream_t s) {
  return cast3(a, ab, na, b, bb, nb, nullptr, nullptr, 0, s);
}

// fp32 [rows, cols] -> bf16 [rows, pitch] (pitch >= cols, multiple of 8; the pad columns are zeroed).  Lets operands whose
// inner dimension is not a multiple of 8 (the word tower's E = 300, configs/word2vec_skipgram.yml:23) meet TMA's 16-byte
// row-pitch rule: the tensor map keeps the true column count, so the pad never enters a product.
__global__ void __launch_bounds__(256)
cast_rows_kernel(const float* __restrict__ src, int64_t rows, int cols, int pitch, __nv_bfloat16* __restrict__ dst) {
  const int64_t n = rows * (int64_t)pitch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / pitch;
    const int c = (int)(i - r * pitch);
    dst[i] = __float2bfloat16(c < cols ? src[r * cols + c] : 0.f);
  }
}
static int cast_rows(const float* src, int64_t rows, int cols, int pitch, __nv_bfloat16* dst, cudaStream_t s) {
  int64_t blocks = ceil_div(rows * (int64_t)pitch, 256);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  cast_rows_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, rows, cols, pitch, dst);
  TT_LAUNCH_CHECK("cast_rows_kernel");
  return TT_OK;
}

// y = z / max(|z|,1e-12) (fp32 + bf16)
__global__ void __launch_bounds__(256)
l2norm_fwd_bf16_kernel(const float* __restrict__ z, int64_t R, int H, float* __restrict__ y, __nv_bfloat16* __restrict__ yb,
                       float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* zr = z + row * H;
  float ss = 0.f;
  for (int e = lane; e < H; e += 32) { float v = zr[e]; ss = fmaf(v, v, ss); }
  const float denom = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  if (inv_norm && lane == 0) inv_norm[row] = 1.0f / denom;
  for (int e = lane; e < H; e += 32) {
    const float o = zr[e] / denom;
    if (y) y[row * H + e] = o;
    if (yb) yb[row * H + e] = __float2bfloat16(o);
  }
}
// Tail of the embedding-fused tower backward: with M = P^T da1 [V,H] (fp32),
//   dw1[h,e]     = sum_v M[v,h] table[v,e]        (== da1^T x   for x = P table)
//   d_table[v,e] = sum_h M[v,h] w1[h,e]           (== P^T dx    for dx = da1 w1)
// Two small fp32 products.  A block owns an 8-row x 64-column output tile; K is walked in chunks of 128 staged in
// shared memory with wide coalesced loads (one memory round trip per chunk), each thread accumulating two adjacent
// columns in ascending-k order (bitwise reproducible).
// M may arrive as `mparts` split-K slices (mstride elements apart): they are summed in slice order while the A chunk
// is staged, so no separate reduction of M is needed.  The launch also carries the step's remaining split reductions
// (dW2 partials, bias-gradient column sums): blocks [0, red_units) run reduce_parts_block, the rest the two products --
// both depend only on the GEMMs before them, so one launch replaces reduce_parts_kernel + embed_finish_kernel.
constexpr int kFinRows = 8, kFinCols = 64, kFinK = 128;
__global__ void __launch_bounds__(256)
embed_finish_kernel(const ReduceJobs jobs, int red_units, int fin_x,
                    const float* __restrict__ M, int mparts, int64_t mstride, const float* __restrict__ table,
                    const float* __restrict__ w1, int V, int H, int E, float* __restrict__ dw1, float* __restrict__ d_table,
                    int accumulate) {
  __shared__ __align__(16) float As[kFinK][kFinRows];
  __shared__ __align__(16) float Bs[kFinK][kFinCols];
  pdl_trigger();
  pdl_wait();
  if ((int)blockIdx.x < red_units) { reduce_parts_block(jobs, (int)blockIdx.x); return; }
  const int fb = (int)blockIdx.x - red_units;
  const int bx = fb % fin_x, by = fb / fin_x;
  const int tiles1 = (H + kFinRows - 1) / kFinRows;       // dw1 row tiles come first, then d_table row tiles
  const bool second = bx >= tiles1;
  const int r0 = (second ? bx - tiles1 : bx) * kFinRows;
  const int c0 = by * kFinCols;
  const int rows = second ? V : H, K = second ? H : V;
  const float* B = second ? w1 : table;                   // [K, E]
  const int tr = threadIdx.x >> 5, tc2 = (threadIdx.x & 31) * 2;
  float acc0 = 0.f, acc1 = 0.f;
  for (int k0 = 0; k0 < K; k0 += kFinK) {
    const int kc = min(kFinK, K - k0);
    // B chunk first (its loads are in flight under the slice sums below): kFinK * kFinCols / 4 / 256 = 8 float4 per thread
    constexpr int NB = kFinK * (kFinCols / 4) / 256;
    float4 bv[NB];
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      const int i = threadIdx.x + e * 256;
      const int k = i / (kFinCols / 4), c = (i % (kFinCols / 4)) * 4;
      bv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < kc && c0 + c < E) bv[e] = __ldg(reinterpret_cast<const float4*>(B + (size_t)(k0 + k) * E + c0 + c));   // E % 4 == 0
    }
    // A chunk: As[k][r] = second ? M[(r0 + r) * H + k0 + k] : M[(k0 + k) * H + r0 + r]
    // kFinK * kFinRows / 256 = 4 elements per thread; the slice loads of all four are issued together (32 in flight per
    // thread and pass over eight slices): with one element at a time this staging ran at eight serial round trips per chunk
    {
      constexpr int NE = kFinK * kFinRows / 256;
      const float* mp[NE];
      float v[NE];
      int kk[NE], rr[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int i = threadIdx.x + e * 256;
        int k, r;
        if (second) { r = i / kFinK; k = i % kFinK; } else { k = i / kFinRows; r = i % kFinRows; }
        kk[e] = k; rr[e] = r; v[e] = 0.f;
        mp[e] = (k < kc && r0 + r < rows) ? (second ? M + (size_t)(r0 + r) * H + k0 + k : M + (size_t)(k0 + k) * H + r0 + r) : nullptr;
      }
      for (int pz = 0; pz < mparts; pz += 8) {             // slice order (bitwise reproducible)
        float t[NE][8];
#pragma unroll
        for (int e = 0; e < NE; ++e)
#pragma unroll
          for (int u = 0; u < 8; ++u) t[e][u] = (mp[e] && pz + u < mparts) ? __ldg(mp[e] + (size_t)(pz + u) * mstride) : 0.f;
#pragma unroll
        for (int e = 0; e < NE; ++e)
#pragma unroll
          for (int u = 0; u < 8; ++u) v[e] += t[e][u];
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) As[kk[e]][rr[e]] = v[e];
    }
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      const int i = threadIdx.x + e * 256;
      *reinterpret_cast<float4*>(&Bs[i / (kFinCols / 4)][(i % (kFinCols / 4)) * 4]) = bv[e];
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kFinK; ++k) {
      const float a = As[k][tr];
      const float2 b = *reinterpret_cast<const float2*>(&Bs[k][tc2]);
      acc0 = fmaf(a, b.x, acc0); acc1 = fmaf(a, b.y, acc1);
    }
    __syncthreads();
  }
  const int r = r0 + tr, c = c0 + tc2;
  if (r < rows && c < E) {                                // E % 4 == 0: c and c + 1 are both in range
    float* out = (second ? d_table : dw1) + (size_t)r * E + c;
    if (second && r == 0) { acc0 = 0.f; acc1 = 0.f; }      // padding row: never looked up (ids > 0), gradient 0 like the reference
    if (second && accumulate) { acc0 += out[0]; acc1 += out[1]; }
    out[0] = acc0; out[1] = acc1;
  }
}

constexpr int kNormRowsPerBlock = 32;
// dz = (dy - y (y.dy)) / |z|  -> fp32 + bf16; also per-block column sums of dz (-> db2), H <= 32 * KMAX <= 512.
// 16 warps x 2 rows: both rows of a warp are loaded together, so a block pays one memory round trip, not four.
// FROM_Y: the normalise step is described by (y in bf16, 1/|z|) instead of z:  dz = (dy - y (y.dy)) * inv_norm.
template <int KMAX, bool FROM_Y>
__global__ void __launch_bounds__(512)
l2norm_bwd_colsum_kernel(const float* __restrict__ dy, int dy_parts, int64_t dy_stride, const float* __restrict__ z,
                         const __nv_bfloat16* __restrict__ yb, const float* __restrict__ inv_norm,
                         int64_t R, int H, float* __restrict__ dz, __nv_bfloat16* __restrict__ dzb,
                         float* __restrict__ colsum_part) {
  __shared__ float s_part[16][32 * KMAX];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kNormRowsPerBlock + warp * 2;
  float g[2][KMAX], zv[2][KMAX];
  // dy may arrive as `dy_parts` split slices (the loss kernel's per-split partial gradients): summed in slice order
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool ok = row0 + i < R;
    const float* gr = dy + (row0 + i) * H;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int e = lane + 32 * k;
      if (FROM_Y) zv[i][k] = (ok && e < H) ? __bfloat162float(yb[(row0 + i) * H + e]) : 0.f;
      else        zv[i][k] = (ok && e < H) ? z[(row0 + i) * H + e] : 0.f;
      g[i][k] = (ok && e < H) ? gr[e] : 0.f;
    }
  }
  float inv[2] = {0.f, 0.f};
  if (FROM_Y) {
#pragma unroll
    for (int i = 0; i < 2; ++i) inv[i] = (row0 + i < R) ? __ldg(inv_norm + row0 + i) : 0.f;
  }
  for (int pp = 1; pp < dy_parts; ++pp) {
    float t[2][KMAX];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool ok = row0 + i < R;
      const float* gr = dy + (row0 + i) * H + (int64_t)pp * dy_stride;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) { const int e = lane + 32 * k; t[i][k] = (ok && e < H) ? gr[e] : 0.f; }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int k = 0; k < KMAX; ++k) g[i][k] += t[i][k];
  }
  float cs[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) cs[k] = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float ss = 0.f, dot = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { const float v = zv[i][k]; ss = fmaf(v, v, ss); dot = fmaf(v, g[i][k], dot); }
    ss = warp_sum(ss); dot = warp_sum(dot);
    const float n = sqrtf(ss), denom = FROM_Y ? 1.0f : fmaxf(n, 1e-12f);
    const float inner = FROM_Y ? dot : ((n > 1e-12f) ? dot / (denom * denom) : 0.f);
    const float oscale = FROM_Y ? inv[i] : 1.0f / denom;
    if (row0 + i < R) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int e = lane + 32 * k;
        if (e < H) {
          const float o = (g[i][k] - zv[i][k] * inner) * oscale;
          if (dz) dz[(row0 + i) * H + e] = o;
          dzb[(row0 + i) * H + e] = __float2bfloat16(o);
          cs[k] += o;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) s_part[warp][lane + 32 * k] = cs[k];
  __syncthreads();
  for (int e = threadIdx.x; e < H; e += 512) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) t += s_part[w][e];
    colsum_part[(size_t)blockIdx.x * H + e] = t;
  }
}

// dz = (dy - y (y.dy)) / |z|  -> fp32 + bf16
__global__ void __launch_bounds__(256)
l2norm_bwd_bf16_kernel(const float* __restrict__ dy, const float* __restrict__ z, int64_t R, int H,
                       float* __restrict__ dz, __nv_bfloat16* __restrict__ dzb) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* zr = z + row * H; const float* gr = dy + row * H;
  float ss = 0.f, dot = 0.f;
  for (int e = lane; e < H; e += 32) { float v = zr[e]; ss = fmaf(v, v, ss); dot = fmaf(v, gr[e], dot); }
  ss = warp_sum(ss); dot = warp_sum(dot);
  const float n = sqrtf(ss), denom = fmaxf(n, 1e-12f);
  const float inner = (n > 1e-12f) ? dot / (denom * denom) : 0.f;
  for (int e = lane; e < H; e += 32) {
    const float o = (gr[e] - zr[e] * inner) / denom;
    dz[row * H + e] = o;
    dzb[row * H + e] = __float2bfloat16(o);
  }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------
// tower MLP on the tensor cores
// ---------------------------------------------------------------------------------------
struct TcMlpPlan {
  int s_dw2, s_dw1;
  size_t xb, w1b, w2b, h1b, act_f, act_b, partial, partial2, cs1, cs2, colsum, total;
};
static inline int pad8(int n) { return (n + 7) / 8 * 8; }
static TcMlpPlan plan_tc_mlp(int64_t R, int E, int H) {
  TcMlpPlan p{};
  p.s_dw2 = tc::pick_splits(H, H, (int)R);
  p.s_dw1 = tc::pick_splits(H, E, (int)R);
  p.xb = align_up((size_t)R * pad8(E) * 2);               // E % 8 != 0: bf16 operands are kept with a row pitch of pad8(E)
  p.w1b = align_up((size_t)H * pad8(E) * 2);
  p.w2b = align_up((size_t)H * H * 2);
  p.h1b = align_up((size_t)R * H * 2);
  p.act_f = align_up((size_t)R * H * 4);
  p.act_b = align_up((size_t)R * H * 2);
  p.partial = align_up(p.s_dw2 > 1 ? (size_t)p.s_dw2 * H * H * 4 : 0);
  p.partial2 = align_up(p.s_dw1 > 1 ? (size_t)p.s_dw1 * H * E * 4 : 0);
  p.cs1 = align_up((size_t)ceil_div(R, 32) * H * 4);                        // da1 column-sum partials (GEMM epilogue)
  p.cs2 = align_up((size_t)ceil_div(R, tc::kNormRowsPerBlock) * H * 4);     // dz column-sum partials (normalise bwd)
  p.colsum = align_up((size_t)colsum_partial_rows(R) * H * 4);
  // bwd: xb, w1b, w2b, h1b, dz (f32+bf16), da1 (f32+bf16), split partials x2, column-sum partials x2
  p.total = p.xb + p.w1b + p.w2b + p.h1b + 2 * p.act_f + 2 * p.act_b + p.partial + p.partial2 + p.cs1 + p.cs2 +
            p.colsum + 1024;
  return p;
}

size_t tc_mlp_workspace(int64_t R, int E, int H) { return plan_tc_mlp(R, E, H).total; }

// E may be anything (e.g. the word tower's 300): x / W1 are then converted into pad8(E)-pitch bf16 rows inside the call
static bool tc_mlp_supported(int E, int H) { return E > 0 && (H % 8 == 0); }

int tc_mlp_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t R, int E,
               int H, float* h1, float* z, float* y, __nv_bfloat16* y_bf16, const __nv_bfloat16* x_bf16,
               const __nv_bfloat16* w1_bf16, const __nv_bfloat16* w2_bf16, __nv_bfloat16* h1_bf16, float* inv_norm,
               const tt_mlp_embed_t* embed, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!tc_mlp_supported(E, H)) { set_error("TT_PREC_BF16 mlp needs H %% 8 == 0 (E=%d H=%d)", E, H); return TT_ERR_UNSUPPORTED; }
  const TcMlpPlan plan = plan_tc_mlp(R, E, H);
  if (ws == nullptr || ws_bytes < plan.total) { set_error("tc_mlp_fwd: workspace too small (%zu < %zu)", ws_bytes, plan.total); return TT_ERR_WORKSPACE; }
  Workspace w(ws, ws_bytes);
  const int Ep = pad8(E);
  __nv_bfloat16* xb = w.take<__nv_bfloat16>((size_t)R * Ep);
  __nv_bfloat16* w1b = w.take<__nv_bfloat16>((size_t)H * Ep);
  __nv_bfloat16* w2b = w.take<__nv_bfloat16>((size_t)H * H);
  __nv_bfloat16* h1b = w.take<__nv_bfloat16>((size_t)R * H);
  int rc = TT_OK;
  if (embed) {
    if (!w1_bf16 || !w2_bf16 || !tc_mlp_fwd_pool_supported(E, H, embed->V)) { set_error("tc_mlp_fwd: embed needs bf16 weight shadows and a supported shape (tt_mlp_fwd_embed_ok)"); return TT_ERR_UNSUPPORTED; }
    return tc_mlp_fwd_fused(nullptr, w1_bf16, b1, w2_bf16, b2, R, E, H, h1_bf16 ? h1_bf16 : reinterpret_cast<__nv_bfloat16*>(h1), z, y, y_bf16,
                            inv_norm, (const __nv_bfloat16*)embed->pool_bf16, embed->V, (const __nv_bfloat16*)embed->table_bf16, s,
                            embed->ids, embed->id_bytes, embed->L, embed->inv_len);
  }
  if (Ep != E) {                               // padded pitch: contiguous caller shadows of x / W1 cannot feed TMA, convert here
    if (embed) { set_error("tc_mlp_fwd: embed needs E %% 8 == 0"); return TT_ERR_UNSUPPORTED; }
    rc = tc::cast_rows(x, R, E, Ep, xb, s); if (rc) return rc;
    rc = tc::cast_rows(w1, H, E, Ep, w1b, s); if (rc) return rc;
    x_bf16 = xb; w1_bf16 = w1b;
  }
  if (!x_bf16 || !w1_bf16 || !w2_bf16) {       // only the operands without a caller-provided shadow are converted
    rc = tc::cast3(x_bf16 ? nullptr : x, xb, x_bf16 ? 0 : R * E, w1_bf16 ? nullptr : w1, w1b, w1_bf16 ? 0 : (int64_t)H * E,
                   w2_bf16 ? nullptr : w2, w2b, w2_bf16 ? 0 : (int64_t)H * H, s);
    if (rc) return rc;
  }
  const __nv_bfloat16* xa = x_bf16 ? x_bf16 : xb;
  const __nv_bfloat16* w1a = w1_bf16 ? w1_bf16 : w1b;
  const __nv_bfloat16* w2a = w2_bf16 ? w2_bf16 : w2b;
  // saved hidden activation: bf16.  Without a caller-owned shadow it is stored in the (4-byte per element,
  // so large enough) h1 buffer itself -- TT_PREC_BF16 treats h1 as opaque saved state.
  h1b = h1_bf16 ? h1_bf16 : reinterpret_cast<__nv_bfloat16*>(h1);
  if (tc_mlp_fused_supported(E, H))
    return tc_mlp_fwd_fused(xa, w1a, b1, w2a, b2, R, E, H, h1b, z, y, y_bf16, inv_norm, nullptr, 0, nullptr, s);
  if (z == nullptr) z = w.take<float>((size_t)R * H);     // unfused shapes: the pre-normalise tensor lives in the workspace
  tc::TcGemm g{};
  g.M = (int)R; g.N = H; g.K = E; g.A = xa; g.a_mn = 0; g.B = w1a; g.b_mn = 0; g.lda = g.ldb = Ep;
  g.C = nullptr; g.Cb = h1b; g.ldc = H; g.bias = b1; g.act = 1;
  rc = tc::tc_gemm(g, s); if (rc) return rc;
  g = tc::TcGemm{};
  g.M = (int)R; g.N = H; g.K = H; g.A = h1b; g.a_mn = 0; g.B = w2a; g.b_mn = 0;
  g.C = z; g.ldc = H; g.bias = b2;
  rc = tc::tc_gemm(g, s); if (rc) return rc;
  tc::l2norm_fwd_bf16_kernel<<<(unsigned)ceil_div(R, 8), 256, 0, s>>>(z, R, H, y, y_bf16, inv_norm);
  TT_LAUNCH_CHECK("l2norm_fwd_bf16_kernel");
  return TT_OK;
}

int tc_mlp_bwd(const float* dy, const float* x, const float* w1, const float* w2, const float* h1, const float* z,
               int64_t R, int E, int H, float* dx, float* dw1, float* db1, float* dw2, float* db2,
               const __nv_bfloat16* x_bf16, const __nv_bfloat16* w1_bf16, const __nv_bfloat16* w2_bf16,
               const __nv_bfloat16* h1_bf16, int dy_parts, int64_t dy_part_stride, const tt_mlp_embed_t* embed,
               const __nv_bfloat16* y_bf16, const float* inv_norm, const __nv_bfloat16* dz_bf16, const float* dz_colsum,
               void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!tc_mlp_supported(E, H)) { set_error("TT_PREC_BF16 mlp needs H %% 8 == 0"); return TT_ERR_UNSUPPORTED; }
  const int Ep = pad8(E);
  if (embed && (embed->V % 8 != 0 || E % 4 != 0)) { set_error("tc_mlp_bwd: embed needs V %% 8 == 0 and E %% 4 == 0"); return TT_ERR_UNSUPPORTED; }
  if (dy_parts < 1) dy_parts = 1;
  if (dy_parts > 1 && H > 512) { set_error("tc_mlp_bwd: split dy needs H <= 512"); return TT_ERR_UNSUPPORTED; }
  const TcMlpPlan plan = plan_tc_mlp(R, E, H);
  if (ws == nullptr || ws_bytes < plan.total) { set_error("tc_mlp_bwd: workspace too small (%zu < %zu)", ws_bytes, plan.total); return TT_ERR_WORKSPACE; }
  Workspace w(ws, ws_bytes);
  __nv_bfloat16* xb = w.take<__nv_bfloat16>((size_t)R * Ep);
  __nv_bfloat16* w1b = w.take<__nv_bfloat16>((size_t)H * Ep);
  __nv_bfloat16* w2b = w.take<__nv_bfloat16>((size_t)H * H);
  __nv_bfloat16* h1b = w.take<__nv_bfloat16>((size_t)R * H);
  float* dz = w.take<float>((size_t)R * H);
  float* da1 = w.take<float>((size_t)R * H);
  __nv_bfloat16* dzb = w.take<__nv_bfloat16>((size_t)R * H);
  __nv_bfloat16* da1b = w.take<__nv_bfloat16>((size_t)R * H);
  float* partial = plan.partial ? w.take<float>(plan.partial / 4) : nullptr;
  float* partial2 = plan.partial2 ? w.take<float>(plan.partial2 / 4) : nullptr;
  float* cs1 = w.take<float>(plan.cs1 / 4);
  float* cs2 = w.take<float>(plan.cs2 / 4);
  float* cpart = w.take<float>(plan.colsum / 4);
  int rc = TT_OK;
  if (Ep != E) {                               // see tc_mlp_fwd
    if (embed) { set_error("tc_mlp_bwd: embed needs E %% 8 == 0"); return TT_ERR_UNSUPPORTED; }
    rc = tc::cast_rows(x, R, E, Ep, xb, s); if (rc) return rc;
    rc = tc::cast_rows(w1, H, E, Ep, w1b, s); if (rc) return rc;
    x_bf16 = xb; w1_bf16 = w1b;
  }
  if (!x_bf16 || !w1_bf16 || !w2_bf16) {
    rc = tc::cast3(x_bf16 ? nullptr : x, xb, x_bf16 ? 0 : R * E, w1_bf16 ? nullptr : w1, w1b, w1_bf16 ? 0 : (int64_t)H * E,
                   w2_bf16 ? nullptr : w2, w2b, w2_bf16 ? 0 : (int64_t)H * H, s);
    if (rc) return rc;
  }
  const __nv_bfloat16* xa = x_bf16 ? x_bf16 : xb;
  const __nv_bfloat16* w1a = w1_bf16 ? w1_bf16 : w1b;
  const __nv_bfloat16* w2a = w2_bf16 ? w2_bf16 : w2b;
  const __nv_bfloat16* h1a = h1_bf16 ? h1_bf16 : reinterpret_cast<const __nv_bfloat16*>(h1);   // see tc_mlp_fwd
  const bool fused_cs = H <= 512;
  const int nblk2 = (int)ceil_div(R, tc::kNormRowsPerBlock);
  if (dz_bf16) {                                            // normalise backward already done by the loss kernel
    dzb = const_cast<__nv_bfloat16*>(dz_bf16);
    cs2 = const_cast<float*>(dz_colsum);
  } else if (fused_cs) {
    // fp32 dz is not needed: db2 comes from cs2
    const bool from_y = y_bf16 != nullptr && inv_norm != nullptr;
    auto kern = H <= 256 ? (from_y ? tc::l2norm_bwd_colsum_kernel<8, true> : tc::l2norm_bwd_colsum_kernel<8, false>)
                         : (from_y ? tc::l2norm_bwd_colsum_kernel<16, true> : tc::l2norm_bwd_colsum_kernel<16, false>);
    TT_CUDA(launch_kernel(kern, dim3((unsigned)nblk2), dim3(512), 0, s, true, dy, dy_parts, dy_part_stride, z, y_bf16, inv_norm, R, H,
                          (float*)nullptr, dzb, cs2));
    TT_LAUNCH_CHECK("l2norm_bwd_colsum_kernel");
  } else {
    if (z == nullptr) { set_error("tc_mlp_bwd: H > 512 needs z"); return TT_ERR_UNSUPPORTED; }
    tc::l2norm_bwd_bf16_kernel<<<(unsigned)ceil_div(R, 8), 256, 0, s>>>(dy, z, R, H, dz, dzb);
    TT_LAUNCH_CHECK("l2norm_bwd_bf16_kernel");
    rc = colsum(dz, R, H, H, db2, cpart, s); if (rc) return rc;
  }
  // launch 1: { dw2[H,H] = dz^T h1 (split-K partials),  da1[R,H] = (dz w2) * (h1 > 0) } -- independent, one grid
  tc::TcGemm gw2{}, ga1{};
  gw2.M = H; gw2.N = H; gw2.K = (int)R; gw2.A = dzb; gw2.a_mn = 1; gw2.B = h1a; gw2.b_mn = 1; gw2.C = dw2; gw2.ldc = H;
  gw2.splits = plan.s_dw2; gw2.partial = partial; gw2.defer_reduce = true;
  ga1.M = (int)R; ga1.N = H; ga1.K = H; ga1.A = dzb; ga1.a_mn = 0; ga1.B = w2a; ga1.b_mn = 1; ga1.C = nullptr; ga1.Cb = da1b;
  ga1.ldc = H; ga1.mask = h1a; ga1.ldmask = H; ga1.colsum_part = cs1;   // fp32 da1 never needed: db1 comes from cs1
  rc = tc::tc_gemm2(ga1, gw2, s); if (rc) return rc;
  if (embed) {
    // x = P table: M[V,H] = P^T da1 (split-K partials) replaces dx, dw1 = da1^T x and the embedding backward
    const int V = (int)embed->V;
    const int sm = tc::pick_splits(V, H, (int)R);
    const size_t need = tc_mlp_embed_workspace(V, H, R);
    if (embed->workspace == nullptr || embed->workspace_bytes < need) { set_error("tc_mlp_bwd: embed workspace too small (%zu < %zu)", embed->workspace_bytes, need); return TT_ERR_WORKSPACE; }
    Workspace we(embed->workspace, embed->workspace_bytes);
    float* Mbuf = we.take<float>((size_t)V * H);
    float* Mpart = sm > 1 ? we.take<float>((size_t)sm * V * H) : nullptr;
    tc::TcGemm gm{};
    gm.M = V; gm.N = H; gm.K = (int)R; gm.A = (const __nv_bfloat16*)embed->pool_bf16; gm.a_mn = 1; gm.B = da1b; gm.b_mn = 1;
    gm.C = Mbuf; gm.ldc = H; gm.splits = sm; gm.partial = Mpart; gm.defer_reduce = true;
    rc = tc::tc_gemm(gm, s); if (rc) return rc;
    // ONE launch finishes the call: the remaining split reductions (dW2, db1, db2; fixed order) and the two small
    // products of the embedding-fused backward, which read M's split-K slices directly
    ReduceJobs jobs{};
    int nj = 0;
    if (plan.s_dw2 > 1) jobs.job[nj++] = make_job(partial, plan.s_dw2, (int64_t)H * H, (int64_t)H * H, dw2);
    jobs.job[nj++] = make_job(cs1, (int)ceil_div(R, 32), H, H, db1);
    if (fused_cs) jobs.job[nj++] = make_job(cs2, nblk2, H, H, db2);
    jobs.njobs = nj;
    const int red_units = plan_reduce(jobs);
    const int fin_x = (int)(ceil_div(H, tc::kFinRows) + ceil_div(V, tc::kFinRows)), fin_y = (int)ceil_div(E, tc::kFinCols);
    TT_CUDA(launch_kernel(tc::embed_finish_kernel, dim3((unsigned)(red_units + fin_x * fin_y)), dim3(256), 0, s, true, jobs, red_units, fin_x,
                          (const float*)(sm > 1 ? Mpart : Mbuf), sm > 1 ? sm : 1, (int64_t)V * H, embed->table, w1,
                          V, H, E, dw1, embed->d_table, embed->accumulate));
    TT_LAUNCH_CHECK("embed_finish_kernel");
    return TT_OK;
  }
  // launch 2: { dw1[H,E] = da1^T x (split-K partials),  dx[R,E] = da1 w1 }
  tc::TcGemm gw1{}, gdx{};
  gw1.M = H; gw1.N = E; gw1.K = (int)R; gw1.A = da1b; gw1.a_mn = 1; gw1.B = xa; gw1.b_mn = 1; gw1.ldb = Ep; gw1.C = dw1; gw1.ldc = E;
  gw1.splits = plan.s_dw1; gw1.partial = partial2; gw1.defer_reduce = true;
  if (dx) {
    gdx.M = (int)R; gdx.N = E; gdx.K = H; gdx.A = da1b; gdx.a_mn = 0; gdx.B = w1a; gdx.b_mn = 1; gdx.ldb = Ep; gdx.C = dx; gdx.ldc = E;
    rc = tc::tc_gemm2(gdx, gw1, s); if (rc) return rc;
  } else {
    rc = tc::tc_gemm(gw1, s); if (rc) return rc;
  }
  // one launch finishes every reduction of this call in a fixed order: dW2, dW1 split-K partials, db1, db2
  ReduceJobs jobs{};
  int nj = 0;
  if (plan.s_dw2 > 1) jobs.job[nj++] = make_job(partial, plan.s_dw2, (int64_t)H * H, (int64_t)H * H, dw2);
  if (plan.s_dw1 > 1) jobs.job[nj++] = make_job(partial2, plan.s_dw1, (int64_t)H * E, (int64_t)H * E, dw1);
  jobs.job[nj++] = make_job(cs1, (int)ceil_div(R, 32), H, H, db1);
  if (fused_cs) jobs.job[nj++] = make_job(cs2, nblk2, H, H, db2);
  jobs.njobs = nj;
  return reduce_parts(jobs, s);
}

size_t tc_mlp_embed_workspace(int64_t V, int H, int64_t R) {
  const int sm = tc::pick_splits((int)V, H, (int)R);
  return align_up((size_t)V * H * 4) + align_up(sm > 1 ? (size_t)sm * V * H * 4 : 0) + 256;
}

// ---------------------------------------------------------------------------------------
// avg_pool tower projection on the tensor cores (TT_PREC_BF16 path of K3', encoders.py:100-104,144-147):
// the Linear(E,H) forward and its two backward products; Dropout / LayerNorm / normalise stay in tower.cu's fp32
// row kernels (they are HBM-bound elementwise work).  Any E and H: operands are converted to pad8-pitch bf16 rows.
// ---------------------------------------------------------------------------------------
struct TcProjPlan { int s_dw; size_t xb, wb, dab, partial, total; };
static TcProjPlan plan_tc_proj(int64_t R, int E, int H) {
  TcProjPlan p{};
  p.s_dw = tc::pick_splits(H, E, (int)R);
  p.xb = align_up((size_t)R * pad8(E) * 2);
  p.wb = align_up((size_t)H * pad8(E) * 2);
  p.dab = align_up((size_t)R * pad8(H) * 2);
  p.partial = align_up(p.s_dw > 1 ? (size_t)p.s_dw * H * E * 4 : 0);
  p.total = p.xb + p.wb + p.dab + p.partial + 1024;
  return p;
}
size_t tc_proj_workspace(int64_t R, int E, int H) { return plan_tc_proj(R, E, H).total; }

// a[R,H] = x[R,E] w[H,E]^T + b
int tc_proj_fwd(const float* x, const float* w, const float* b, int64_t R, int E, int H, float* a, void* ws, size_t ws_bytes,
                cudaStream_t s) {
  const TcProjPlan plan = plan_tc_proj(R, E, H);
  if (ws == nullptr || ws_bytes < plan.total) { set_error("tc_proj_fwd: workspace too small (%zu < %zu)", ws_bytes, plan.total); return TT_ERR_WORKSPACE; }
  Workspace wk(ws, ws_bytes);
  const int Ep = pad8(E);
  __nv_bfloat16* xb = wk.take<__nv_bfloat16>((size_t)R * Ep);
  __nv_bfloat16* wb = wk.take<__nv_bfloat16>((size_t)H * Ep);
  int rc = tc::cast_rows(x, R, E, Ep, xb, s); if (rc) return rc;
  rc = tc::cast_rows(w, H, E, Ep, wb, s); if (rc) return rc;
  tc::TcGemm g{};
  g.M = (int)R; g.N = H; g.K = E; g.A = xb; g.a_mn = 0; g.B = wb; g.b_mn = 0; g.lda = g.ldb = Ep;
  g.C = a; g.ldc = H; g.bias = b;
  return tc::tc_gemm(g, s);
}

// dw[H,E] = da^T x (fixed-order split-K), dx[R,E] = da w (nullable)
int tc_proj_bwd(const float* da, const float* x, const float* w, int64_t R, int E, int H, float* dx, float* dw, void* ws,
                size_t ws_bytes, cudaStream_t s) {
  const TcProjPlan plan = plan_tc_proj(R, E, H);
  if (ws == nullptr || ws_bytes < plan.total) { set_error("tc_proj_bwd: workspace too small (%zu < %zu)", ws_bytes, plan.total); return TT_ERR_WORKSPACE; }
  Workspace wk(ws, ws_bytes);
  const int Ep = pad8(E), Hp = pad8(H);
  __nv_bfloat16* xb = wk.take<__nv_bfloat16>((size_t)R * Ep);
  __nv_bfloat16* wb = wk.take<__nv_bfloat16>((size_t)H * Ep);
  __nv_bfloat16* dab = wk.take<__nv_bfloat16>((size_t)R * Hp);
  float* partial = plan.partial ? wk.take<float>(plan.partial / 4) : nullptr;
  int rc = tc::cast_rows(x, R, E, Ep, xb, s); if (rc) return rc;
  rc = tc::cast_rows(da, R, H, Hp, dab, s); if (rc) return rc;
  tc::TcGemm gw{};
  gw.M = H; gw.N = E; gw.K = (int)R; gw.A = dab; gw.a_mn = 1; gw.lda = Hp; gw.B = xb; gw.b_mn = 1; gw.ldb = Ep;
  gw.C = dw; gw.ldc = E; gw.splits = plan.s_dw; gw.partial = partial;
  rc = tc::tc_gemm(gw, s); if (rc) return rc;
  if (dx) {
    rc = tc::cast_rows(w, H, E, Ep, wb, s); if (rc) return rc;
    tc::TcGemm gx{};
    gx.M = (int)R; gx.N = E; gx.K = H; gx.A = dab; gx.a_mn = 0; gx.lda = Hp; gx.B = wb; gx.b_mn = 1; gx.ldb = Ep;
    gx.C = dx; gx.ldc = E;
    rc = tc::tc_gemm(gx, s); if (rc) return rc;
  }
  return TT_OK;
}

// self-test hook used by the GPU test-suite: C = A * B on the tensor cores with either operand major
int tc_gemm_selftest(const __nv_bfloat16* A, int a_mn, const __nv_bfloat16* B, int b_mn, int M, int N, int K,
                     float* C, int splits, float* partial, cudaStream_t s) {
  tc::TcGemm g{};
  g.M = M; g.N = N; g.K = K; g.A = A; g.a_mn = a_mn; g.B = B; g.b_mn = b_mn; g.C = C; g.ldc = N;
  g.splits = splits; g.partial = partial;
  return tc::tc_gemm(g, s);
}

}  // namespace tt

extern "C" int tt_selftest_tc_gemm(const void* a_bf16, int a_mn_major, const void* b_bf16, int b_mn_major, int M, int N,
                                   int K, float* c, int splits, float* partial, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(a_bf16 && b_bf16 && c && M > 0 && N > 0 && K > 0 && splits >= 1, "selftest_tc_gemm: bad arguments");
  return tt::tc_gemm_selftest((const __nv_bfloat16*)a_bf16, a_mn_major, (const __nv_bfloat16*)b_bf16, b_mn_major, M, N, K,
                              c, splits, partial, static_cast<cudaStream_t>(stream));
}
