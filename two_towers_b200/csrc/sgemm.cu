// sgemm.cu -- fp32 FFMA GEMM (register-blocked, smem double-buffered) + fixed-order split-K.
#include "sgemm.cuh"
#include "reduce.cuh"

namespace tt {

template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(SgemmArgs a, int kchunk) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TXN = BN / TN;                  // threads along n
  constexpr int MB = TM / 4, NB = TN / 4;       // 4-wide sub-blocks per thread
  constexpr int MSTRIDE = BM / MB, NSTRIDE = BN / NB;
  constexpr int A_ELEMS = BM * BK / NT, B_ELEMS = BN * BK / NT;
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/thread mismatch");

  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(a.K, kbeg + kchunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[A_ELEMS], rb[B_ELEMS];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_ELEMS; ++i) {
      int e = tid + i * NT;
      int m, k;
      if (TA) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < a.M && gk < kend)
        v = TA ? __ldg(a.A + (size_t)gk * a.lda + gm) : __ldg(a.A + (size_t)gm * a.lda + gk);
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_ELEMS; ++i) {
      int e = tid + i * NT;
      int n, k;
      if (TB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < a.N && gk < kend)
        v = TB ? __ldg(a.B + (size_t)gn * a.ldb + gk) : __ldg(a.B + (size_t)gk * a.ldb + gn);
      rb[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_ELEMS; ++i) {
      int e = tid + i * NT;
      int m, k;
      if (TA) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
      As[buf][k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_ELEMS; ++i) {
      int e = tid + i * NT;
      int n, k;
      if (TB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      Bs[buf][k][n] = rb[i];
    }
  };

  if (kbeg < kend) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK, buf ^= 1) {
    const bool more = (k0 + BK) < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int ib = 0; ib < MB; ++ib) {
        float4 t = *reinterpret_cast<const float4*>(&As[buf][k][ib * MSTRIDE + ty * 4]);
        av[ib * 4 + 0] = t.x; av[ib * 4 + 1] = t.y; av[ib * 4 + 2] = t.z; av[ib * 4 + 3] = t.w;
      }
#pragma unroll
      for (int jb = 0; jb < NB; ++jb) {
        float4 t = *reinterpret_cast<const float4*>(&Bs[buf][k][jb * NSTRIDE + tx * 4]);
        bv[jb * 4 + 0] = t.x; bv[jb * 4 + 1] = t.y; bv[jb * 4 + 2] = t.z; bv[jb * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
  }

  const bool split = a.splits > 1;
  float* out = split ? a.partial + (size_t)blockIdx.z * a.M * a.N : a.C;
  const int ldo = split ? a.N : a.ldc;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + (i / 4) * MSTRIDE + ty * 4 + (i % 4);
    if (gm >= a.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int gn = n0 + (j / 4) * NSTRIDE + tx * 4 + (j % 4);
      if (gn >= a.N) continue;
      float v = acc[i][j];
      if (!split) {
        if (a.bias) v += a.bias[gn];
        if (a.act == 1) v = fmaxf(v, 0.f);
        if (a.mask) v = (a.mask[(size_t)gm * a.ldmask + gn] > 0.f) ? v : 0.f;
      }
      out[(size_t)gm * ldo + gn] = v;
    }
  }
}

int sgemm_pick_splits(int M, int N, int K) {
  const bool big = (M >= 128 && N >= 128);
  const int bm = big ? 128 : 64, bn = big ? 128 : 64;
  int64_t tiles = ceil_div(M, bm) * ceil_div(N, bn);
  if (tiles >= kNumSMs || K <= 256) return 1;
  int64_t want = ceil_div(2 * kNumSMs, tiles);
  int64_t maxs = K / 128;               // keep >= 128 of K per split
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (want > 64) want = 64;
  return (int)(want < 1 ? 1 : want);
}

template <int BM, int BN, int BK, int TM, int TN>
static void launch_cfg(const SgemmArgs& a, int kchunk, dim3 grid, cudaStream_t s) {
  constexpr int NT = (BM / TM) * (BN / TN);
  if (a.transA) {
    if (a.transB) (void)launch_kernel(sgemm_kernel<BM, BN, BK, TM, TN, true, true>, grid, dim3(NT), 0, s, true, a, kchunk);
    else          (void)launch_kernel(sgemm_kernel<BM, BN, BK, TM, TN, true, false>, grid, dim3(NT), 0, s, true, a, kchunk);
  } else {
    if (a.transB) (void)launch_kernel(sgemm_kernel<BM, BN, BK, TM, TN, false, true>, grid, dim3(NT), 0, s, true, a, kchunk);
    else          (void)launch_kernel(sgemm_kernel<BM, BN, BK, TM, TN, false, false>, grid, dim3(NT), 0, s, true, a, kchunk);
  }
}

int sgemm(const SgemmArgs& a, cudaStream_t stream) {
  TT_CHECK_ARG(a.M > 0 && a.N > 0 && a.K >= 0, "sgemm: bad shape %d %d %d", a.M, a.N, a.K);
  TT_CHECK_ARG(a.splits >= 1 && (a.splits == 1 || a.partial != nullptr),
               "sgemm: split-K needs a partial buffer");
  int kchunk = (int)ceil_div(a.K > 0 ? a.K : 1, a.splits);
  const bool big = (a.M >= 128 && a.N >= 128) &&
                   (ceil_div(a.M, 128) * ceil_div(a.N, 128) * a.splits >= kNumSMs / 2);
  if (big) {
    kchunk = (int)ceil_div(kchunk, 8) * 8;
    dim3 grid((unsigned)ceil_div(a.N, 128), (unsigned)ceil_div(a.M, 128), (unsigned)a.splits);
    launch_cfg<128, 128, 8, 8, 8>(a, kchunk, grid, stream);
  } else {
    kchunk = (int)ceil_div(kchunk, 16) * 16;
    dim3 grid((unsigned)ceil_div(a.N, 64), (unsigned)ceil_div(a.M, 64), (unsigned)a.splits);
    launch_cfg<64, 64, 16, 4, 4>(a, kchunk, grid, stream);
  }
  TT_LAUNCH_CHECK("sgemm_kernel");
  if (a.splits > 1) return splitk_reduce(a, stream);
  return TT_OK;
}

int splitk_reduce(const SgemmArgs& a, cudaStream_t stream) {
  ReduceJobs jobs{};
  ReduceJob& j = jobs.job[0];
  j = make_job(a.partial, a.splits, (int64_t)a.M * a.N, (int64_t)a.M * a.N, a.C);
  j.bias = a.bias; j.act = a.act; j.mask = a.mask; j.ldmask = a.ldmask; j.ncols = a.N; j.ldo = a.ldc;
  jobs.njobs = 1;
  return reduce_parts(jobs, stream);
}

// ---------------------------------------------------------------------------------------
// column sums (bias gradients)
// ---------------------------------------------------------------------------------------
__global__ void colsum_stage1(const float* __restrict__ X, int64_t M, int N, int ldx, int64_t rows_per,
                              float* __restrict__ partial) {
  __shared__ float sm[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per;
  const int64_t r1 = (r0 + rows_per < M) ? r0 + rows_per : M;
  float s = 0.f;
  if (col < N)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += X[r * ldx + col];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += sm[y][threadIdx.x];
    partial[(size_t)blockIdx.y * N + col] = t;
  }
}
__global__ void colsum_stage2(const float* __restrict__ partial, int P, int N, float* __restrict__ out) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float t = 0.f;
  for (int p = 0; p < P; ++p) t += partial[(size_t)p * N + col];
  out[col] = t;
}

int colsum_partial_rows(int64_t M) {
  int64_t p = ceil_div(M, 128);
  if (p > 64) p = 64;
  if (p < 1) p = 1;
  return (int)p;
}

int colsum(const float* X, int64_t M, int N, int ldx, float* out, float* partial, cudaStream_t stream) {
  const int P = colsum_partial_rows(M);
  const int64_t rows_per = ceil_div(M, P);
  dim3 grid((unsigned)ceil_div(N, 32), (unsigned)P);
  colsum_stage1<<<grid, dim3(32, 8), 0, stream>>>(X, M, N, ldx, rows_per, partial);
  TT_LAUNCH_CHECK("colsum_stage1");
  colsum_stage2<<<(unsigned)ceil_div(N, 128), 128, 0, stream>>>(partial, P, N, out);
  TT_LAUNCH_CHECK("colsum_stage2");
  return TT_OK;
}

}  // namespace tt
