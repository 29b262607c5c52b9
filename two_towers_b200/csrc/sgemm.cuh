// sgemm.cuh -- fp32 CUDA-core GEMM used by the TT_PREC_FP32 (parity) mode.
//
//   C[M,N] = epilogue( op(A)[M,K] * op(B)[K,N] )        all row-major with leading dims
//   op(A): transA ? A[k*lda+m] : A[m*lda+k]       op(B): transB ? B[n*ldb+k] : B[k*ldb+n]
//   epilogue: + bias[n]  -> relu (act==1) -> * (mask[m,n] > 0)
// Split-K (splits > 1) writes per-split partials and reduces them in a FIXED order, so the
// result is bitwise reproducible.  Pure FFMA: every product and accumulation is IEEE fp32,
// which is what keeps this mode inside rel 1e-5 of the reference's fp32 ATen path.
#pragma once
#include "common.cuh"

namespace tt {

struct SgemmArgs {
  int M, N, K;
  const float* A; int lda; int transA;
  const float* B; int ldb; int transB;
  float* C; int ldc;
  const float* bias = nullptr;     // [N]
  int act = 0;                     // 0 none, 1 relu
  const float* mask = nullptr;     // [M,ldmask]: C *= (mask > 0)
  int ldmask = 0;
  int splits = 1;                  // split-K factor
  float* partial = nullptr;        // [splits, M, N] when splits > 1
};

// bytes of `partial` needed for a given problem/splits
inline size_t sgemm_partial_bytes(int M, int N, int splits) {
  return splits > 1 ? (size_t)splits * M * N * sizeof(float) : 0;
}
// heuristic split-K factor so that the grid covers the 148 SMs
int sgemm_pick_splits(int M, int N, int K);
int sgemm(const SgemmArgs& a, cudaStream_t stream);
// fixed-order sum of a.partial[0..splits) + epilogue -> a.C (used by the fp32 and the tcgen05 GEMMs)
int splitk_reduce(const SgemmArgs& a, cudaStream_t stream);

// column sums: out[n] = sum_m X[m,n]  (bias gradients), fixed-order two-stage reduction.
// partial must hold colsum_partial_rows(M) * N floats.
int colsum_partial_rows(int64_t M);
int colsum(const float* X, int64_t M, int N, int ldx, float* out, float* partial, cudaStream_t stream);

}  // namespace tt
