// inbatch_ce.cu -- K4: in-batch sampled-softmax loss, fused similarity GEMM + online
// logsumexp cross-entropy, forward and backward (TT_PREC_FP32 path + the finalisation
// kernels shared with the tcgen05 path).
//
// Reference: in_batch_sampled_softmax_loss, twotower/losses.py:107 (S = Q D^T), :110
// (S / temperature), :113 (labels = arange), :116 (F.cross_entropy, mean).  The reference
// writes S, logits, log-softmax and their gradient (4 x [B,B] fp32); here a 64x64 tile of S
// lives in registers, each row keeps a running (max, sum-exp), and backward recomputes the
// tile from Q, D and the saved row logsumexp -- nothing of size B x B touches HBM.
//
// Backward runs the same tile loop twice with the roles swapped ("X rows, Y columns"):
//   dQ = c (P - I) D      X = Q, Y = D, row-lse
//   dD = c (P - I)^T Q    X = D, Y = Q, column-lse           c = grad * loss_scale / temp
// Every output row has one owner and a fixed summation order (split partials are reduced in
// split order), so gradients are bitwise reproducible.
#include <math_constants.h>

#include "common.cuh"
#include "tensor_core.cuh"
#include "reduce.cuh"

namespace tt {

constexpr int CE_T = 64;      // tile edge (X rows, Y rows)
constexpr int CE_K = 16;      // k-chunk
constexpr int CE_HC = 256;    // output-column chunk held in registers by the backward pass

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// S[64x64] tile = X[x0:x0+64, :] * Y[y0:y0+64, :]^T ; thread (ty,tx) gets rows ty*4+i, cols tx*4+j
template <bool VEC>
__device__ __forceinline__ void s_tile(const float* __restrict__ X, const float* __restrict__ Y,
                                       int64_t x0, int64_t y0, int64_t Bx, int64_t By, int H,
                                       float (&acc)[4][4], float (*Xs)[CE_K][CE_T + 4],
                                       float (*Ys)[CE_K][CE_T + 4]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float rx[4], ry[4];
  auto gload = [&](int k0) {
    const int64_t xr = x0 + lrow, yr = y0 + lrow;
    if (VEC) {
      float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vy = vx;
      if (xr < Bx && k0 + lk < H) vx = __ldg(reinterpret_cast<const float4*>(X + xr * H + k0 + lk));
      if (yr < By && k0 + lk < H) vy = __ldg(reinterpret_cast<const float4*>(Y + yr * H + k0 + lk));
      rx[0] = vx.x; rx[1] = vx.y; rx[2] = vx.z; rx[3] = vx.w;
      ry[0] = vy.x; ry[1] = vy.y; ry[2] = vy.z; ry[3] = vy.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        rx[u] = (xr < Bx && k0 + lk + u < H) ? __ldg(X + xr * H + k0 + lk + u) : 0.f;
        ry[u] = (yr < By && k0 + lk + u < H) ? __ldg(Y + yr * H + k0 + lk + u) : 0.f;
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { Xs[buf][lk + u][lrow] = rx[u]; Ys[buf][lk + u][lrow] = ry[u]; }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < H; k0 += CE_K, buf ^= 1) {
    const bool more = k0 + CE_K < H;
    if (more) gload(k0 + CE_K);
#pragma unroll
    for (int k = 0; k < CE_K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ys[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
  }
}

// ---- forward: per-(row, split) running (max, sum-exp) + the positive logit -------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ D, int64_t Bq, int64_t Bd, int H,
              float inv_temp, int64_t label_offset, int tiles_per_split, float* __restrict__ part_ml,
              float* __restrict__ pos_logit) {
  __shared__ __align__(16) float Xs[2][CE_K][CE_T + 4];
  __shared__ __align__(16) float Ys[2][CE_K][CE_T + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t x0 = (int64_t)blockIdx.x * CE_T;
  const int ntiles = (int)ceil_div(Bd, CE_T);
  const int t_beg = blockIdx.y * tiles_per_split;
  const int t_end = min(ntiles, t_beg + tiles_per_split);
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -CUDART_INF_F; l[i] = 0.f; }
  for (int t = t_beg; t < t_end; ++t) {
    const int64_t y0 = (int64_t)t * CE_T;
    float acc[4][4];
    s_tile<VEC>(Q, D, x0, y0, Bq, Bd, H, acc, Xs, Ys);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t row = x0 + ty * 4 + i;
      const int64_t pcol = row + label_offset;
      float tmax = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t col = y0 + tx * 4 + j;
        const float v = (col < Bd) ? acc[i][j] * inv_temp : -CUDART_INF_F;
        acc[i][j] = v;
        tmax = fmaxf(tmax, v);
        if (col == pcol && row < Bq && col < Bd) pos_logit[row] = v;
      }
      tmax = group16_max(tmax);
      const float mnew = fmaxf(m[i], tmax);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) sum += expf(acc[i][j] - mnew);      // exp(-inf) = 0 for padded cols
      sum = group16_sum(sum);
      l[i] = l[i] * expf(m[i] - mnew) + sum;
      m[i] = mnew;
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t row = x0 + ty * 4 + i;
      if (row < Bq) {
        part_ml[((int64_t)blockIdx.y * Bq + row) * 2 + 0] = m[i];
        part_ml[((int64_t)blockIdx.y * Bq + row) * 2 + 1] = l[i];
      }
    }
  }
}

// combine splits (fixed order) -> lse; loss = loss_scale * sum(lse - pos); pos_mean = mean S_ii
__global__ void __launch_bounds__(1024)
ce_finalize_kernel(const float* __restrict__ part_ml, const float* __restrict__ pos_logit, int nsplit,
                   int64_t Bq, float inv_temp, float loss_scale, float* __restrict__ lse,
                   float* __restrict__ loss, float* __restrict__ pos_mean) {
  __shared__ float s_loss[1024], s_pos[1024];
  pdl_trigger();
  pdl_wait();
  float tl = 0.f, tp = 0.f;
  for (int64_t row = threadIdx.x; row < Bq; row += 1024) {
    // (max, sum) pairs of all splits: loads issued back to back (8 in flight), then combined in split order
    float M = -CUDART_INF_F, Lsum = 0.f;
    for (int s0 = 0; s0 < nsplit; s0 += 8) {
      float2 ml[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        ml[u] = (s0 + u < nsplit) ? __ldg(reinterpret_cast<const float2*>(part_ml) + (int64_t)(s0 + u) * Bq + row)
                                  : make_float2(-CUDART_INF_F, 0.f);
      float Mn = M;
#pragma unroll
      for (int u = 0; u < 8; ++u) Mn = fmaxf(Mn, ml[u].x);
      Lsum *= expf(M - Mn);
#pragma unroll
      for (int u = 0; u < 8; ++u) Lsum += ml[u].y * expf(ml[u].x - Mn);
      M = Mn;
    }
    const float v = M + logf(Lsum);
    lse[row] = v;
    tl += v - pos_logit[row];
    tp += pos_logit[row];
  }
  s_loss[threadIdx.x] = tl; s_pos[threadIdx.x] = tp;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s_loss[threadIdx.x] += s_loss[threadIdx.x + o]; s_pos[threadIdx.x] += s_pos[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *loss = s_loss[0] * loss_scale;
    if (pos_mean) *pos_mean = s_pos[0] / (inv_temp * (float)Bq);
  }
}

int inbatch_finalize(const float* part_ml, const float* pos_logit, int nsplit, int64_t Bq, float inv_temp,
                     float loss_scale, float* lse, float* loss, float* pos_mean, float* /*scratch*/,
                     cudaStream_t s) {
  TT_CUDA(launch_kernel(ce_finalize_kernel, dim3(1), dim3(1024), 0, s, true, part_ml, pos_logit, nsplit, Bq, inv_temp, loss_scale, lse, loss, pos_mean));
  TT_LAUNCH_CHECK("ce_finalize_kernel");
  return TT_OK;
}

// ---- backward ------------------------------------------------------------------------------
// COL_LSE == false: X = Q, Y = D  (lse indexed by X row;  positive at col == row + off)
// COL_LSE == true : X = D, Y = Q  (lse indexed by Y row;  positive at row == col + off)
template <bool VEC, bool COL_LSE>
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ lse,
              int64_t Bx, int64_t By, int H, float inv_temp, int64_t label_offset, int tiles_per_split,
              const float* __restrict__ grad_out, float coef, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float (*Xs)[CE_K][CE_T + 4] = reinterpret_cast<float (*)[CE_K][CE_T + 4]>(smem_raw);
  float (*Ys)[CE_K][CE_T + 4] = Xs + 2;
  float (*Ps)[CE_T + 1] = reinterpret_cast<float (*)[CE_T + 1]>(Ys + 2);
  float (*Y2)[CE_HC + 4] = reinterpret_cast<float (*)[CE_HC + 4]>(
      reinterpret_cast<unsigned char*>(Ps) + ((CE_T * (CE_T + 1) * sizeof(float) + 15) / 16) * 16);

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t x0 = (int64_t)blockIdx.x * CE_T;
  const int ntiles = (int)ceil_div(By, CE_T);
  const int t_beg = blockIdx.y * tiles_per_split;
  const int t_end = min(ntiles, t_beg + tiles_per_split);
  const float scale = coef * (grad_out ? *grad_out : 1.0f);

  for (int hc = 0; hc < H; hc += CE_HC) {
    float O[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) O[i][j] = 0.f;
    for (int t = t_beg; t < t_end; ++t) {
      const int64_t y0 = (int64_t)t * CE_T;
      float acc[4][4];
      s_tile<VEC>(X, Y, x0, y0, Bx, By, H, acc, Xs, Ys);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = x0 + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t col = y0 + tx * 4 + j;
          float p = 0.f;
          if (row < Bx && col < By) {
            const float lv = COL_LSE ? lse[col] : lse[row];
            p = expf(acc[i][j] * inv_temp - lv);
            const bool pos = COL_LSE ? (row == col + label_offset) : (col == row + label_offset);
            if (pos) p -= 1.0f;
          }
          Ps[ty * 4 + i][tx * 4 + j] = p;
        }
      }
      // stage the Y tile (all 64 rows, CE_HC columns) for the second product
      for (int e = tid; e < CE_T * (CE_HC / 4); e += 256) {
        const int k = e / (CE_HC / 4), c4 = (e % (CE_HC / 4)) * 4;
        const int64_t yr = y0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yr < By) {
          if (VEC) {
            if (hc + c4 < H) v = __ldg(reinterpret_cast<const float4*>(Y + yr * H + hc + c4));
          } else {
            float t4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t4[u] = (hc + c4 + u < H) ? __ldg(Y + yr * H + hc + c4 + u) : 0.f;
            v = make_float4(t4[0], t4[1], t4[2], t4[3]);
          }
        }
        *reinterpret_cast<float4*>(&Y2[k][c4]) = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < CE_T; ++k) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Ps[ty * 4 + i][k];
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
          const float4 b = *reinterpret_cast<const float4*>(&Y2[k][jb * 64 + tx * 4]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            O[i][jb * 4 + 0] = fmaf(a[i], b.x, O[i][jb * 4 + 0]);
            O[i][jb * 4 + 1] = fmaf(a[i], b.y, O[i][jb * 4 + 1]);
            O[i][jb * 4 + 2] = fmaf(a[i], b.z, O[i][jb * 4 + 2]);
            O[i][jb * 4 + 3] = fmaf(a[i], b.w, O[i][jb * 4 + 3]);
          }
        }
      }
      __syncthreads();
    }
    float* o = out + (int64_t)blockIdx.y * Bx * H;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t row = x0 + ty * 4 + i;
      if (row >= Bx) continue;
#pragma unroll
      for (int jb = 0; jb < 4; ++jb)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = hc + jb * 64 + tx * 4 + u;
          if (c < H) o[row * H + c] = O[i][jb * 4 + u] * scale;
        }
    }
  }
}

int split_sum(const float* partial, int nsplit, int64_t n, float* out, cudaStream_t s) {
  ReduceJobs jobs{};
  jobs.job[0] = make_job(partial, nsplit, n, n, out);
  jobs.njobs = 1;
  return reduce_parts(jobs, s);
}

static int pick_split(int64_t Bx, int64_t By) {
  const int64_t xt = ceil_div(Bx, CE_T), yt = ceil_div(By, CE_T);
  int64_t s = ceil_div(2 * kNumSMs, xt);
  if (s > yt) s = yt;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  // make every split non-empty
  const int64_t per = ceil_div(yt, s);
  return (int)ceil_div(yt, per);
}

struct CePlan { int ns_f, ns_q, ns_d; size_t ml_bytes, pos_bytes, bwd_bytes, total; };
static CePlan plan_ce(int64_t Bq, int64_t Bd, int H) {
  CePlan p{};
  p.ns_f = pick_split(Bq, Bd);
  p.ns_q = pick_split(Bq, Bd);
  p.ns_d = pick_split(Bd, Bq);
  p.ml_bytes = align_up((size_t)p.ns_f * Bq * 2 * sizeof(float));
  p.pos_bytes = align_up((size_t)Bq * sizeof(float));
  size_t bq = p.ns_q > 1 ? (size_t)p.ns_q * Bq * H * sizeof(float) : 0;
  size_t bd = p.ns_d > 1 ? (size_t)p.ns_d * Bd * H * sizeof(float) : 0;
  p.bwd_bytes = align_up(bq > bd ? bq : bd);
  p.total = p.ml_bytes + p.pos_bytes + p.bwd_bytes + 256;
  return p;
}

constexpr size_t kCeBwdSmem = 2 * 2 * CE_K * (CE_T + 4) * sizeof(float) +
                              ((CE_T * (CE_T + 1) * sizeof(float) + 15) / 16) * 16 +
                              CE_T * (CE_HC + 4) * sizeof(float);

template <bool VEC, bool COL>
static int launch_bwd(const float* X, const float* Y, const float* lse, int64_t Bx, int64_t By, int H,
                      float inv_temp, int64_t off, int nsplit, const float* grad_out, float coef,
                      float* out, float* partial, cudaStream_t s) {
  TT_CUDA(cudaFuncSetAttribute(ce_bwd_kernel<VEC, COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCeBwdSmem));
  const int yt = (int)ceil_div(By, CE_T);
  const int per = (int)ceil_div(yt, nsplit);
  dim3 grid((unsigned)ceil_div(Bx, CE_T), (unsigned)nsplit);
  float* dst = nsplit > 1 ? partial : out;
  ce_bwd_kernel<VEC, COL><<<grid, 256, kCeBwdSmem, s>>>(X, Y, lse, Bx, By, H, inv_temp, off, per, grad_out, coef, dst);
  TT_LAUNCH_CHECK("ce_bwd_kernel");
  if (nsplit > 1) return split_sum(partial, nsplit, Bx * H, out, s);
  return TT_OK;
}

size_t inbatch_ce_fp32_workspace(int64_t Bq, int64_t Bd, int H) { return plan_ce(Bq, Bd, H).total; }

int inbatch_ce_fwd_fp32(const float* q, const float* d, int64_t Bq, int64_t Bd, int H, float inv_temperature,
                        int64_t label_offset, float loss_scale, float* loss, float* lse, float* pos_mean,
                        void* workspace, size_t workspace_bytes, cudaStream_t s) {
  const CePlan plan = plan_ce(Bq, Bd, H);
  if (workspace == nullptr || workspace_bytes < plan.total) { set_error("inbatch_ce_fwd: workspace too small"); return TT_ERR_WORKSPACE; }
  Workspace w(workspace, workspace_bytes);
  float* part_ml = w.take<float>(plan.ml_bytes / sizeof(float));
  float* pos = w.take<float>(Bq);
  const int yt = (int)ceil_div(Bd, CE_T);
  const int per = (int)ceil_div(yt, plan.ns_f);
  dim3 grid((unsigned)ceil_div(Bq, CE_T), (unsigned)plan.ns_f);
  const bool vec = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
  if (vec) ce_fwd_kernel<true><<<grid, 256, 0, s>>>(q, d, Bq, Bd, H, inv_temperature, label_offset, per, part_ml, pos);
  else     ce_fwd_kernel<false><<<grid, 256, 0, s>>>(q, d, Bq, Bd, H, inv_temperature, label_offset, per, part_ml, pos);
  TT_LAUNCH_CHECK("ce_fwd_kernel");
  return inbatch_finalize(part_ml, pos, plan.ns_f, Bq, inv_temperature, loss_scale, lse, loss, pos_mean, nullptr, s);
}

int inbatch_ce_bwd_fp32(const float* q, const float* d, const float* lse, int64_t Bq, int64_t Bd, int H,
                        float inv_temperature, int64_t label_offset, float loss_scale, const float* grad_out,
                        float* dq, float* dd, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  const CePlan plan = plan_ce(Bq, Bd, H);
  if (workspace == nullptr || workspace_bytes < plan.total) { set_error("inbatch_ce_bwd: workspace too small"); return TT_ERR_WORKSPACE; }
  Workspace w(workspace, workspace_bytes);
  (void)w.take<float>(plan.ml_bytes / sizeof(float));
  (void)w.take<float>(Bq);
  float* partial = plan.bwd_bytes ? w.take<float>(plan.bwd_bytes / sizeof(float)) : nullptr;
  const float coef = loss_scale * inv_temperature;
  const bool vec = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(d)) & 15) == 0;
  int rc = TT_OK;
  if (dq) {
    rc = vec ? launch_bwd<true, false>(q, d, lse, Bq, Bd, H, inv_temperature, label_offset, plan.ns_q, grad_out, coef, dq, partial, s)
             : launch_bwd<false, false>(q, d, lse, Bq, Bd, H, inv_temperature, label_offset, plan.ns_q, grad_out, coef, dq, partial, s);
    if (rc) return rc;
  }
  if (dd) {
    rc = vec ? launch_bwd<true, true>(d, q, lse, Bd, Bq, H, inv_temperature, label_offset, plan.ns_d, grad_out, coef, dd, partial, s)
             : launch_bwd<false, true>(d, q, lse, Bd, Bq, H, inv_temperature, label_offset, plan.ns_d, grad_out, coef, dd, partial, s);
  }
  return rc;
}

}  // namespace tt

extern "C" {

size_t tt_inbatch_ce_workspace(int64_t Bq, int64_t Bd, int H, int precision) {
  if (Bq <= 0 || Bd <= 0 || H <= 0) return 256;
  size_t fp32 = tt::plan_ce(Bq, Bd, H).total;
  if (precision == TT_PREC_BF16) return tt::tc_inbatch_workspace(Bq, Bd, H);
  return fp32;
}

int tt_inbatch_ce_fwd(const float* q, const float* d, const void* q_bf16, const void* d_bf16, int64_t Bq,
                      int64_t Bd, int H, float inv_temperature, int64_t label_offset, float loss_scale,
                      float* loss, float* lse, float* pos_mean, int precision, void* workspace,
                      size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q && d && loss && lse && Bq > 0 && Bd > 0 && H > 0, "inbatch_ce_fwd: bad arguments");
  TT_CHECK_ARG(label_offset >= 0 && Bq + label_offset <= Bd,
               "inbatch_ce_fwd: positives out of range (Bq=%lld off=%lld Bd=%lld)", (long long)Bq,
               (long long)label_offset, (long long)Bd);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == TT_PREC_BF16)
    return tt::tc_inbatch_fwd(q, d, (const __nv_bfloat16*)q_bf16, (const __nv_bfloat16*)d_bf16, Bq, Bd, H,
                              inv_temperature, label_offset, loss_scale, loss, lse, pos_mean, workspace,
                              workspace_bytes, s);
  TT_CHECK_ARG(precision == TT_PREC_FP32, "inbatch_ce_fwd: unknown precision %d", precision);
  return tt::inbatch_ce_fwd_fp32(q, d, Bq, Bd, H, inv_temperature, label_offset, loss_scale, loss, lse, pos_mean,
                                 workspace, workspace_bytes, s);
}

int tt_inbatch_ce_bwd(const float* q, const float* d, const void* q_bf16, const void* d_bf16, const float* lse,
                      int64_t Bq, int64_t Bd, int H, float inv_temperature, int64_t label_offset,
                      float loss_scale, const float* grad_out, float* dq, float* dd, int precision,
                      void* workspace, size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q && d && lse && Bq > 0 && Bd > 0 && H > 0, "inbatch_ce_bwd: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == TT_PREC_BF16)
    return tt::tc_inbatch_bwd(q, d, (const __nv_bfloat16*)q_bf16, (const __nv_bfloat16*)d_bf16, lse, Bq, Bd, H,
                              inv_temperature, label_offset, loss_scale, grad_out, dq, dd, workspace,
                              workspace_bytes, s);
  TT_CHECK_ARG(precision == TT_PREC_FP32, "inbatch_ce_bwd: unknown precision %d", precision);
  return tt::inbatch_ce_bwd_fp32(q, d, lse, Bq, Bd, H, inv_temperature, label_offset, loss_scale, grad_out, dq, dd,
                                 workspace, workspace_bytes, s);
}

int tt_inbatch_ce_bwd_nparts(int64_t Bq, int64_t Bd, int H, int precision) {
  if (precision != TT_PREC_BF16 || Bq <= 0 || Bd <= 0 || H <= 0) return 1;
  return tt::tc_inbatch_bwd_nparts(Bq, Bd, H);
}

int tt_inbatch_ce_bwd_parts(const void* q_bf16, const void* d_bf16, const float* lse, int64_t Bq, int64_t Bd, int H,
                            float inv_temperature, int64_t label_offset, float loss_scale, const float* grad_out,
                            float* dq_parts, int64_t dq_part_stride, float* dd_parts, int64_t dd_part_stride,
                            void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q_bf16 && d_bf16 && lse && Bq > 0 && Bd > 0 && H > 0 && (dq_parts || dd_parts), "inbatch_ce_bwd_parts: bad arguments");
  return tt::tc_inbatch_bwd_parts((const __nv_bfloat16*)q_bf16, (const __nv_bfloat16*)d_bf16, lse, Bq, Bd, H,
                                  inv_temperature, label_offset, loss_scale, grad_out, dq_parts, dq_part_stride,
                                  dd_parts, dd_part_stride, static_cast<cudaStream_t>(stream));
}

size_t tt_inbatch_ce_fwd_ex_workspace(int64_t Bq, int64_t Bd) {
  if (Bq <= 0 || Bd <= 0) return 256;
  return tt::tc_inbatch_fwd_ex_workspace(Bq, Bd);
}

int tt_inbatch_ce_fwd_ex(const void* q_bf16, int64_t Bq, const void* d_bf16, int64_t Bd, int64_t d_buf_rows, int64_t d_blk,
                         int64_t d_blk_stride, int64_t d_blk_off, int H, float inv_temperature, int64_t label_offset,
                         float loss_scale, float* loss, float* lse, float* pos_mean, void* workspace,
                         size_t workspace_bytes, void* sync_scratch, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q_bf16 && d_bf16 && loss && lse && Bq > 0 && Bd > 0 && H > 0 && d_buf_rows >= 1 && d_blk >= 1,
               "inbatch_ce_fwd_ex: bad arguments");
  TT_CHECK_ARG(H % 64 == 0 && H <= 256 && (d_blk % 128 == 0 || d_blk >= Bd), "inbatch_ce_fwd_ex: needs H %% 64 == 0, H <= 256, d_blk %% 128 == 0");
  TT_CHECK_ARG(label_offset >= 0 && Bq + label_offset <= Bd, "inbatch_ce_fwd_ex: positives out of range");
  const size_t need = tt::tc_inbatch_fwd_ex_workspace(Bq, Bd);
  if (workspace == nullptr || workspace_bytes < need) { tt::set_error("inbatch_ce_fwd_ex: workspace too small"); return TT_ERR_WORKSPACE; }
  tt::Workspace w(workspace, workspace_bytes);
  float* pos = w.take<float>(Bq);
  float* part_ml = w.take<float>((need - 256 - tt::align_up((size_t)Bq * 4)) / 4);
  return tt::tc_inbatch_fwd_ex((const __nv_bfloat16*)q_bf16, Bq, (const __nv_bfloat16*)d_bf16, Bd, d_buf_rows, d_blk,
                               d_blk_stride, d_blk_off, H, inv_temperature, label_offset, loss_scale, loss, lse, pos_mean,
                               part_ml, pos, sync_scratch, static_cast<cudaStream_t>(stream));
}

size_t tt_inbatch_ce_sync_bytes(int64_t Bq) { return Bq > 0 ? tt::tc_inbatch_fwd_sync_bytes(Bq) : 16; }

int tt_inbatch_ce_bwd_nparts_ex(int64_t q_x_rows, int64_t q_y_rows, int64_t d_x_rows, int64_t d_y_rows, int H) {
  if (q_x_rows <= 0 || q_y_rows <= 0 || d_x_rows <= 0 || d_y_rows <= 0 || H <= 0) return 1;
  return tt::tc_inbatch_bwd_nparts2(q_x_rows, q_y_rows, d_x_rows, d_y_rows, H);
}

int tt_inbatch_ce_bwd_fused_ok(int64_t q_x_rows, int64_t q_y_rows, int64_t d_x_rows, int64_t d_y_rows, int H) {
  if (q_x_rows <= 0 || q_y_rows <= 0 || d_x_rows <= 0 || d_y_rows <= 0 || H <= 0 || H % 64 != 0 || H > 256) return 0;
  return tt::tc_inbatch_bwd_nparts2(q_x_rows, q_y_rows, d_x_rows, d_y_rows, H) <= 2 ? 1 : 0;
}

size_t tt_inbatch_ce_onepass_sync_bytes(int64_t Bq) { return Bq > 0 ? tt::tc_inbatch_onepass_sync_bytes(Bq) : 16; }

int tt_inbatch_ce_onepass_ok(int64_t Bq, int64_t Bd, int H, float logit_bound) {
  if (Bq <= 0 || Bd <= 0 || H <= 0) return 0;
  return tt::tc_inbatch_onepass_ok(Bq, Bd, H, logit_bound);
}

int tt_inbatch_ce_fwd_dq(const tt_ce_pass_t* q_pass, int H, float inv_temperature, float logit_bound, float loss_scale,
                         const float* grad_out, float* loss, float* lse, float* pos_mean, void* sync_scratch, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q_pass && H > 0 && loss && lse && sync_scratch, "inbatch_ce_fwd_dq: bad arguments");
  TT_CHECK_ARG(q_pass->x_bf16 && q_pass->y_bf16 && q_pass->x_rows > 0 && q_pass->y_rows > 0 && q_pass->y_buf_rows >= 1,
               "inbatch_ce_fwd_dq: null operand");
  TT_CHECK_ARG(q_pass->label_offset >= 0 && q_pass->x_rows + q_pass->label_offset <= q_pass->y_rows,
               "inbatch_ce_fwd_dq: positives out of range");
  return tt::tc_inbatch_fwd_dq(q_pass, H, inv_temperature, logit_bound, loss_scale, grad_out, loss, lse, pos_mean, sync_scratch,
                               nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int tt_inbatch_ce_stash_ok(int64_t Bq, int64_t Bd, int H) {
  if (Bq <= 0 || Bd <= 0 || H <= 0) return 0;
  return tt::tc_inbatch_stash_ok(Bq, Bd, H);
}
size_t tt_inbatch_ce_stash_bytes(int64_t Bq, int64_t Bd, int H) {
  return (Bq > 0 && Bd > 0 && H > 0) ? tt::tc_inbatch_stash_bytes(Bq, Bd, H) : 0;
}

int tt_inbatch_ce_fwd_dq_stash(const tt_ce_pass_t* q_pass, int H, float inv_temperature, float logit_bound, float loss_scale,
                               const float* grad_out, float* loss, float* lse, float* pos_mean, void* sync_scratch, void* stash,
                               void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q_pass && H > 0 && loss && lse && sync_scratch && stash, "inbatch_ce_fwd_dq_stash: bad arguments");
  TT_CHECK_ARG(q_pass->x_bf16 && q_pass->y_bf16 && q_pass->x_rows > 0 && q_pass->y_rows > 0 && q_pass->y_buf_rows >= 1,
               "inbatch_ce_fwd_dq_stash: null operand");
  TT_CHECK_ARG(q_pass->label_offset >= 0 && q_pass->x_rows + q_pass->label_offset <= q_pass->y_rows,
               "inbatch_ce_fwd_dq_stash: positives out of range");
  return tt::tc_inbatch_fwd_dq(q_pass, H, inv_temperature, logit_bound, loss_scale, grad_out, loss, lse, pos_mean, sync_scratch,
                               nullptr, nullptr, static_cast<cudaStream_t>(stream), stash);
}

int tt_inbatch_ce_dd_stash(const tt_ce_pass_t* d_pass, int H, float inv_temperature, float loss_scale, const float* grad_out,
                           const void* stash, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(d_pass && H > 0 && d_pass->x_bf16 && d_pass->y_bf16 && d_pass->x_rows > 0 && d_pass->y_rows > 0 && stash,
               "inbatch_ce_dd_stash: bad arguments");
  TT_CHECK_ARG(d_pass->dz_bf16 ? (d_pass->dz_colsum && d_pass->inv_norm) : d_pass->out_parts != nullptr,
               "inbatch_ce_dd_stash: needs dz_bf16 + dz_colsum + inv_norm, or out_parts");
  return tt::tc_inbatch_dd_stored(d_pass, H, inv_temperature, loss_scale, grad_out, stash, static_cast<cudaStream_t>(stream));
}

int tt_inbatch_ce_fwd_dq_p2p(const tt_ce_pass_t* q_pass, int H, float inv_temperature, float logit_bound, float loss_scale,
                             const float* grad_out, float* loss, float* lse, float* pos_mean, void* sync_scratch,
                             const tt_p2p_t* y_exchange, const void* y_own, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q_pass && H > 0 && loss && lse && sync_scratch && y_exchange && y_own, "inbatch_ce_fwd_dq_p2p: bad arguments");
  TT_CHECK_ARG(q_pass->x_bf16 && q_pass->y_bf16 && q_pass->x_rows > 0 && q_pass->y_rows > 0 && q_pass->y_buf_rows >= 1,
               "inbatch_ce_fwd_dq_p2p: null operand");
  TT_CHECK_ARG(q_pass->label_offset >= 0 && q_pass->x_rows + q_pass->label_offset <= q_pass->y_rows,
               "inbatch_ce_fwd_dq_p2p: positives out of range");
  return tt::tc_inbatch_fwd_dq(q_pass, H, inv_temperature, logit_bound, loss_scale, grad_out, loss, lse, pos_mean, sync_scratch,
                               y_exchange, y_own, static_cast<cudaStream_t>(stream));
}

int tt_inbatch_ce_onepass(const tt_ce_pass_t* q_pass, const tt_ce_pass_t* d_pass, int H, float inv_temperature, float logit_bound,
                          float loss_scale, const float* grad_out, float* loss, float* lse, float* pos_mean, void* sync_scratch,
                          void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q_pass && d_pass && H > 0 && loss && lse && sync_scratch, "inbatch_ce_onepass: bad arguments");
  TT_CHECK_ARG(q_pass->x_bf16 && q_pass->y_bf16 && d_pass->x_bf16 && d_pass->y_bf16 && q_pass->x_rows > 0, "inbatch_ce_onepass: null operand");
  return tt::tc_inbatch_onepass_single(q_pass, d_pass, H, inv_temperature, logit_bound, loss_scale, grad_out, loss, lse, pos_mean,
                                       sync_scratch, static_cast<cudaStream_t>(stream));
}

int tt_inbatch_ce_dd_nparts(int64_t d_x_rows, int64_t d_y_rows, int H) {
  if (d_x_rows <= 0 || d_y_rows <= 0 || H <= 0) return 1;
  return tt::tc_inbatch_dd_nparts(d_x_rows, d_y_rows);
}

int tt_inbatch_ce_dd(const tt_ce_pass_t* d_pass, int H, float inv_temperature, float loss_scale, const float* grad_out,
                     void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(d_pass && H > 0 && d_pass->x_bf16 && d_pass->y_bf16 && d_pass->lse && d_pass->x_rows > 0 && d_pass->y_rows > 0,
               "inbatch_ce_dd: bad arguments");
  TT_CHECK_ARG(d_pass->dz_bf16 ? (d_pass->dz_colsum && d_pass->inv_norm) : d_pass->out_parts != nullptr,
               "inbatch_ce_dd: needs dz_bf16 + dz_colsum + inv_norm, or out_parts");
  return tt::tc_inbatch_dd(d_pass, H, inv_temperature, loss_scale, grad_out, static_cast<cudaStream_t>(stream));
}

int tt_inbatch_ce_bwd_parts_ex(const tt_ce_pass_t* q_pass, const tt_ce_pass_t* d_pass, int H, float inv_temperature,
                               float loss_scale, const float* grad_out, int nparts, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q_pass && d_pass && H > 0 && nparts >= 1, "inbatch_ce_bwd_parts_ex: bad arguments");
  TT_CHECK_ARG(q_pass->x_bf16 && q_pass->y_bf16 && d_pass->x_bf16 && d_pass->y_bf16 && q_pass->lse && d_pass->lse,
               "inbatch_ce_bwd_parts_ex: null operand");
  return tt::tc_inbatch_bwd_parts_ex(q_pass, d_pass, H, inv_temperature, loss_scale, grad_out, nparts,
                                     static_cast<cudaStream_t>(stream));
}

}  // extern "C"
