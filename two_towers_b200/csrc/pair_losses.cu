// pair_losses.cu -- K5 / K6: row-paired cosine losses (triplet margin, multiple negatives).
//
// Reference: contrastive_triplet_loss, twotower/losses.py:28-35
//              relu(margin - cos(q,p) + cos(q,n)).mean(), F.cosine_similarity eps 1e-8;
//            multiple_negatives_loss, twotower/losses.py:65-83
//              cos(q, [p; negs]) / temperature -> cross_entropy(label 0).
// One warp per batch row; the batch mean is a fixed-order two-level reduction.
#include "common.cuh"

namespace tt {

constexpr float kCosEps = 1e-8f;

struct Dot3 { float xy, xx, yy; };
__device__ __forceinline__ Dot3 warp_dot3(const float* __restrict__ x, const float* __restrict__ y, int H, int lane) {
  Dot3 r{0.f, 0.f, 0.f};
  for (int e = lane; e < H; e += 32) {
    const float a = x[e], b = y[e];
    r.xy = fmaf(a, b, r.xy); r.xx = fmaf(a, a, r.xx); r.yy = fmaf(b, b, r.yy);
  }
  r.xy = warp_sum(r.xy); r.xx = warp_sum(r.xx); r.yy = warp_sum(r.yy);
  return r;
}

// fixed-order sum of v[0..n) * scale -> out[0]; up to 3 vectors at once
__global__ void __launch_bounds__(1024)
reduce3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
               int64_t n, int stride, float scale, float* __restrict__ oa, float* __restrict__ ob,
               float* __restrict__ oc) {
  __shared__ float s[3][1024];
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    if (a) t0 += a[i * stride];
    if (b) t1 += b[i * stride];
    if (c) t2 += c[i * stride];
  }
  s[0][threadIdx.x] = t0; s[1][threadIdx.x] = t1; s[2][threadIdx.x] = t2;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s[0][threadIdx.x] += s[0][threadIdx.x + o];
      s[1][threadIdx.x] += s[1][threadIdx.x + o];
      s[2][threadIdx.x] += s[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (oa) *oa = s[0][0] * scale;
    if (ob) *ob = s[1][0] * scale;
    if (oc) *oc = s[2][0] * scale;
  }
}

// sims[row] = (cos_qp, cos_qn); rowloss[row] = relu(margin - cos_qp + cos_qn)
__global__ void __launch_bounds__(256)
triplet_fwd_kernel(const float* __restrict__ q, const float* __restrict__ p, const float* __restrict__ n,
                   int64_t B, int H, float margin, float* __restrict__ sims, float* __restrict__ rowloss) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const Dot3 a = warp_dot3(q + row * H, p + row * H, H, lane);
  const Dot3 b = warp_dot3(q + row * H, n + row * H, H, lane);
  const float nq = fmaxf(sqrtf(a.xx), kCosEps);
  const float sp = a.xy / (nq * fmaxf(sqrtf(a.yy), kCosEps));
  const float sn = b.xy / (nq * fmaxf(sqrtf(b.yy), kCosEps));
  if (lane == 0) {
    sims[2 * row] = sp; sims[2 * row + 1] = sn;
    rowloss[row] = fmaxf(margin - sp + sn, 0.f);
  }
}

// d cos(x,y)/dx = y/(nx ny) - c x/nx^2
__global__ void __launch_bounds__(256)
triplet_bwd_kernel(const float* __restrict__ q, const float* __restrict__ p, const float* __restrict__ n,
                   const float* __restrict__ sims, int64_t B, int H, float margin,
                   const float* __restrict__ grad_out, float* __restrict__ dq, float* __restrict__ dp,
                   float* __restrict__ dn) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* qr = q + row * H; const float* pr = p + row * H; const float* nr = n + row * H;
  float qq = 0.f, pp = 0.f, nn = 0.f;
  for (int e = lane; e < H; e += 32) { qq = fmaf(qr[e], qr[e], qq); pp = fmaf(pr[e], pr[e], pp); nn = fmaf(nr[e], nr[e], nn); }
  qq = warp_sum(qq); pp = warp_sum(pp); nn = warp_sum(nn);
  const float nq = fmaxf(sqrtf(qq), kCosEps), np_ = fmaxf(sqrtf(pp), kCosEps), nn_ = fmaxf(sqrtf(nn), kCosEps);
  const float sp = sims[2 * row], sn = sims[2 * row + 1];
  const float g = ((margin - sp + sn) > 0.f ? 1.0f : 0.f) * (grad_out ? *grad_out : 1.0f) / (float)B;
  // loss_i = margin - sp + sn  ->  d/dsp = -g, d/dsn = +g
  for (int e = lane; e < H; e += 32) {
    const float qe = qr[e], pe = pr[e], ne = nr[e];
    const float dsp_dq = pe / (nq * np_) - sp * qe / (nq * nq);
    const float dsn_dq = ne / (nq * nn_) - sn * qe / (nq * nq);
    if (dq) dq[row * H + e] = g * (dsn_dq - dsp_dq);
    if (dp) dp[row * H + e] = -g * (qe / (nq * np_) - sp * pe / (np_ * np_));
    if (dn) dn[row * H + e] = g * (qe / (nq * nn_) - sn * ne / (nn_ * nn_));
  }
}

constexpr int kMaxNeg = 63;     // N+1 <= 64 candidates per row

// probs[row, 0..N] = softmax(cos(q,[p;negs]) * inv_temp); rowloss = -log probs[row,0]
__global__ void __launch_bounds__(256)
multineg_fwd_kernel(const float* __restrict__ q, const float* __restrict__ p, const float* __restrict__ negs,
                    int64_t B, int N, int H, float inv_temp, float* __restrict__ probs,
                    float* __restrict__ rowloss) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  float logit[kMaxNeg + 1];
  float mx = -3.0e38f;
  for (int c = 0; c <= N; ++c) {
    const float* y = (c == 0) ? p + row * H : negs + ((int64_t)row * N + (c - 1)) * H;
    const Dot3 a = warp_dot3(q + row * H, y, H, lane);
    const float v = a.xy / (fmaxf(sqrtf(a.xx), kCosEps) * fmaxf(sqrtf(a.yy), kCosEps)) * inv_temp;
    logit[c] = v;
    mx = fmaxf(mx, v);
  }
  float sum = 0.f;
  for (int c = 0; c <= N; ++c) sum += expf(logit[c] - mx);
  const float lse = mx + logf(sum);
  if (lane == 0) {
    for (int c = 0; c <= N; ++c) probs[row * (N + 1) + c] = expf(logit[c] - lse);
    rowloss[row] = lse - logit[0];
  }
}

__global__ void __launch_bounds__(256)
multineg_bwd_kernel(const float* __restrict__ q, const float* __restrict__ p, const float* __restrict__ negs,
                    const float* __restrict__ probs, int64_t B, int N, int H, float inv_temp,
                    const float* __restrict__ grad_out, float* __restrict__ dq, float* __restrict__ dp,
                    float* __restrict__ dnegs) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* qr = q + row * H;
  const float gs = (grad_out ? *grad_out : 1.0f) * inv_temp / (float)B;
  float qq = 0.f;
  for (int e = lane; e < H; e += 32) qq = fmaf(qr[e], qr[e], qq);
  const float nq = fmaxf(sqrtf(warp_sum(qq)), kCosEps);
  // first pass over candidates: write d(candidate); accumulate dq in registers (H <= 32*16)
  constexpr int kMaxPerLane = 16;
  float dqa[kMaxPerLane];
#pragma unroll
  for (int u = 0; u < kMaxPerLane; ++u) dqa[u] = 0.f;
  for (int c = 0; c <= N; ++c) {
    const float* y = (c == 0) ? p + row * H : negs + ((int64_t)row * N + (c - 1)) * H;
    float* dy = (c == 0) ? (dp ? dp + row * H : nullptr) : (dnegs ? dnegs + ((int64_t)row * N + (c - 1)) * H : nullptr);
    const Dot3 a = warp_dot3(qr, y, H, lane);
    const float ny = fmaxf(sqrtf(a.yy), kCosEps);
    const float cs = a.xy / (nq * ny);
    const float g = (probs[row * (N + 1) + c] - (c == 0 ? 1.0f : 0.f)) * gs;     // dL/dcos_c
#pragma unroll
    for (int u = 0; u < kMaxPerLane; ++u) {
      const int e = lane + 32 * u;
      if (e < H) {
        const float qe = qr[e], ye = y[e];
        dqa[u] = fmaf(g, ye / (nq * ny) - cs * qe / (nq * nq), dqa[u]);
        if (dy) dy[e] = g * (qe / (nq * ny) - cs * ye / (ny * ny));
      }
    }
  }
  if (dq) {
#pragma unroll
    for (int u = 0; u < kMaxPerLane; ++u) {
      const int e = lane + 32 * u;
      if (e < H) dq[row * H + e] = dqa[u];
    }
  }
}

static inline unsigned rows_grid(int64_t B) { return (unsigned)ceil_div(B, 8); }

}  // namespace tt

extern "C" {

int tt_triplet_fwd(const float* q, const float* p, const float* n, int64_t B, int H, float margin, float* loss,
                   float* sims, float* pos_mean, float* neg_mean, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q && p && n && loss && sims && B > 0 && H > 0, "triplet_fwd: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* rowloss = sims + 2 * B;        // sims buffer is 3*B floats: [B,2] cosines, then [B] row losses
  tt::triplet_fwd_kernel<<<tt::rows_grid(B), 256, 0, s>>>(q, p, n, B, H, margin, sims, rowloss);
  TT_LAUNCH_CHECK("triplet_fwd_kernel");
  tt::reduce3_kernel<<<1, 1024, 0, s>>>(rowloss, nullptr, nullptr, B, 1, 1.0f / (float)B, loss, nullptr, nullptr);
  TT_LAUNCH_CHECK("reduce3_kernel");
  if (pos_mean || neg_mean) {
    tt::reduce3_kernel<<<1, 1024, 0, s>>>(nullptr, sims, sims + 1, B, 2, 1.0f / (float)B, nullptr, pos_mean, neg_mean);
    TT_LAUNCH_CHECK("reduce3_kernel");
  }
  return TT_OK;
}

int tt_triplet_bwd(const float* q, const float* p, const float* n, const float* sims, int64_t B, int H,
                   float margin, const float* grad_out, float* dq, float* dp, float* dn, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q && p && n && sims && B > 0 && H > 0, "triplet_bwd: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  tt::triplet_bwd_kernel<<<tt::rows_grid(B), 256, 0, s>>>(q, p, n, sims, B, H, margin, grad_out, dq, dp, dn);
  TT_LAUNCH_CHECK("triplet_bwd_kernel");
  return TT_OK;
}

int tt_multineg_fwd(const float* q, const float* p, const float* negs, int64_t B, int N, int H,
                    float inv_temperature, float* loss, float* probs, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q && p && negs && loss && probs && B > 0 && N > 0 && H > 0, "multineg_fwd: bad arguments");
  TT_CHECK_ARG(N <= tt::kMaxNeg, "multineg_fwd: at most %d negatives per row", tt::kMaxNeg);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* rowloss = probs + (size_t)B * (N + 1);
  tt::multineg_fwd_kernel<<<tt::rows_grid(B), 256, 0, s>>>(q, p, negs, B, N, H, inv_temperature, probs, rowloss);
  TT_LAUNCH_CHECK("multineg_fwd_kernel");
  tt::reduce3_kernel<<<1, 1024, 0, s>>>(rowloss, nullptr, nullptr, B, 1, 1.0f / (float)B, loss, nullptr, nullptr);
  TT_LAUNCH_CHECK("reduce3_kernel");
  return TT_OK;
}

int tt_multineg_bwd(const float* q, const float* p, const float* negs, const float* probs, int64_t B, int N,
                    int H, float inv_temperature, const float* grad_out, float* dq, float* dp, float* dnegs,
                    void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(q && p && negs && probs && B > 0 && N > 0 && H > 0, "multineg_bwd: bad arguments");
  TT_CHECK_ARG(N <= tt::kMaxNeg && H <= 512, "multineg_bwd: N <= %d and H <= 512 supported", tt::kMaxNeg);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  tt::multineg_bwd_kernel<<<tt::rows_grid(B), 256, 0, s>>>(q, p, negs, probs, B, N, H, inv_temperature, grad_out, dq, dp, dnegs);
  TT_LAUNCH_CHECK("multineg_bwd_kernel");
  return TT_OK;
}

}  // extern "C"
