// optim.cu -- fused AdamW over a flat fp32 parameter buffer.
//
// Reference: torch.optim.AdamW(model.parameters(), lr) at twotower/train.py:359 with the
// defaults betas=(0.9,0.999), eps=1e-8, weight_decay=0.01, amsgrad=False; semantics of ATen
// _single_tensor_adamw:  p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
// The step counter lives on the device so the whole training step can sit in a CUDA graph.
#include "common.cuh"

namespace tt {

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, float lr, float beta1, float beta2, float eps, float wd,
             const int64_t* __restrict__ step_count, __nv_bfloat16* __restrict__ p_bf16) {
  __shared__ float s_step_size, s_inv_sqrt_bc2;
  if (threadIdx.x == 0) {
    const double t = (double)(*step_count + 1);
    const double bc1 = 1.0 - pow((double)beta1, t);
    const double bc2 = 1.0 - pow((double)beta2, t);
    s_step_size = (float)((double)lr / bc1);
    s_inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step_size, inv_sqrt_bc2 = s_inv_sqrt_bc2;
  const float decay = 1.0f - lr * wd;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = p[i] * decay;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16(pi);
  }
}

__global__ void step_inc_kernel(int64_t* step_count) { *step_count += 1; }

}  // namespace tt

extern "C" int tt_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             float lr, float beta1, float beta2, float eps, float weight_decay,
                             int64_t* step_count, void* param_bf16, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step_count && n >= 0, "adamw_step: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n > 0) {
    int64_t blocks = tt::ceil_div(n, 256);
    if (blocks > 8 * tt::kNumSMs) blocks = 8 * tt::kNumSMs;
    tt::adamw_kernel<<<(unsigned)blocks, 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                      weight_decay, step_count, (__nv_bfloat16*)param_bf16);
    TT_LAUNCH_CHECK("adamw_kernel");
  }
  tt::step_inc_kernel<<<1, 1, 0, s>>>(step_count);
  TT_LAUNCH_CHECK("step_inc_kernel");
  return TT_OK;
}
