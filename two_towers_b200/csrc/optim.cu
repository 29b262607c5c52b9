// optim.cu -- fused AdamW over a flat fp32 parameter buffer.
//
// Reference: torch.optim.AdamW(model.parameters(), lr) at twotower/train.py:359 with the
// defaults betas=(0.9,0.999), eps=1e-8, weight_decay=0.01, amsgrad=False; semantics of ATen
// _single_tensor_adamw:  p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
// The step counter lives on the device so the whole training step can sit in a CUDA graph.
#include "common.cuh"

namespace tt {

// gradient source of one AdamW launch: a plain buffer, or the slots of a peer-memory all-gather (tt_p2p_t, double
// buffered) that are summed in rank order on the fly -- the all-reduce's reduction costs no launch of its own
struct GradSlots {
  const unsigned char* base;   // local exchange buffer (null: plain gradient)
  int world;
  size_t slot_floats;
  float* sum_out;              // nullable: the summed gradient is also written here (p.grad stays meaningful)
  // a second gradient for one segment of the flat buffer (parameters extra_lo .. extra_lo + extra_n, both multiples of 4):
  // the embedding-table gradient of the tower whose backward ran concurrently on a second stream (FusedTrainer, untied
  // towers) -- added after `g`, i.e. in the order the serial schedule accumulates it
  const float* extra;
  int64_t extra_lo, extra_n;
};

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, double lr, double beta1, double beta2, double eps, double wd,
             int64_t* step_count, __nv_bfloat16* __restrict__ p_bf16, int vec_ok,
             const float* __restrict__ publish_src, float* publish_dst, const GradSlots gs) {
  // scalars are formed in double (as Python does in torch.optim) and rounded to fp32 once
  __shared__ float s_neg_step_size, s_sqrt_bc2;
  pdl_trigger();
  pdl_wait();
  // the step's loss leaves for the host from here (publish_dst: mapped pinned memory): no copy node in the captured step
  if (publish_dst && blockIdx.x == 0 && threadIdx.x == 0) *publish_dst = *publish_src;
  if (threadIdx.x == 0) {
    const double t = (double)(*step_count + 1);
    const double bc1 = 1.0 - pow(beta1, t);
    const double bc2 = 1.0 - pow(beta2, t);
    s_neg_step_size = (float)(-(lr / bc1));
    s_sqrt_bc2 = (float)sqrt(bc2);
  }
  __syncthreads();
  const float neg_step_size = s_neg_step_size, sqrt_bc2 = s_sqrt_bc2;
  const float decay = (float)(1.0 - lr * wd);
  const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), epsf = (float)eps;
  auto update = [&](float gi, float& pi, float& mi, float& vi) {
    pi = pi * decay;                                           // param.mul_(1 - lr*wd)
    mi = (w1 < 0.5f) ? mi + w1 * (gi - mi) : gi - (gi - mi) * (1.0f - w1);   // exp_avg.lerp_(grad, 1-beta1)
    vi = vi * b2;                                              // exp_avg_sq.mul_(beta2)
    vi = vi + (w2 * gi) * gi;                                  //   .addcmul_(grad, grad, value=1-beta2)
    const float denom = sqrtf(vi) / sqrt_bc2 + epsf;           // sqrt / bias_correction2_sqrt + eps
    pi = pi + neg_step_size * (mi / denom);                    // addcdiv_(exp_avg, denom, value=-step_size)
  };
  // 16-byte accesses (the flat buffers are 256-byte aligned): 28 bytes move per parameter, so a 120 M-row word-embedding
  // table is 3.6 GB per step -- with 4-byte accesses this loop ran at half the HBM rate
  const int64_t n4 = vec_ok ? n / 4 : 0;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const float* slots = nullptr;
  if (gs.base) {                                             // the round that has just completed selects the slot set
    const unsigned round = reinterpret_cast<const unsigned*>(gs.base)[32];
    slots = reinterpret_cast<const float*>(gs.base + 256) + ((round & 1u) ? (size_t)gs.world * gs.slot_floats : 0);
  }
  for (int64_t i = tid; i < n4; i += nth) {
    float4 g4;
    if (slots) {                                             // rank order: bitwise identical on every rank
      g4 = __ldg(reinterpret_cast<const float4*>(slots) + i);
      for (int r = 1; r < gs.world; ++r) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(slots + (size_t)r * gs.slot_floats) + i);
        g4.x += t.x; g4.y += t.y; g4.z += t.z; g4.w += t.w;
      }
      if (gs.sum_out) reinterpret_cast<float4*>(gs.sum_out)[i] = g4;
    } else {
      g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    }
    if (gs.extra) {
      const int64_t e = i * 4 - gs.extra_lo;
      if (e >= 0 && e < gs.extra_n) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(gs.extra + e));
        g4.x += t.x; g4.y += t.y; g4.z += t.z; g4.w += t.w;
        if (gs.sum_out) reinterpret_cast<float4*>(gs.sum_out)[i] = g4;
      }
    }
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    update(g4.x, p4.x, m4.x, v4.x); update(g4.y, p4.y, m4.y, v4.y);
    update(g4.z, p4.z, m4.z, v4.z); update(g4.w, p4.w, m4.w, v4.w);
    reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
    if (p_bf16) reinterpret_cast<uint2*>(p_bf16)[i] = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += nth) {
    float pi = p[i], mi = m[i], vi = v[i];
    float gi;
    if (slots) {
      gi = slots[i];
      for (int r = 1; r < gs.world; ++r) gi += slots[(size_t)r * gs.slot_floats + i];
      if (gs.sum_out) gs.sum_out[i] = gi;
    } else {
      gi = g[i];
    }
    if (gs.extra && i >= gs.extra_lo && i < gs.extra_lo + gs.extra_n) {
      gi += gs.extra[i - gs.extra_lo];
      if (gs.sum_out) gs.sum_out[i] = gi;
    }
    update(gi, pi, mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16(pi);
  }
  // every block has read step_count[0] above; the LAST block to arrive advances the counter, so the
  // increment needs no extra launch and the whole optimizer step is one graph node.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(step_count + 1), 1ull);
    if (t == (unsigned long long)gridDim.x - 1ull) {
      step_count[1] = 0;
      step_count[0] += 1;
    }
  }
}

}  // namespace tt

static int adamw_launch(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                        double lr, double beta1, double beta2, double eps, double weight_decay,
                        int64_t* step_count, void* param_bf16, const float* publish_src, float* publish_dst,
                        const tt::GradSlots& gs, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG((publish_src == nullptr) == (publish_dst == nullptr), "adamw_step: publish_src and publish_dst go together");
  TT_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step_count && n >= 0, "adamw_step: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t blocks = tt::ceil_div(n > 0 ? n : 1, 1024);         // 4 parameters per thread and iteration
  if (blocks > 8 * tt::kNumSMs) blocks = 8 * tt::kNumSMs;
  const uintptr_t al = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                       reinterpret_cast<uintptr_t>(exp_avg_sq);
  const int vec_ok = ((al & 15) == 0 && (reinterpret_cast<uintptr_t>(param_bf16) & 7) == 0 && (gs.slot_floats % 4) == 0) ? 1 : 0;
  TT_CUDA(tt::launch_kernel(tt::adamw_kernel, dim3((unsigned)blocks), dim3(256), 0, s, true, param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                            beta2, eps, weight_decay, step_count, (__nv_bfloat16*)param_bf16, vec_ok, publish_src, publish_dst, gs));
  TT_LAUNCH_CHECK("adamw_kernel");
  return TT_OK;
}

extern "C" int tt_adamw_step_publish(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                     double lr, double beta1, double beta2, double eps, double weight_decay,
                                     int64_t* step_count, void* param_bf16, const float* publish_src, float* publish_dst,
                                     void* stream) {
  return adamw_launch(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_count, param_bf16,
                      publish_src, publish_dst, tt::GradSlots{}, stream);
}

extern "C" int tt_adamw_step_p2p(float* param, float* grad_sum, const tt_p2p_t* grad_exchange, float* exp_avg, float* exp_avg_sq,
                                 int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay,
                                 int64_t* step_count, void* param_bf16, const float* publish_src, float* publish_dst,
                                 void* stream) {
  TT_CHECK_ARG(grad_exchange && grad_exchange->double_buffered && grad_exchange->world >= 1 && grad_exchange->world <= 8 &&
               grad_exchange->rank >= 0 && grad_exchange->rank < grad_exchange->world && grad_exchange->base[grad_exchange->rank] &&
               (size_t)n * 4 <= grad_exchange->slot_bytes, "adamw_step_p2p: needs a double-buffered exchange whose slots hold n floats");
  tt::GradSlots gs{};
  gs.base = static_cast<const unsigned char*>(grad_exchange->base[grad_exchange->rank]);
  gs.world = grad_exchange->world; gs.slot_floats = grad_exchange->slot_bytes / 4; gs.sum_out = grad_sum;
  return adamw_launch(param, grad_sum ? grad_sum : param, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_count,
                      param_bf16, publish_src, publish_dst, gs, stream);
}

extern "C" int tt_adamw_step_extra(float* param, float* grad, const float* extra_grad, int64_t extra_offset, int64_t extra_n,
                                   float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1, double beta2, double eps,
                                   double weight_decay, int64_t* step_count, void* param_bf16, const float* publish_src,
                                   float* publish_dst, void* stream) {
  TT_CHECK_ARG(extra_grad && extra_offset >= 0 && extra_n > 0 && extra_offset + extra_n <= n && extra_offset % 4 == 0 && extra_n % 4 == 0 &&
               (reinterpret_cast<uintptr_t>(extra_grad) & 15) == 0,
               "adamw_step_extra: the extra gradient must cover a 4-aligned segment of the flat buffer");
  tt::GradSlots gs{};
  gs.extra = extra_grad; gs.extra_lo = extra_offset; gs.extra_n = extra_n; gs.sum_out = grad;
  return adamw_launch(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_count, param_bf16,
                      publish_src, publish_dst, gs, stream);
}

extern "C" int tt_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             double lr, double beta1, double beta2, double eps, double weight_decay,
                             int64_t* step_count, void* param_bf16, void* stream) {
  return tt_adamw_step_publish(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_count, param_bf16,
                               nullptr, nullptr, stream);
}
