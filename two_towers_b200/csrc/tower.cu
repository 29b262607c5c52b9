// tower.cu -- K3 / K3': the dense part of the towers (TT_PREC_FP32 path) and the row kernels
// shared by both precisions (L2 normalise fwd/bwd, LayerNorm fwd/bwd, dropout mask).
//
// Reference: MeanPoolingTower.feed_forward + F.normalize, twotower/encoders.py:38-42,77;
//            AveragePoolingTower.projection + F.normalize, twotower/encoders.py:100-104,144-150.
#include "common.cuh"
#include "sgemm.cuh"
#include "tensor_core.cuh"

namespace tt {

// ---- row L2 normalise: y = z / max(||z||, 1e-12)  (F.normalize, encoders.py:77) ------------
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ z, int64_t R, int H, float* __restrict__ y,
                  __nv_bfloat16* __restrict__ y_bf16) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* zr = z + row * H;
  float ss = 0.f;
  for (int e = lane; e < H; e += 32) { float v = zr[e]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  for (int e = lane; e < H; e += 32) {
    float o = zr[e] / denom;
    y[row * H + e] = o;
    if (y_bf16) y_bf16[row * H + e] = __float2bfloat16(o);
  }
}

// dz = (dy - y (y.dy)) / max(||z||, eps);  below eps the clamp makes it dy / eps
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, int64_t R, int H,
                  float* __restrict__ dz) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* zr = z + row * H;
  const float* gr = dy + row * H;
  float ss = 0.f, dot = 0.f;
  for (int e = lane; e < H; e += 32) { float v = zr[e]; ss = fmaf(v, v, ss); dot = fmaf(v, gr[e], dot); }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float n = sqrtf(ss);
  const float denom = fmaxf(n, 1e-12f);
  const float inner = (n > 1e-12f) ? dot / (denom * denom) : 0.f;   // (y.dy)/denom with y = z/denom
  for (int e = lane; e < H; e += 32) dz[row * H + e] = (gr[e] - zr[e] * inner) / denom;
}

// ---- counter-based dropout mask ---------------------------------------------------------
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t idx) {
  uint64_t x = seed + idx * 0x9E3779B97F4A7C15ull;            // splitmix64
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x = x ^ (x >> 31);
  return (float)(x >> 40) * (1.0f / 16777216.0f);
}

// seed_step (nullable): device counter read at run time (the trainer passes AdamW's step count), so a CUDA-graph replay
// draws a fresh mask every step although `seed` itself is baked into the graph; forward and backward of one step see
// the same value because the optimizer bumps it last.
__device__ __forceinline__ uint64_t mix_seed(uint64_t seed, const int64_t* seed_step) {
  return seed_step ? seed + (uint64_t)(*seed_step + 1) * 0xD1B54A32D192ED03ull : seed;
}

// ---- LayerNorm(H) + normalise, one warp per row (encoders.py:103,150) -----------------------
// a: in/out (dropout applied in place when active); stats[row] = (mean, rstd); z = LN(a); y = z/|z|
__global__ void __launch_bounds__(256)
ln_norm_fwd_kernel(float* __restrict__ a, const float* __restrict__ gamma, const float* __restrict__ beta,
                   int64_t R, int H, float drop_p, int drop_on, uint64_t seed, const int64_t* __restrict__ seed_step,
                   float* __restrict__ stats, float* __restrict__ z, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  seed = mix_seed(seed, seed_step);
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  float* ar = a + row * H;
  if (drop_on) {
    const float keep_scale = 1.0f / (1.0f - drop_p);
    for (int e = lane; e < H; e += 32) {
      float u = uniform01(seed, (uint64_t)row * H + e);
      ar[e] = (u >= drop_p) ? ar[e] * keep_scale : 0.f;
    }
    __syncwarp();
  }
  float s = 0.f;
  for (int e = lane; e < H; e += 32) s += ar[e];
  const float mean = warp_sum(s) / (float)H;
  float v = 0.f;
  for (int e = lane; e < H; e += 32) { float d = ar[e] - mean; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(warp_sum(v) / (float)H + 1e-5f);
  float ss = 0.f;
  for (int e = lane; e < H; e += 32) {
    float o = (ar[e] - mean) * rstd * gamma[e] + beta[e];
    z[row * H + e] = o;
    ss = fmaf(o, o, ss);
  }
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  for (int e = lane; e < H; e += 32) y[row * H + e] = z[row * H + e] / denom;
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// normalise-bwd + LN-bwd.  Writes da (grad wrt the Linear output, dropout mask applied),
// t_gamma = dz * xhat and t_beta = dz (column-summed afterwards for dgamma / dbeta).
__global__ void __launch_bounds__(256)
ln_norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ a, const float* __restrict__ stats,
                   const float* __restrict__ z, const float* __restrict__ gamma, int64_t R, int H,
                   float drop_p, int drop_on, uint64_t seed, const int64_t* __restrict__ seed_step,
                   float* __restrict__ da, float* __restrict__ t_gamma, float* __restrict__ t_beta) {
  const int lane = threadIdx.x & 31;
  seed = mix_seed(seed, seed_step);
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* zr = z + row * H;
  const float* gr = dy + row * H;
  const float* ar = a + row * H;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  float ss = 0.f, dot = 0.f;
  for (int e = lane; e < H; e += 32) { float v = zr[e]; ss = fmaf(v, v, ss); dot = fmaf(v, gr[e], dot); }
  ss = warp_sum(ss); dot = warp_sum(dot);
  const float n = sqrtf(ss), denom = fmaxf(n, 1e-12f);
  const float inner = (n > 1e-12f) ? dot / (denom * denom) : 0.f;
  float m1 = 0.f, m2 = 0.f;                           // mean(dxhat), mean(dxhat*xhat)
  for (int e = lane; e < H; e += 32) {
    float dz = (gr[e] - zr[e] * inner) / denom;
    float xhat = (ar[e] - mean) * rstd;
    float dxh = dz * gamma[e];
    t_gamma[row * H + e] = dz * xhat;
    t_beta[row * H + e] = dz;
    m1 += dxh; m2 = fmaf(dxh, xhat, m2);
  }
  m1 = warp_sum(m1) / (float)H; m2 = warp_sum(m2) / (float)H;
  const float keep_scale = drop_on ? 1.0f / (1.0f - drop_p) : 1.0f;
  for (int e = lane; e < H; e += 32) {
    float xhat = (ar[e] - mean) * rstd;
    float dxh = t_beta[row * H + e] * gamma[e];
    float g = (dxh - m1 - xhat * m2) * rstd;
    if (drop_on) g = (uniform01(seed, (uint64_t)row * H + e) >= drop_p) ? g * keep_scale : 0.f;
    da[row * H + e] = g;
  }
}

static inline unsigned row_grid(int64_t R) { return (unsigned)ceil_div(R, 8); }

// ---------------------------------------------------------------------------------------
// fp32 MLP
// ---------------------------------------------------------------------------------------
struct MlpPlan {
  int s_fwd1, s_fwd2, s_dw2, s_da1, s_dw1, s_dx;
  size_t partial_bytes, colsum_bytes, act_bytes, total;
};
static MlpPlan plan_mlp(int64_t R, int E, int H) {
  MlpPlan p{};
  p.s_fwd1 = sgemm_pick_splits((int)R, H, E);
  p.s_fwd2 = sgemm_pick_splits((int)R, H, H);
  p.s_dw2 = sgemm_pick_splits(H, H, (int)R);
  p.s_da1 = sgemm_pick_splits((int)R, H, H);
  p.s_dw1 = sgemm_pick_splits(H, E, (int)R);
  p.s_dx = sgemm_pick_splits((int)R, E, H);
  size_t pb = 0;
  auto upd = [&](int M, int N, int s) { size_t b = sgemm_partial_bytes(M, N, s); if (b > pb) pb = b; };
  upd((int)R, H, p.s_fwd1); upd((int)R, H, p.s_fwd2); upd(H, H, p.s_dw2); upd((int)R, H, p.s_da1);
  upd(H, E, p.s_dw1); upd((int)R, E, p.s_dx);
  p.partial_bytes = align_up(pb);
  p.colsum_bytes = align_up((size_t)colsum_partial_rows(R) * H * sizeof(float));
  p.act_bytes = align_up((size_t)R * H * sizeof(float));
  p.total = p.partial_bytes + p.colsum_bytes + 2 * p.act_bytes + 256;
  return p;
}

static int mlp_fwd_fp32(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                        int64_t R, int E, int H, float* h1, float* z, float* y, __nv_bfloat16* y_bf16,
                        void* ws, size_t ws_bytes, cudaStream_t s) {
  const MlpPlan plan = plan_mlp(R, E, H);
  if (ws_bytes < plan.total) { set_error("mlp_fwd: workspace too small (%zu < %zu)", ws_bytes, plan.total); return TT_ERR_WORKSPACE; }
  Workspace w(ws, ws_bytes);
  float* partial = w.take<float>(plan.partial_bytes / sizeof(float));
  SgemmArgs g{};
  g.M = (int)R; g.N = H; g.K = E; g.A = x; g.lda = E; g.transA = 0; g.B = w1; g.ldb = E; g.transB = 1;
  g.C = h1; g.ldc = H; g.bias = b1; g.act = 1; g.splits = plan.s_fwd1; g.partial = partial;
  int rc = sgemm(g, s); if (rc) return rc;
  g = SgemmArgs{};
  g.M = (int)R; g.N = H; g.K = H; g.A = h1; g.lda = H; g.transA = 0; g.B = w2; g.ldb = H; g.transB = 1;
  g.C = z; g.ldc = H; g.bias = b2; g.splits = plan.s_fwd2; g.partial = partial;
  rc = sgemm(g, s); if (rc) return rc;
  l2norm_fwd_kernel<<<row_grid(R), 256, 0, s>>>(z, R, H, y, y_bf16);
  TT_LAUNCH_CHECK("l2norm_fwd_kernel");
  return TT_OK;
}

static int mlp_bwd_fp32(const float* dy, const float* x, const float* w1, const float* w2, const float* h1,
                        const float* z, int64_t R, int E, int H, float* dx, float* dw1, float* db1,
                        float* dw2, float* db2, void* ws, size_t ws_bytes, cudaStream_t s) {
  const MlpPlan plan = plan_mlp(R, E, H);
  if (ws_bytes < plan.total) { set_error("mlp_bwd: workspace too small (%zu < %zu)", ws_bytes, plan.total); return TT_ERR_WORKSPACE; }
  Workspace w(ws, ws_bytes);
  float* partial = w.take<float>(plan.partial_bytes / sizeof(float));
  float* cpart = w.take<float>(plan.colsum_bytes / sizeof(float));
  float* dz = w.take<float>((size_t)R * H);
  float* da1 = w.take<float>((size_t)R * H);
  l2norm_bwd_kernel<<<row_grid(R), 256, 0, s>>>(dy, z, R, H, dz);
  TT_LAUNCH_CHECK("l2norm_bwd_kernel");
  int rc;
  SgemmArgs g{};
  // dw2[H,H] = dz^T h1
  g.M = H; g.N = H; g.K = (int)R; g.A = dz; g.lda = H; g.transA = 1; g.B = h1; g.ldb = H; g.transB = 0;
  g.C = dw2; g.ldc = H; g.splits = plan.s_dw2; g.partial = partial;
  rc = sgemm(g, s); if (rc) return rc;
  rc = colsum(dz, R, H, H, db2, cpart, s); if (rc) return rc;
  // da1[R,H] = (dz w2) * (h1 > 0)
  g = SgemmArgs{};
  g.M = (int)R; g.N = H; g.K = H; g.A = dz; g.lda = H; g.transA = 0; g.B = w2; g.ldb = H; g.transB = 0;
  g.C = da1; g.ldc = H; g.mask = h1; g.ldmask = H; g.splits = plan.s_da1; g.partial = partial;
  rc = sgemm(g, s); if (rc) return rc;
  // dw1[H,E] = da1^T x
  g = SgemmArgs{};
  g.M = H; g.N = E; g.K = (int)R; g.A = da1; g.lda = H; g.transA = 1; g.B = x; g.ldb = E; g.transB = 0;
  g.C = dw1; g.ldc = E; g.splits = plan.s_dw1; g.partial = partial;
  rc = sgemm(g, s); if (rc) return rc;
  rc = colsum(da1, R, H, H, db1, cpart, s); if (rc) return rc;
  if (dx) {
    g = SgemmArgs{};
    g.M = (int)R; g.N = E; g.K = H; g.A = da1; g.lda = H; g.transA = 0; g.B = w1; g.ldb = E; g.transB = 0;
    g.C = dx; g.ldc = E; g.splits = plan.s_dx; g.partial = partial;
    rc = sgemm(g, s); if (rc) return rc;
  }
  return TT_OK;
}

// ---------------------------------------------------------------------------------------
// avg_pool projection
// ---------------------------------------------------------------------------------------
struct ProjPlan { int s_fwd, s_dw, s_dx; size_t partial_bytes, colsum_bytes, act_bytes, total; };
static ProjPlan plan_proj(int64_t R, int E, int H) {
  ProjPlan p{};
  p.s_fwd = sgemm_pick_splits((int)R, H, E);
  p.s_dw = sgemm_pick_splits(H, E, (int)R);
  p.s_dx = sgemm_pick_splits((int)R, E, H);
  size_t pb = sgemm_partial_bytes((int)R, H, p.s_fwd);
  size_t b2 = sgemm_partial_bytes(H, E, p.s_dw); if (b2 > pb) pb = b2;
  b2 = sgemm_partial_bytes((int)R, E, p.s_dx); if (b2 > pb) pb = b2;
  p.partial_bytes = align_up(pb);
  p.colsum_bytes = align_up((size_t)colsum_partial_rows(R) * H * sizeof(float));
  p.act_bytes = align_up((size_t)R * H * sizeof(float));
  p.total = p.partial_bytes + p.colsum_bytes + 3 * p.act_bytes + 256;
  return p;
}

}  // namespace tt

extern "C" {

size_t tt_mlp_workspace(int64_t R, int E, int H, int precision) {
  if (R <= 0 || E <= 0 || H <= 0) return 256;
  size_t fp32 = tt::plan_mlp(R, E, H).total;
  if (precision == TT_PREC_BF16) {
    size_t tc = tt::tc_mlp_workspace(R, E, H);
    return tc > fp32 ? tc : fp32;
  }
  return fp32;
}

int tt_mlp_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
               int64_t R, int E, int H, float* h1, float* z, float* y, void* y_bf16, const void* x_bf16,
               const void* w1_bf16, const void* w2_bf16, void* h1_bf16, float* inv_norm,
               const tt_mlp_embed_t* embed, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG((x || embed) && w1 && b1 && w2 && b2 && h1 && R >= 0 && E > 0 && H > 0, "mlp_fwd: bad arguments");
  TT_CHECK_ARG(embed == nullptr || (precision == TT_PREC_BF16 && embed->pool_bf16 && embed->table_bf16 && embed->V > 0),
               "mlp_fwd: embed needs TT_PREC_BF16, pool_bf16 and table_bf16");
  TT_CHECK_ARG(z || (precision == TT_PREC_BF16 && y_bf16 && inv_norm), "mlp_fwd: z may be null only in TT_PREC_BF16 with y_bf16 and inv_norm given");
  TT_CHECK_ARG(inv_norm == nullptr || precision == TT_PREC_BF16, "mlp_fwd: inv_norm is a TT_PREC_BF16 output");
  TT_CHECK_ARG(y || (precision == TT_PREC_BF16 && y_bf16), "mlp_fwd: y may be null only in TT_PREC_BF16 with y_bf16 given");
  TT_CHECK_ARG(R < (1ll << 31), "mlp_fwd: R too large");
  if (R == 0) return TT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == TT_PREC_BF16)
    return tt::tc_mlp_fwd(x, w1, b1, w2, b2, R, E, H, h1, z, y, (__nv_bfloat16*)y_bf16, (const __nv_bfloat16*)x_bf16,
                          (const __nv_bfloat16*)w1_bf16, (const __nv_bfloat16*)w2_bf16, (__nv_bfloat16*)h1_bf16, inv_norm, embed,
                          workspace, workspace_bytes, s);
  TT_CHECK_ARG(precision == TT_PREC_FP32, "mlp_fwd: unknown precision %d", precision);
  return tt::mlp_fwd_fp32(x, w1, b1, w2, b2, R, E, H, h1, z, y, (__nv_bfloat16*)y_bf16, workspace, workspace_bytes, s);
}

int tt_mlp_bwd(const float* dy, const float* x, const float* w1, const float* w2, const float* h1,
               const float* z, int64_t R, int E, int H, float* dx, float* dw1, float* db1, float* dw2,
               float* db2, const void* x_bf16, const void* w1_bf16, const void* w2_bf16, const void* h1_bf16,
               int dy_parts, int64_t dy_part_stride, const tt_mlp_embed_t* embed, const void* y_bf16,
               const float* inv_norm, const void* dz_bf16, const float* dz_colsum, int precision, void* workspace,
               size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(x && w1 && w2 && h1 && dw1 && db1 && dw2 && db2 && R > 0 && E > 0 && H > 0, "mlp_bwd: bad arguments");
  const bool have_dz = dz_bf16 != nullptr;
  TT_CHECK_ARG(!have_dz || (precision == TT_PREC_BF16 && dz_colsum && H <= 512), "mlp_bwd: dz_bf16 needs TT_PREC_BF16, dz_colsum and H <= 512");
  TT_CHECK_ARG(have_dz || dy, "mlp_bwd: dy missing");
  TT_CHECK_ARG(have_dz || z || (precision == TT_PREC_BF16 && y_bf16 && inv_norm), "mlp_bwd: z may be null only in TT_PREC_BF16 with y_bf16 and inv_norm given");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == TT_PREC_BF16) {
    if (embed) {
      TT_CHECK_ARG(embed->pool_bf16 && embed->table && embed->d_table && embed->V > 0 && dx == nullptr,
                   "mlp_bwd: embed needs pool_bf16, table, d_table and dx == NULL");
    }
    return tt::tc_mlp_bwd(dy, x, w1, w2, h1, z, R, E, H, dx, dw1, db1, dw2, db2, (const __nv_bfloat16*)x_bf16,
                          (const __nv_bfloat16*)w1_bf16, (const __nv_bfloat16*)w2_bf16, (const __nv_bfloat16*)h1_bf16,
                          dy_parts, dy_part_stride, embed, (const __nv_bfloat16*)y_bf16, inv_norm, (const __nv_bfloat16*)dz_bf16,
                          dz_colsum, workspace, workspace_bytes, s);
  }
  TT_CHECK_ARG(embed == nullptr, "mlp_bwd: embed is a TT_PREC_BF16 feature");
  TT_CHECK_ARG(precision == TT_PREC_FP32, "mlp_bwd: unknown precision %d", precision);
  TT_CHECK_ARG(dy_parts <= 1, "mlp_bwd: split dy slices are a TT_PREC_BF16 feature");
  return tt::mlp_bwd_fp32(dy, x, w1, w2, h1, z, R, E, H, dx, dw1, db1, dw2, db2, workspace, workspace_bytes, s);
}

int tt_mlp_fwd_embed_ok(int E, int H, int64_t V) { return (E > 0 && H > 0 && V > 0 && tt::tc_mlp_fwd_pool_supported(E, H, V)) ? 1 : 0; }

size_t tt_mlp_embed_workspace(int64_t V, int H, int64_t R) {
  if (V <= 0 || H <= 0 || R <= 0) return 256;
  return tt::tc_mlp_embed_workspace(V, H, R);
}

size_t tt_proj_ln_workspace(int64_t R, int E, int H, int precision) {
  if (R <= 0 || E <= 0 || H <= 0) return 256;
  size_t n = tt::plan_proj(R, E, H).total;
  if (precision == TT_PREC_BF16) n += tt::align_up(tt::tc_proj_workspace(R, E, H));
  return n;
}

int tt_proj_ln_fwd(const float* x, const float* w, const float* b, const float* gamma, const float* beta,
                   int64_t R, int E, int H, int has_projection, float dropout_p, int training, uint64_t seed,
                   const int64_t* seed_step, float* a, float* stats, float* z, float* y, int precision,
                   void* workspace, size_t workspace_bytes, void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(x && y && R >= 0 && E > 0 && H > 0, "proj_ln_fwd: bad arguments");
  TT_CHECK_ARG(precision == TT_PREC_FP32 || precision == TT_PREC_BF16, "proj_ln_fwd: unknown precision %d", precision);
  TT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "proj_ln_fwd: dropout_p must be in [0,1)");
  if (R == 0) return TT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!has_projection) {
    TT_CHECK_ARG(E == H, "proj_ln_fwd: no projection requires E == H");
    tt::l2norm_fwd_kernel<<<tt::row_grid(R), 256, 0, s>>>(x, R, H, y, nullptr);
    TT_LAUNCH_CHECK("l2norm_fwd_kernel");
    return TT_OK;
  }
  TT_CHECK_ARG(w && b && gamma && beta && a && stats && z, "proj_ln_fwd: projection buffers missing");
  const tt::ProjPlan plan = tt::plan_proj(R, E, H);
  if (workspace_bytes < plan.total) { tt::set_error("proj_ln_fwd: workspace too small"); return TT_ERR_WORKSPACE; }
  tt::Workspace wsp(workspace, workspace_bytes);
  float* partial = wsp.take<float>(plan.partial_bytes / sizeof(float));
  int rc;
  if (precision == TT_PREC_BF16) {                          // Linear(E,H) on the tcgen05 tensor cores (bf16 operands, fp32 accumulate)
    if (workspace_bytes < tt_proj_ln_workspace(R, E, H, precision)) { tt::set_error("proj_ln_fwd: workspace too small"); return TT_ERR_WORKSPACE; }
    rc = tt::tc_proj_fwd(x, w, b, R, E, H, a, static_cast<char*>(workspace) + plan.total, workspace_bytes - plan.total, s);
    if (rc) return rc;
  } else {
    tt::SgemmArgs g{};
    g.M = (int)R; g.N = H; g.K = E; g.A = x; g.lda = E; g.transA = 0; g.B = w; g.ldb = E; g.transB = 1;
    g.C = a; g.ldc = H; g.bias = b; g.splits = plan.s_fwd; g.partial = partial;
    rc = tt::sgemm(g, s); if (rc) return rc;
  }
  const int drop_on = (training && dropout_p > 0.f) ? 1 : 0;
  tt::ln_norm_fwd_kernel<<<tt::row_grid(R), 256, 0, s>>>(a, gamma, beta, R, H, dropout_p, drop_on, seed, seed_step, stats, z, y);
  TT_LAUNCH_CHECK("ln_norm_fwd_kernel");
  return TT_OK;
}

int tt_proj_ln_bwd(const float* dy, const float* x, const float* w, const float* gamma, const float* a,
                   const float* stats, const float* z, int64_t R, int E, int H, int has_projection,
                   float dropout_p, int training, uint64_t seed, const int64_t* seed_step, float* dx, float* dw,
                   float* db, float* dgamma, float* dbeta, int precision, void* workspace, size_t workspace_bytes,
                   void* stream) {
  TT_REQUIRE_DEVICE();
  TT_CHECK_ARG(dy && x && R > 0 && E > 0 && H > 0, "proj_ln_bwd: bad arguments");
  TT_CHECK_ARG(precision == TT_PREC_FP32 || precision == TT_PREC_BF16, "proj_ln_bwd: unknown precision %d", precision);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!has_projection) {
    TT_CHECK_ARG(E == H && dx, "proj_ln_bwd: no projection requires E == H and dx");
    tt::l2norm_bwd_kernel<<<tt::row_grid(R), 256, 0, s>>>(dy, x, R, H, dx);
    TT_LAUNCH_CHECK("l2norm_bwd_kernel");
    return TT_OK;
  }
  TT_CHECK_ARG(w && gamma && a && stats && z && dw && db && dgamma && dbeta, "proj_ln_bwd: buffers missing");
  const tt::ProjPlan plan = tt::plan_proj(R, E, H);
  if (workspace_bytes < plan.total) { tt::set_error("proj_ln_bwd: workspace too small"); return TT_ERR_WORKSPACE; }
  tt::Workspace wsp(workspace, workspace_bytes);
  float* partial = wsp.take<float>(plan.partial_bytes / sizeof(float));
  float* cpart = wsp.take<float>(plan.colsum_bytes / sizeof(float));
  float* da = wsp.take<float>((size_t)R * H);
  float* tg = wsp.take<float>((size_t)R * H);
  float* tb = wsp.take<float>((size_t)R * H);
  const int drop_on = (training && dropout_p > 0.f) ? 1 : 0;
  tt::ln_norm_bwd_kernel<<<tt::row_grid(R), 256, 0, s>>>(dy, a, stats, z, gamma, R, H, dropout_p, drop_on, seed, seed_step, da, tg, tb);
  TT_LAUNCH_CHECK("ln_norm_bwd_kernel");
  int rc = tt::colsum(tg, R, H, H, dgamma, cpart, s); if (rc) return rc;
  rc = tt::colsum(tb, R, H, H, dbeta, cpart, s); if (rc) return rc;
  if (precision == TT_PREC_BF16) {                          // dw = da^T x and dx = da w on the tcgen05 tensor cores
    if (workspace_bytes < tt_proj_ln_workspace(R, E, H, precision)) { tt::set_error("proj_ln_bwd: workspace too small"); return TT_ERR_WORKSPACE; }
    rc = tt::colsum(da, R, H, H, db, cpart, s); if (rc) return rc;
    return tt::tc_proj_bwd(da, x, w, R, E, H, dx, dw, static_cast<char*>(workspace) + plan.total, workspace_bytes - plan.total, s);
  }
  tt::SgemmArgs g{};
  g.M = H; g.N = E; g.K = (int)R; g.A = da; g.lda = H; g.transA = 1; g.B = x; g.ldb = E; g.transB = 0;
  g.C = dw; g.ldc = E; g.splits = plan.s_dw; g.partial = partial;
  rc = tt::sgemm(g, s); if (rc) return rc;
  rc = tt::colsum(da, R, H, H, db, cpart, s); if (rc) return rc;
  if (dx) {
    g = tt::SgemmArgs{};
    g.M = (int)R; g.N = E; g.K = H; g.A = da; g.lda = H; g.transA = 0; g.B = w; g.ldb = E; g.transB = 0;
    g.C = dx; g.ldc = E; g.splits = plan.s_dx; g.partial = partial;
    rc = tt::sgemm(g, s); if (rc) return rc;
  }
  return TT_OK;
}

}  // extern "C"
