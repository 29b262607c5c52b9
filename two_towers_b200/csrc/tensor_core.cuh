// tensor_core.cuh -- entry points of the TT_PREC_BF16 path (tcgen05 / TMEM / TMA kernels,
// implemented in tc_*.cu).  Called by the C-ABI functions in tower.cu / inbatch_ce.cu.
#pragma once
#include "common.cuh"

namespace tt {

// tower MLP on tcgen05 (tc_mlp.cu)
size_t tc_mlp_workspace(int64_t R, int E, int H);
int tc_mlp_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
               int64_t R, int E, int H, float* h1, float* z, float* y, __nv_bfloat16* y_bf16,
               const __nv_bfloat16* x_bf16, const __nv_bfloat16* w1_bf16, const __nv_bfloat16* w2_bf16,
               __nv_bfloat16* h1_bf16, float* inv_norm, const tt_mlp_embed_t* embed, void* ws, size_t ws_bytes, cudaStream_t s);
int tc_mlp_bwd(const float* dy, const float* x, const float* w1, const float* w2, const float* h1,
               const float* z, int64_t R, int E, int H, float* dx, float* dw1, float* db1, float* dw2,
               float* db2, const __nv_bfloat16* x_bf16, const __nv_bfloat16* w1_bf16,
               const __nv_bfloat16* w2_bf16, const __nv_bfloat16* h1_bf16, int dy_parts, int64_t dy_part_stride,
               const tt_mlp_embed_t* embed, const __nv_bfloat16* y_bf16, const float* inv_norm, const __nv_bfloat16* dz_bf16,
               const float* dz_colsum, void* ws, size_t ws_bytes, cudaStream_t s);
size_t tc_mlp_embed_workspace(int64_t V, int H, int64_t R);

// avg_pool projection Linear(E,H) on tcgen05 (tc_gemm.cu): forward a = x w^T + b, backward dw = da^T x, dx = da w
size_t tc_proj_workspace(int64_t R, int E, int H);
int tc_proj_fwd(const float* x, const float* w, const float* b, int64_t R, int E, int H, float* a, void* ws, size_t ws_bytes,
                cudaStream_t s);
int tc_proj_bwd(const float* da, const float* x, const float* w, int64_t R, int E, int H, float* dx, float* dw, void* ws,
                size_t ws_bytes, cudaStream_t s);

// whole-MLP forward in one kernel (tc_mlp.cu)
bool tc_mlp_fused_supported(int E, int H);
int tc_mlp_fwd_fused(const __nv_bfloat16* xb, const __nv_bfloat16* w1b, const float* b1, const __nv_bfloat16* w2b,
                     const float* b2, int64_t R, int E, int H, __nv_bfloat16* h1b, float* z, float* y,
                     __nv_bfloat16* yb, float* inv_norm, const __nv_bfloat16* pool, int64_t V, const __nv_bfloat16* table_bf16,
                     cudaStream_t s, const void* ids = nullptr, int id_bytes = 0, int L = 0, float* inv_len = nullptr);
bool tc_mlp_fwd_pool_supported(int E, int H, int64_t V);

// fused similarity GEMM + online-LSE cross entropy on tcgen05 (tc_inbatch.cu)
size_t tc_inbatch_workspace(int64_t Bq, int64_t Bd, int H);
int tc_inbatch_fwd(const float* q, const float* d, const __nv_bfloat16* q_bf16, const __nv_bfloat16* d_bf16,
                   int64_t Bq, int64_t Bd, int H, float inv_temp, int64_t label_offset, float loss_scale,
                   float* loss, float* lse, float* pos_mean, void* ws, size_t ws_bytes, cudaStream_t s);
int tc_inbatch_bwd(const float* q, const float* d, const __nv_bfloat16* q_bf16, const __nv_bfloat16* d_bf16,
                   const float* lse, int64_t Bq, int64_t Bd, int H, float inv_temp, int64_t label_offset,
                   float loss_scale, const float* grad_out, float* dq, float* dd, void* ws, size_t ws_bytes,
                   cudaStream_t s);

// partial-slice backward for the fused trainer: both gradients in one launch, slices summed by the consumer
int tc_inbatch_bwd_nparts(int64_t Bq, int64_t Bd, int H);
int tc_inbatch_bwd_parts(const __nv_bfloat16* q_bf16, const __nv_bfloat16* d_bf16, const float* lse, int64_t Bq, int64_t Bd,
                         int H, float inv_temp, int64_t label_offset, float loss_scale, const float* grad_out,
                         float* dq_parts, int64_t stride_q, float* dd_parts, int64_t stride_d, cudaStream_t s);

int tc_inbatch_bwd_nparts2(int64_t Bx0, int64_t By0, int64_t Bx1, int64_t By1, int H);
size_t tc_inbatch_onepass_sync_bytes(int64_t Bq);
int tc_inbatch_onepass_ok(int64_t Bq, int64_t Bd, int H, float logit_bound);
int tc_inbatch_dd_nparts(int64_t x_rows, int64_t y_rows);
int tc_inbatch_fwd_dq(const tt_ce_pass_t* q_pass, int H, float inv_temp, float logit_bound, float loss_scale, const float* grad_out,
                      float* loss, float* lse_out, float* pos_mean, void* sync_scratch, const tt_p2p_t* y_exchange, const void* y_own, cudaStream_t s, void* stash = nullptr);
int tc_inbatch_onepass_single(const tt_ce_pass_t* q_pass, const tt_ce_pass_t* d_pass, int H, float inv_temp, float logit_bound, float loss_scale,
                              const float* grad_out, float* loss, float* lse_out, float* pos_mean, void* sync_scratch, cudaStream_t s);
size_t tc_inbatch_stash_bytes(int64_t Bq, int64_t Bd, int H);
int tc_inbatch_stash_ok(int64_t Bq, int64_t Bd, int H);
int tc_inbatch_dd_stored(const tt_ce_pass_t* d_pass, int H, float inv_temp, float loss_scale, const float* grad_out, const void* stash,
                         cudaStream_t s);
int tc_inbatch_dd(const tt_ce_pass_t* d_pass, int H, float inv_temp, float loss_scale, const float* grad_out, cudaStream_t s);
int tc_inbatch_bwd_parts_ex(const tt_ce_pass_t* q_pass, const tt_ce_pass_t* d_pass, int H, float inv_temp, float loss_scale,
                            const float* grad_out, int nparts, cudaStream_t s);
size_t tc_inbatch_fwd_ex_workspace(int64_t Bq, int64_t Bd);
int tc_inbatch_fwd_ex(const __nv_bfloat16* qa, int64_t Bq, const __nv_bfloat16* da, int64_t Bd, int64_t d_buf_rows,
                      int64_t d_blk, int64_t d_blk_stride, int64_t d_blk_off, int H, float inv_temp, int64_t label_offset,
                      float loss_scale, float* loss, float* lse, float* pos_mean, float* part_ml, float* pos, void* sync_scratch,
                      cudaStream_t s);
size_t tc_inbatch_fwd_sync_bytes(int64_t Bq);

// shared by both precisions (inbatch_ce.cu): lse/loss finalisation from per-split (max,sum)
int inbatch_finalize(const float* part_ml, const float* pos_logit, int nsplit, int64_t Bq, float inv_temp,
                     float loss_scale, float* lse, float* loss, float* pos_mean, float* scratch,
                     cudaStream_t s);

}  // namespace tt
