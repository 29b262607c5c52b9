/* tt_b200.h -- C ABI of libtt_b200.so: the B200 (sm_100a) two-tower hot path.
 *
 * This is the drop-in boundary.  The reference (k0r1g/two-towers) is pure Python/PyTorch;
 * its "FFI" for this path is the set of ATen calls made by its registered classes.  Each
 * entry point below names the reference call site (file:line under /root/reference) it
 * replaces.  The host-side mirror of the reference's registries (two_towers_b200/*.py) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  All data pointers are DEVICE pointers
 *     on the current CUDA device unless the name ends in _host.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); nothing
 *     allocates, synchronises or spawns threads.  Workspaces are caller-allocated; query
 *     their size with the matching *_workspace() function (bytes, 256-B aligned base needed).
 *   - return value: TT_OK (0) or a negative TT_ERR_*; tt_last_error() gives the message for
 *     the calling thread.  There is NO CPU fallback: without an sm_100 device every compute
 *     entry point returns TT_ERR_ARCH.
 *   - matrices are row-major and contiguous; `ids` are token ids of `id_bytes` in {4, 8}
 *     (int32 / int64; the reference uses int64, dataset.py:274-277).
 *   - precision: TT_PREC_FP32 = CUDA-core fp32 FFMA (parity mode, rel 1e-5);
 *                TT_PREC_BF16 = bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM
 *                (performance mode, rel 2e-2).
 *   - all reductions use a fixed order: results are bitwise reproducible run to run.
 */
#ifndef TT_B200_H
#define TT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_ABI_VERSION 6

enum { TT_OK = 0, TT_ERR_INVALID = -1, TT_ERR_CUDA = -2, TT_ERR_ARCH = -3, TT_ERR_WORKSPACE = -4,
       TT_ERR_UNSUPPORTED = -5 };
enum { TT_PREC_FP32 = 0, TT_PREC_BF16 = 1 };

/* ---- library -------------------------------------------------------------------------- */
int tt_abi_version(void);
const char* tt_last_error(void);
/* TT_OK iff `device` exists and is compute capability 10.x (B200).  Cached per device. */
int tt_require_sm100(int device);
/* Number of kernel launches issued by this library in the calling process so far. */
int64_t tt_launch_count(void);
/* Token ids outside [0, V): the reference's nn.Embedding raises IndexError (twotower/embeddings.py:33-40).  The gather
 * kernels treat such a token as padding (no out-of-bounds read) and record it in a mapped host word; this call returns 1
 * and the last offending id if any completed kernel has seen one (no synchronisation here: call it after the stream /
 * event wait that made the results visible), optionally clearing the record.  The Python layer turns it into IndexError. */
int tt_bad_token_id(int64_t* id, int clear);

/* ---- K1: token-id gather + masked mean pool ----------------------------------------------
 * tt_embed_gather     : LookupEmbedding.forward, twotower/embeddings.py:33-40 (W[ids]).
 * tt_embed_pool_fwd   : MeanPoolingTower.forward twotower/encoders.py:62,67,72 (and :125-138):
 *                       mask=(ids>0); pooled = sum_l W[ids]*mask / (sum mask + 1e-9).
 *                       Never materialises [rows,L,E].  inv_len[r] = 1/(count_r + 1e-9).
 *                       pooled_bf16 (nullable) receives a bf16 copy for the tensor-core path.
 *                       pool_bf16 (nullable, V <= 1024): the pooling matrix P [rows,V] in bf16,
 *                       P[r,v] = count_r(v) / (count_r + 1e-9), so that pooled = P * table; consumed by
 *                       tt_mlp_bwd(embed = ...) which then needs neither dx nor tt_embed_pool_bwd.
 *                       pooled == NULL (with pool_bf16): only P and inv_len are produced -- for tt_mlp_fwd(embed = ...),
 *                       which forms x = P * table on the tensor cores inside the tower kernel.
 * tt_embed_pool_bwd   : autograd of the above -> ATen embedding_dense_backward
 *                       (loss.backward(), twotower/train.py:138).  Deterministic: small tables
 *                       use a pooling-matrix GEMM with fixed split order, large tables a
 *                       stable radix sort of (id,row) + ordered segment reduction.
 *                       d_table [V,E] is OVERWRITTEN (row 0 = 0, untouched rows = 0).
 */
int tt_embed_gather(const void* ids, int id_bytes, const float* table, int64_t n_tokens,
                    int64_t V, int E, float* out, void* stream);
int tt_embed_pool_fwd(const void* ids, int id_bytes, const float* table, int64_t rows, int L,
                      int64_t V, int E, float* pooled, float* inv_len, void* pooled_bf16,
                      void* pool_bf16, void* stream);
size_t tt_embed_pool_bwd_workspace(int64_t rows, int L, int64_t V, int E);
int tt_embed_pool_bwd(const void* ids, int id_bytes, const float* inv_len, const float* d_pooled,
                      int64_t rows, int L, int64_t V, int E, float* d_table,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3: tower MLP  Linear(E,H) -> ReLU -> Linear(H,H) -> L2 normalise ---------------------
 * Replaces MeanPoolingTower.feed_forward + F.normalize, twotower/encoders.py:38-42,77.
 * Weights are nn.Linear layout: w1 [H,E], b1 [H], w2 [H,H], b2 [H].
 * Saved for backward (caller-allocated): h1 [R,H] (post-ReLU), z [R,H] (pre-normalise).
 * y [R,H] has unit rows: y = z / max(||z||, 1e-12).  y_bf16 (nullable): bf16 copy of y.
 * TT_PREC_BF16: y may be null when y_bf16 is given; h1 is opaque saved state (it holds the hidden activation as
 * R*H bf16 values unless a separate h1_bf16 buffer is passed); E % 64 == 0, H % 64 == 0, H <= 256 runs the whole
 * forward as ONE kernel (both weight matrices resident in shared memory).
 * tt_mlp_bwd: given dy, produces dx [R,E] (nullable), dw1, db1, dw2, db2 (OVERWRITTEN).
 * bf16 operand shadows (TT_PREC_BF16 only, all nullable -- missing ones are produced inside the call):
 *   x_bf16 [R,E], w1_bf16 [H,E], w2_bf16 [H,H] (e.g. the shadow tt_adamw_step maintains),
 *   h1_bf16 [R,H] written by fwd and read by bwd.
 * dy_parts > 1 (TT_PREC_BF16): dy is given as dy_parts slices, dy_part_stride elements apart, that are summed
 *   in slice order inside the first backward kernel (see tt_inbatch_ce_bwd_parts); pass 1, 0 otherwise.
 * inv_norm (TT_PREC_BF16, nullable) [R]: tt_mlp_fwd also stores 1 / max(||z||, 1e-12); with y_bf16 AND inv_norm the
 *   normalise step is fully described by them -- fwd may then be given z == NULL (8 bytes/element less to write) and
 *   bwd, given the same y_bf16 / inv_norm, ignores z:  dz = (dy - y (y . dy)) * inv_norm.
 * dz_bf16 + dz_colsum (TT_PREC_BF16, nullable): the normalise backward has already been done by the loss kernel
 *   (tt_ce_pass_t.dz_bf16): dz [R,H] bf16 and its per-32-row column sums [ceil(R/32),H]; dy, z, y_bf16, inv_norm are
 *   then ignored (may be null).
 * embed (TT_PREC_BF16, nullable): x is the mean-pooled lookup embedding x = P * table (P from tt_embed_pool_fwd).
 *   The backward then forms M = P^T da1 [V,H] once on the tensor cores and finishes with two tiny products,
 *   dw1 = M^T table and d_table = M w1 (+= when accumulate != 0), instead of computing dx [R,E], dw1 = da1^T x and
 *   the separate embedding backward; dx must be null.  workspace: tt_mlp_embed_workspace(V, H, R) bytes.
 *   tt_mlp_fwd(embed = ...) (when tt_mlp_fwd_embed_ok): x = P * table_bf16 is formed on the tensor cores inside the
 *   tower kernel (x and x_bf16 are ignored): the pooled activations never reach HBM.
 */

typedef struct tt_mlp_embed_s {
  const void* pool_bf16;      /* P [R,V] bf16 */
  int64_t V;
  const float* table;         /* [V,E] */
  const void* table_bf16;     /* [V,E] bf16 shadow (forward only; e.g. the one tt_adamw_step maintains) */
  float* d_table;             /* [V,E] */
  int accumulate;             /* 0: d_table is overwritten, else added to (second tower sharing the table) */
  void* workspace; size_t workspace_bytes;
  /* tt_mlp_fwd only, optional (ids == NULL: pool_bf16 is an INPUT produced by tt_embed_pool_fwd): the tower kernel builds
   * the pooling matrix itself from the token ids [R,L] (int32: id_bytes 4, int64: 8; 0 = padding; L <= 255, V <= 128 E / 64) -- mask +
   * mean of twotower/encoders.py:62-72 as one integer histogram per row -- and writes it to pool_bf16 for the backward; inv_len
   * (nullable) [R] receives 1 / (non-pad tokens + 1e-9).  The histogram launch of the step disappears. */
  const void* ids; int id_bytes; int L; float* inv_len;
} tt_mlp_embed_t;
size_t tt_mlp_embed_workspace(int64_t V, int H, int64_t R);
size_t tt_mlp_workspace(int64_t R, int E, int H, int precision);
int tt_mlp_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
               int64_t R, int E, int H, float* h1, float* z, float* y, void* y_bf16,
               const void* x_bf16, const void* w1_bf16, const void* w2_bf16, void* h1_bf16,
               float* inv_norm, const tt_mlp_embed_t* embed,
               int precision, void* workspace, size_t workspace_bytes, void* stream);
/* 1 when tt_mlp_fwd(embed = ...) can form x = P * table inside the tower kernel for these shapes. */
int tt_mlp_fwd_embed_ok(int E, int H, int64_t V);
int tt_mlp_bwd(const float* dy, const float* x, const float* w1, const float* w2,
               const float* h1, const float* z, int64_t R, int E, int H,
               float* dx, float* dw1, float* db1, float* dw2, float* db2,
               const void* x_bf16, const void* w1_bf16, const void* w2_bf16, const void* h1_bf16,
               int dy_parts, int64_t dy_part_stride, const tt_mlp_embed_t* embed,
               const void* y_bf16, const float* inv_norm, const void* dz_bf16, const float* dz_colsum,
               int precision, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3': avg_pool tower projection  Linear(E,H) -> Dropout(p) -> LayerNorm(H) -> normalise -
 * Replaces AveragePoolingTower.projection + F.normalize, twotower/encoders.py:100-104,144-150.
 * has_projection == 0 (H == E): y = normalise(x) only (w,b,gamma,beta ignored).
 * Dropout: p == 0 or training == 0 -> identity; otherwise nn.Dropout semantics (kept values scaled by 1/(1-p)) with a
 * counter-based keep-mask: element (row, col) is kept iff u01(splitmix64(seed' + (row*H + col) * 0x9E3779B97F4A7C15)) >= p,
 * recomputed in backward.  seed' = seed, or seed + (*seed_step + 1) * 0xD1B54A32D192ED03 when seed_step (nullable device
 * int64, e.g. tt_adamw_step's counter) is given: a CUDA-graph replay then draws a fresh mask every optimizer step.
 * torch's Philox stream cannot be reproduced; parity is checked with the same mask applied to the oracle.
 * TT_PREC_BF16: the Linear(E,H) forward and its two backward products run on the tcgen05 tensor cores (bf16 operands,
 * fp32 accumulate, any E / H); Dropout, LayerNorm and the normalise stay fp32 row kernels.
 * Saved: a [R,H] (post-dropout pre-LN), stats [R,2] (mean, rstd), z [R,H] (LN output).
 */
size_t tt_proj_ln_workspace(int64_t R, int E, int H, int precision);
int tt_proj_ln_fwd(const float* x, const float* w, const float* b, const float* gamma,
                   const float* beta, int64_t R, int E, int H, int has_projection,
                   float dropout_p, int training, uint64_t seed, const int64_t* seed_step,
                   float* a, float* stats, float* z, float* y,
                   int precision, void* workspace, size_t workspace_bytes, void* stream);
int tt_proj_ln_bwd(const float* dy, const float* x, const float* w, const float* gamma,
                   const float* a, const float* stats, const float* z,
                   int64_t R, int E, int H, int has_projection,
                   float dropout_p, int training, uint64_t seed, const int64_t* seed_step,
                   float* dx, float* dw, float* db, float* dgamma, float* dbeta,
                   int precision, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K4: in-batch sampled-softmax loss, fused similarity GEMM + online logsumexp CE --------
 * Replaces in_batch_sampled_softmax_loss, twotower/losses.py:107-116: S = Q D^T; S/temperature;
 * labels = arange; cross_entropy(mean).  The [Bq,Bd] logits never reach HBM.
 *   q [Bq,H], d [Bd,H]; row i's positive is column i + label_offset (0 == the reference;
 *   rank*B_local with all-gathered D gives global in-batch negatives).
 *   loss (device scalar) = loss_scale * sum_i (lse_i - logit_ii); pass loss_scale = 1/Bq for
 *   the reference's mean.  lse [Bq] is saved for backward.  pos_mean (nullable device scalar)
 *   = mean_i S_ii (the train.py:145-147 monitoring value, free here).
 * tt_inbatch_ce_bwd: dq = g*loss_scale/temp * (P - I) D ; dd = g*loss_scale/temp * (P - I)^T Q,
 *   g read from device scalar grad_out (nullable == 1).  dq / dd nullable.
 * q_bf16/d_bf16 (nullable): bf16 copies of q/d used by TT_PREC_BF16 (else converted inside).
 */
size_t tt_inbatch_ce_workspace(int64_t Bq, int64_t Bd, int H, int precision);
int tt_inbatch_ce_fwd(const float* q, const float* d, const void* q_bf16, const void* d_bf16,
                      int64_t Bq, int64_t Bd, int H, float inv_temperature, int64_t label_offset,
                      float loss_scale, float* loss, float* lse, float* pos_mean,
                      int precision, void* workspace, size_t workspace_bytes, void* stream);
int tt_inbatch_ce_bwd(const float* q, const float* d, const void* q_bf16, const void* d_bf16,
                      const float* lse, int64_t Bq, int64_t Bd, int H, float inv_temperature,
                      int64_t label_offset, float loss_scale, const float* grad_out,
                      float* dq, float* dd,
                      int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Partial-slice backward (TT_PREC_BF16, bf16 operands, H % 64 == 0, H <= 256): ONE launch computes both
 * gradients; the Y range is split over n = tt_inbatch_ce_bwd_nparts() CTAs per row tile and slice s of
 * dq/dd is written at dq_parts + s*dq_part_stride (dd likewise).  The true gradient is the sum of the
 * slices in slice order -- tt_mlp_bwd(dy_parts = n) consumes them directly, so no reduction kernel and no
 * extra pass over [B,H] is needed.  A null dq_parts / dd_parts skips that gradient.
 */
int tt_inbatch_ce_bwd_nparts(int64_t Bq, int64_t Bd, int H, int precision);

/* General forms for data-parallel training with GLOBAL in-batch negatives (bf16 operands).  The "all" operand may be
 * the raw output of an all-gather of per-rank [Q_r | D_r] blocks: logical row g of the gathered matrix lives at
 * physical row (g / blk) * blk_stride + g % blk + blk_off of a buffer with buf_rows rows (blk % 64 == 0), so no
 * re-packing copy is needed between the collective and the kernel.
 *   tt_inbatch_ce_fwd_ex : local queries [Bq,H] vs gathered documents (Bd logical rows); label_offset = rank*Bq.
 *   tt_ce_pass_t         : one gradient pass -- rows X get gradients from all logical rows of Y.
 *       q pass: x = local queries,   y = gathered documents, lse = local row lse [x_rows],  positive col = row + off
 *       d pass: x = local documents, y = gathered queries,   lse = gathered lse [y_rows],   positive at row == col + off
 *   tt_inbatch_ce_bwd_parts_ex : both passes in ONE launch; slice s of each output at out_parts + s*part_stride.
 */
typedef struct {
  const void* x_bf16; int64_t x_rows;
  const void* y_bf16; int64_t y_rows;
  int64_t y_buf_rows, y_blk, y_blk_stride, y_blk_off;
  const float* lse; int64_t label_offset;
  float* out_parts; int64_t part_stride;
  /* Fused normalise backward (all three or none; needs nparts <= 2, see tt_inbatch_ce_bwd_fused_ok): x is the unit-norm
   * output y = z / |z| of tt_mlp_fwd and inv_norm its 1/|z|.  Instead of gradient slices the launch then writes
   *   dz_bf16 [x_rows,H] = (dy - y (y . dy)) * inv_norm  and  dz_colsum [ceil(x_rows/32), H] (per-32-row column sums),
   * exactly what tt_mlp_bwd(dz_bf16 = ..., dz_colsum = ...) consumes; out_parts is ignored. */
  void* dz_bf16; float* dz_colsum; const float* inv_norm;
} tt_ce_pass_t;
size_t tt_inbatch_ce_fwd_ex_workspace(int64_t Bq, int64_t Bd);
int tt_inbatch_ce_fwd_ex(const void* q_bf16, int64_t Bq, const void* d_bf16, int64_t Bd, int64_t d_buf_rows,
                         int64_t d_blk, int64_t d_blk_stride, int64_t d_blk_off, int H, float inv_temperature,
                         int64_t label_offset, float loss_scale, float* loss, float* lse, float* pos_mean,
                         void* workspace, size_t workspace_bytes, void* sync_scratch, void* stream);
/* sync_scratch (nullable): tt_inbatch_ce_sync_bytes(Bq) bytes the caller keeps for the life of the call site, zero-filled
 * once before first use; every call re-arms it for the next one (do not share it between concurrent calls).  With it the loss / lse finalisation runs inside the
 * same launch (the last CTA of each row tile merges its splits, fixed order); without it a second launch does. */
size_t tt_inbatch_ce_sync_bytes(int64_t Bq);
int tt_inbatch_ce_bwd_nparts_ex(int64_t q_x_rows, int64_t q_y_rows, int64_t d_x_rows, int64_t d_y_rows, int H);
/* 1 when tt_inbatch_ce_bwd_parts_ex can run the fused normalise backward for these shapes (both passes in one launch: <= 2 splits). */
int tt_inbatch_ce_bwd_fused_ok(int64_t q_x_rows, int64_t q_y_rows, int64_t d_x_rows, int64_t d_y_rows, int H);
int tt_inbatch_ce_bwd_parts_ex(const tt_ce_pass_t* q_pass, const tt_ce_pass_t* d_pass, int H, float inv_temperature,
                               float loss_scale, const float* grad_out, int nparts, void* stream);
/* One-pass step (TT_PREC_BF16 operands, H % 64 == 0, H <= 256), for logits with a known bound: |logit_ij| <= logit_bound
 * (unit-norm tower outputs, twotower/encoders.py:77: logit_bound = inv_temperature).  With the bound as a fixed softmax
 * shift no running maximum is needed, so the forward of twotower/losses.py:107-116 and the query gradient come out of ONE
 * pass over S = X Y^T:  E = exp(S/temp - bound) feeds the second product O += E Y tile by tile while its row sums L
 * accumulate, and the row is normalised once at the end:  dq_i = g*loss_scale/temp * (O_i / L_i - y_pos(i)),
 * lse_i = bound + log L_i.  S is formed once instead of three times (forward, dq pass, dd pass of the two-launch form).
 *   tt_inbatch_ce_fwd_dq : q_pass as for tt_inbatch_ce_bwd_parts_ex, except that q_pass->lse is ignored (lse is an
 *       OUTPUT here, [x_rows]) and, without dz_bf16, out_parts receives the finished dq [x_rows,H] (no slices).
 *       The 1, 2 or 4 CTAs that share a row tile form a cluster and reduce their accumulators through distributed
 *       shared memory in rank order (bitwise reproducible).  sync_scratch: tt_inbatch_ce_onepass_sync_bytes(x_rows)
 *       bytes, zero-filled once by the caller, re-armed by every call.
 *   tt_inbatch_ce_dd     : the document gradient as its own launch (it needs the lse of every -- gathered -- query):
 *       d_pass as for tt_inbatch_ce_bwd_parts_ex; with out_parts the gradient is left as tt_inbatch_ce_dd_nparts slices.
 *   tt_inbatch_ce_onepass_ok : 1 when the shapes and the bound qualify (2 * logit_bound * log2(e) < 120 keeps E a normal
 *       fp32 number for every admissible logit).
 */
int tt_inbatch_ce_onepass_ok(int64_t Bq, int64_t Bd, int H, float logit_bound);
size_t tt_inbatch_ce_onepass_sync_bytes(int64_t Bq);
int tt_inbatch_ce_fwd_dq(const tt_ce_pass_t* q_pass, int H, float inv_temperature, float logit_bound, float loss_scale,
                         const float* grad_out, float* loss, float* lse, float* pos_mean, void* sync_scratch, void* stream);
/* Both launches as ONE for the square single-process case (x_rows == y_rows on both passes, offsets 0, dz form,
 * d_pass->lse == lse): forward + query gradient, a grid-wide barrier inside the kernel (every CTA of the <= 148-CTA grid is
 * resident; checked with the occupancy calculator), document gradient.  Results are bitwise those of the two calls above.
 * Returns TT_ERR_UNSUPPORTED (-> make the two calls) when the shapes or the device do not allow it. */
int tt_inbatch_ce_onepass(const tt_ce_pass_t* q_pass, const tt_ce_pass_t* d_pass, int H, float inv_temperature, float logit_bound,
                          float loss_scale, const float* grad_out, float* loss, float* lse, float* pos_mean, void* sync_scratch,
                          void* stream);
int tt_inbatch_ce_dd_nparts(int64_t d_x_rows, int64_t d_y_rows, int H);
/* tt_inbatch_ce_fwd_dq fused with the all-gather of the documents (data-parallel training, global in-batch negatives,
 * twotower/losses.py:107-116 over the concatenated batch): q_pass->y_bf16 must be the gathered slots of `y_exchange`
 * (base[rank] + 256: world x y_rows/world x H bf16), and the call must follow tt_p2p_allgather(y_exchange, ...) on the same
 * stream, whose `src` is y_own_bf16 (this rank's y_rows/world rows; x_rows == y_rows/world, label_offset == rank * x_rows).
 * The kernel is launched programmatically behind the exchange kernel and does not wait for it to retire: it starts on this
 * rank's own rows, read where the towers left them, and its TMA warp polls the exchange's per-source arrival counters and
 * consumes each peer's block as it lands -- the NVLink transfer and the wait for the slowest peer run under the loss
 * kernel's main loop. */
struct tt_p2p_s;
int tt_inbatch_ce_fwd_dq_p2p(const tt_ce_pass_t* q_pass, int H, float inv_temperature, float logit_bound, float loss_scale,
                             const float* grad_out, float* loss, float* lse, float* pos_mean, void* sync_scratch,
                             const struct tt_p2p_s* y_exchange, const void* y_own_bf16, void* stream);
int tt_inbatch_ce_dd(const tt_ce_pass_t* d_pass, int H, float inv_temperature, float loss_scale, const float* grad_out,
                     void* stream);
/* Stored-E form of the one-pass step (single process, or per-rank negatives): while the B_q x B_d bf16 matrix
 * E = exp(S/temp - bound) fits in L2 (tt_inbatch_ce_stash_ok: H >= 128 and at most TT_CE_STASH_MAX_MB, default 48, MB),
 * tt_inbatch_ce_fwd_dq_stash also writes every E tile it forms (TMA store of the tile it has just fed to the tensor cores; the
 * positives left out), the rows x_i / L_i and 1 - P_pos(i) into `stash` (tt_inbatch_ce_stash_bytes, caller-owned, no
 * initialisation), and tt_inbatch_ce_dd_stash forms the document gradient of twotower/losses.py:107-116 as ONE plain product
 *   dd_j = g*loss_scale/temp * ( sum_i E_ij x_i / L_i  -  (1 - P_pos(i(j))) x_i(j) )
 * (d_pass: x = documents, y = the queries of the first call, label_offset as there; lse is not used) -- S is formed once per
 * step (6 B^2 H FLOP executed for the 6 B^2 H the loss and its two gradients need) and no exponential is taken twice.  The two
 * calls must use the same stash on the same stream.  Results differ from tt_inbatch_ce_dd only by bf16 rounding of x / L. */
int tt_inbatch_ce_stash_ok(int64_t Bq, int64_t Bd, int H);
size_t tt_inbatch_ce_stash_bytes(int64_t Bq, int64_t Bd, int H);
int tt_inbatch_ce_fwd_dq_stash(const tt_ce_pass_t* q_pass, int H, float inv_temperature, float logit_bound, float loss_scale,
                               const float* grad_out, float* loss, float* lse, float* pos_mean, void* sync_scratch, void* stash,
                               void* stream);
int tt_inbatch_ce_dd_stash(const tt_ce_pass_t* d_pass, int H, float inv_temperature, float loss_scale, const float* grad_out,
                           const void* stash, void* stream);
int tt_inbatch_ce_bwd_parts(const void* q_bf16, const void* d_bf16, const float* lse, int64_t Bq, int64_t Bd,
                            int H, float inv_temperature, int64_t label_offset, float loss_scale,
                            const float* grad_out, float* dq_parts, int64_t dq_part_stride,
                            float* dd_parts, int64_t dd_part_stride, void* stream);

/* ---- K5/K6: row-paired cosine losses ------------------------------------------------------
 * tt_triplet_*   : contrastive_triplet_loss, twotower/losses.py:28-35:
 *                  mean(relu(margin - cos(q,p) + cos(q,n))), cosine eps 1e-8.
 *                  sims: 3*B floats saved -- [B,2] (cos_qp, cos_qn) followed by [B] row
 *                  losses; pos_mean/neg_mean nullable device scalars (train.py:145-151
 *                  monitoring).
 * tt_multineg_*  : multiple_negatives_loss, twotower/losses.py:65-83: cos(q,[p;negs])/temp,
 *                  CE vs label 0.  negs [B,N,H], N <= 63; probs: B*(N+1)+B floats saved --
 *                  [B,N+1] softmax followed by [B] row losses.
 */
int tt_triplet_fwd(const float* q, const float* p, const float* n, int64_t B, int H, float margin,
                   float* loss, float* sims, float* pos_mean, float* neg_mean, void* stream);
int tt_triplet_bwd(const float* q, const float* p, const float* n, const float* sims,
                   int64_t B, int H, float margin, const float* grad_out,
                   float* dq, float* dp, float* dn, void* stream);
int tt_multineg_fwd(const float* q, const float* p, const float* negs, int64_t B, int N, int H,
                    float inv_temperature, float* loss, float* probs, void* stream);
int tt_multineg_bwd(const float* q, const float* p, const float* negs, const float* probs,
                    int64_t B, int N, int H, float inv_temperature, const float* grad_out,
                    float* dq, float* dp, float* dnegs, void* stream);

/* ---- K7: brute-force scan + top-k ---------------------------------------------------------
 * Replaces TwoTowerSearch.search scoring + torch.topk, inference/search/two_tower.py:98-105.
 *   index [N,H] fp32 (index_bf16 == 0) or bf16 (== 1), contiguous (what index_documents
 *   writes, two_tower.py:69).  queries [nq,H] fp32.  cosine == 0: raw dot products (valid for
 *   the unit rows both towers emit); cosine == 1: divide by max(||q||,1e-8)*max(||d||,1e-8)
 *   exactly like F.cosine_similarity.
 *   out_scores [nq,k] fp32 and out_ids [nq,k] int64 (global id = row + id_offset), sorted by
 *   descending score, ties -> LOWER id first; k <= min(N, TT_TOPK_MAX).
 * tt_topk_merge: merge R sorted candidate lists per query (the sharded / multi-GPU step):
 *   list r = scores + r*score_rank_stride, ids + r*id_rank_stride (elements; 0 = dense [R,nq,k]) -> [nq,k], same
 *   ordering rule.  Strides let the merge read (score,id) records straight out of ONE all-gathered byte buffer.
 */
#define TT_TOPK_MAX 1024
size_t tt_topk_scan_workspace(int64_t N, int H, int nq, int k);
int tt_topk_scan(const void* index, int index_bf16, const float* queries, int64_t N, int H,
                 int nq, int k, int cosine, int64_t id_offset,
                 float* out_scores, int64_t* out_ids,
                 void* workspace, size_t workspace_bytes, void* stream);
int tt_topk_merge(const float* scores, const int64_t* ids, int R, int nq, int k,
                  int64_t score_rank_stride, int64_t id_rank_stride,
                  float* out_scores, int64_t* out_ids, void* stream);
/* Batched scan on the tcgen05 tensor cores (bf16 index): S = Q D^T for up to 128 queries per pass over the index
 * (M = 128 queries x N = 128 documents per tcgen05.mma tile, document tiles streamed by TMA), fused with the same exact
 * top-k; the reference API scores one query per call (two_tower.py:72-115), so nq queries read the index nq times there
 * and ceil(nq / 128) times here.  fp32 queries are split into bf16 hi + lo parts, both multiplied into the same fp32
 * accumulator: scores match tt_topk_scan on the same bf16 index to ~1e-5 relative, ids identical except at such ties.
 * row_inv_norms (nullable, [N] from tt_index_row_inv_norms, computed once per index): cosine scores
 * (dot / (max(|q|,1e-8) max(|d|,1e-8))); null = raw dot products.  H % 64 == 0, H <= 256, k <= 128.
 */
int tt_topk_scan_batched_ok(int H, int k);
size_t tt_topk_scan_batched_workspace(int64_t N, int H, int nq);
int tt_index_row_inv_norms(const void* index_bf16, int64_t N, int H, float* out, void* stream);
int tt_topk_scan_batched(const void* index_bf16, const float* queries, int64_t N, int H, int nq, int k,
                         const float* row_inv_norms, int64_t id_offset, float* out_scores, int64_t* out_ids,
                         void* workspace, size_t workspace_bytes, void* stream);
/* fp32 -> bf16 row copy used by index_documents when the index is kept in bf16. */
int tt_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ---- multi-GPU exchange over NVLink peer memory ---------------------------------------------------
 * One process per GPU.  Every rank allocates one exchange buffer (tt_p2p_alloc, tt_p2p_buffer_bytes), exports it
 * (CUDA IPC handle, 64 bytes) and imports every peer's; tt_p2p_t lists all ranks' buffers as mapped in this process
 * (base[rank] is the local one).  Layout: 256-byte header (arrival counters, round, ticket) + world slots of slot_bytes.
 * tt_p2p_allgather : ONE kernel stores `bytes` of src into slot `rank` of every rank's buffer (16-byte stores over
 *                    NVLink), signals, and spins until all ranks' slots for this round have arrived in the LOCAL buffer.
 *                    When it retires the gathered data is base[rank] + 256.  bytes may differ from call to call (<= slot_bytes).
 *                    Capturable in a CUDA graph; replaces an NCCL all-gather whose fixed cost dominates at these sizes.
 * tt_p2p_sum_slots : out[i] = sum_r slot_r[i] in rank order -- with an all-gather of the gradients this is an all-reduce
 *                    whose result is bitwise identical on every rank.
 */
typedef struct tt_p2p_s {
  int world, rank;
  size_t slot_bytes;          /* multiple of 256 */
  void* base[8];              /* exchange buffer of rank p as mapped in this process */
  int double_buffered;        /* rounds alternate between two slot sets: safe when this is the only exchange between two
                                 uses of the same buffer (see p2p.cu); consumers then read via tt_p2p_sum_slots */
  int ctas;                   /* CTAs per exchange kernel, the SAME on every rank and for every call on this exchange
                                 (0: tt_p2p_allgather_ctas(slot_bytes)); calls may then move any byte count <= slot_bytes */
  int timeout_s;              /* seconds a rank waits for its peers before it gives up (0: 600); see tt_p2p_status */
} tt_p2p_t;
size_t tt_p2p_buffer_bytes(int world, size_t slot_bytes, int double_buffered);
int tt_p2p_alloc(size_t bytes, void** ptr);
int tt_p2p_free(void* ptr);
int tt_p2p_export(void* ptr, void* handle64);
int tt_p2p_import(const void* handle64, void** ptr);
int tt_p2p_unimport(void* ptr);
int tt_p2p_allgather_ctas(size_t bytes);
int tt_p2p_allgather(const tt_p2p_t* x, const void* src, size_t bytes, void* stream);
/* Synchronous: *timed_out_rank = the rank an exchange kernel gave up waiting for (its output is then garbage), -1 if none.
 * A time-out never traps: the CUDA context stays usable and the host decides (raise, fall back to NCCL, retry). */
int tt_p2p_status(const tt_p2p_t* x, int* timed_out_rank);
int tt_p2p_sum_slots(const tt_p2p_t* x, size_t n_floats, float* out, void* stream);
/* Row-sharded search in two launches per query batch (inference/search/two_tower.py:98-105 over a row-sharded index,
 * SURVEY 8e): tt_topk_scan's scan kernel, then ONE kernel per query that (1) selects this shard's k best block candidates,
 * (2) stores them -- rebased to global row numbers, id_offset + N < 2^32 -- into every rank's exchange slot over NVLink and
 * bumps the arrival counters (the protocol of tt_p2p_allgather), (3) waits for every shard's k candidates and (4) selects the
 * global top-k.  Every rank gets the same (out_scores, out_ids [nq,k], global ids; ties -> lower global row).  `x`: a
 * double-buffered exchange whose `ctas` field equals nq and whose slots hold nq * k * 8 bytes (tt_topk_scan_p2p_ok); every
 * rank calls it in step, on one stream.  workspace: tt_topk_scan_workspace(N, H, nq, k). */
int tt_topk_scan_p2p_ok(const tt_p2p_t* x, int nq, int k);
int tt_topk_scan_p2p(const void* index, int index_bf16, const float* queries, int64_t N, int H, int nq, int k,
                     int cosine, int64_t id_offset, const tt_p2p_t* x, float* out_scores, int64_t* out_ids, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- optimizer: torch.optim.AdamW(model.parameters(), lr), train.py:359 ------
 * One fused AdamW step over a flat fp32 parameter buffer (ATen _single_tensor_adamw
 * semantics and operation order, amsgrad off; hyper-parameters are doubles, rounded to fp32
 * exactly where torch rounds them).  step_count: device int64[2], zero-initialised by the caller:
 * [0] = number of steps already taken (the kernel uses [0]+1 for bias correction and the last block
 * to finish increments it -- graph-capturable, no host state), [1] = scratch arrival ticket.  param_bf16 (nullable): refreshed bf16 shadow.
 */
int tt_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay,
                  int64_t* step_count, void* param_bf16, void* stream);
/* Same step; in addition the kernel copies the scalar *publish_src (e.g. the step's loss, twotower/train.py:139
 * loss.item()) to *publish_dst, which may be MAPPED PINNED HOST memory: the value reaches the host from inside the
 * step's last launch, so a captured step needs no device-to-host copy node.  Both null == tt_adamw_step. */
int tt_adamw_step_publish(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                          int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay,
                          int64_t* step_count, void* param_bf16, const float* publish_src, float* publish_dst,
                          void* stream);
/* Same step with the gradient taken from the slots of a (double-buffered) peer-memory all-gather that has just completed
 * on this stream (tt_p2p_allgather of every rank's local gradient): the slots are summed in rank order while they are
 * read -- the data-parallel all-reduce (twotower semantics: one model, mean over the global batch) costs no reduction
 * launch, and the parameters stay bitwise identical on all ranks.  grad_sum (nullable) receives the summed gradient. */
/* tt_adamw_step_publish with a SECOND gradient for one segment of the flat buffer (parameters extra_offset .. + extra_n, both
 * multiples of 4): grad[i] + extra_grad[i - extra_offset] is what the update uses, and it is written back to grad.  Used by
 * FusedTrainer with untied towers: the two towers' backward passes run concurrently on two streams, each writing its own
 * embedding-table gradient (twotower/train.py:120-139: one table shared by both towers), and the sum costs no launch. */
int tt_adamw_step_extra(float* param, float* grad, const float* extra_grad, int64_t extra_offset, int64_t extra_n,
                        float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1, double beta2, double eps,
                        double weight_decay, int64_t* step_count, void* param_bf16, const float* publish_src,
                        float* publish_dst, void* stream);
int tt_adamw_step_p2p(float* param, float* grad_sum, const tt_p2p_t* grad_exchange, float* exp_avg, float* exp_avg_sq,
                      int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay,
                      int64_t* step_count, void* param_bf16, const float* publish_src, float* publish_dst,
                      void* stream);

/* ---- self-test hook -------------------------------------------------------------------------
 * C[M,N] (fp32) = A * B on the tcgen05 tensor cores; used by the GPU tests to pin the UMMA / TMA
 * descriptor encodings for both operand majors.  a_mn_major: A stored [K,M] (else [M,K]);
 * b_mn_major: B stored [K,N] (else [N,K]); bf16 operands.  splits > 1 needs partial[splits*M*N].
 */
int tt_selftest_tc_gemm(const void* a_bf16, int a_mn_major, const void* b_bf16, int b_mn_major,
                        int M, int N, int K, float* c, int splits, float* partial, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H */
