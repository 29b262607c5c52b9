"""Fused training step: forward + loss + backward + AdamW as one CUDA-graph replay.

Reference: the body of ``train_epoch``'s batch loop, twotower/train.py:107-154 --
``.to(device)`` x3 (:107-109), ``model(q, pos, neg)`` (:120), ``loss_fn`` (:133),
``zero_grad / backward / step`` (:137-139), monitoring cosines + ``.item()`` x3 (:144-154).
The reference launches ~100 eager ATen kernels per step and synchronises the host three
times; here the same arithmetic is ~12 hand-written kernels captured in one CUDA graph,
all parameters / gradients / Adam moments live in flat buffers (one optimizer launch), no
gradient zeroing is needed (every gradient is overwritten by its single producer) and the
monitoring values are by-products of the loss kernels read back asynchronously.

Data parallel (one process per GPU): rows are sharded; in-batch negatives become GLOBAL
negatives through an all-gather of the document embeddings (``parallel.global_inbatch_*``),
and the flat gradient buffer is all-reduced once per step.

The modules of ``two_towers_b200.encoders`` also work under plain ``loss.backward()`` +
``torch.optim.AdamW`` exactly like the reference loop; this class is the fast path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops, parallel
from ._lib import check
from .encoders import AveragePoolingTower, MeanPoolingTower, TwoTower


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _StopStep(Exception):
    pass


class FusedTrainer:
    def __init__(self, model: TwoTower, loss: str = "in_batch", temperature: float = 0.1, margin: float = 0.2,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 batch_size: int = 256, max_len: int = 64, precision=None, process_group=None,
                 global_negatives: bool = True, use_cuda_graph: bool = True, id_dtype=torch.int64,
                 p2p: Optional[bool] = None):
        if loss not in ("in_batch", "triplet"):
            raise ValueError("FusedTrainer supports loss 'in_batch' or 'triplet'")
        self.model = model
        self.loss_name = loss
        self.temperature, self.margin = float(temperature), float(margin)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.B, self.L = int(batch_size), int(max_len)
        self.prec = ops.resolve_precision(precision)
        self.group = process_group
        self.rank, self.world = parallel.world(process_group)
        self.global_negatives = bool(global_negatives) and self.world > 1 and loss == "in_batch"
        self.use_graph = use_cuda_graph
        self.lib = _lib.load()

        qt, dt = model.query_tower, model.document_tower
        if dt.embedding is not qt.embedding:
            raise ValueError("FusedTrainer: both towers must share one embedding object (build_two_tower always does, "
                             "encoders.py:265,270)")
        self.tied = qt is dt
        self.passes = 2 if loss == "in_batch" else 3
        table = qt.embedding.embedding.weight
        if not table.is_cuda:
            raise RuntimeError("FusedTrainer: model must live on a CUDA (B200) device; there is no CPU fallback")
        self.dev = table.device
        self.table = table
        self.V, self.E = table.shape
        self.H = qt.hidden_dim
        self.train_table = bool(table.requires_grad)
        B, L, P = self.B, self.L, self.passes
        R = P * B
        # row groups: (tower, first row, row count)
        self.groups = [(qt, 0, R)] if self.tied else [(qt, 0, B), (dt, B, (P - 1) * B)]

        # ---- flat parameter / gradient / moment buffers --------------------------------------
        params = [p for p in model.parameters() if p.requires_grad]
        self._params = params
        pad = lambda n: (n + 63) // 64 * 64          # every view starts 256-byte aligned
        total = sum(pad(p.numel()) for p in params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.step_count = torch.zeros(2, dtype=torch.int64, device=self.dev)      # [steps taken, ticket]
        off = 0
        self.offsets = {}
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.offsets[id(p)] = off
                # keep 16-byte alignment of every view (vector loads in the kernels)
                self.flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)
                p.grad = self.flat_grad[off:off + n].view(p.shape)
                off += pad(n)
        self.n_params = total

        # ---- static activations -------------------------------------------------------------
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.ids = torch.zeros(R, L, dtype=id_dtype, device=self.dev)
        # input pipelining (prefetch / step): TWO static id buffers, each with its own captured graph and loss slot, so a
        # host->device copy lands directly in the buffer the next replay reads (no staging copy on the step's stream)
        self._ids_bufs = [self.ids, None]
        self._ids_cur = self.ids
        self._loss_dst = None
        self.pooled = torch.empty(R, self.E, **f32)
        self.inv_len = torch.empty(R, **f32)
        self.y = torch.empty(R, self.H, **f32)
        # the tensor-core loss backward leaves its per-split partial gradients as slices [parts, R, H]; the
        # tower backward sums them while it reads dy (no reduction kernel, no extra pass)
        self.dy_parts = 1
        fast = (self.prec == _lib.TT_PREC_BF16 and loss == "in_batch" and self.H % 64 == 0 and self.H <= 256 and
                all(isinstance(t, MeanPoolingTower) for t, _, _ in self.groups))
        self.global_fast = bool(fast and self.global_negatives and B % 64 == 0)
        self.local_fast = bool(fast and not self.global_negatives)
        if fast:
            # persistent, self-zeroing scratch: lets the loss forward finish lse / loss inside its own launch
            self.ce_sync = torch.zeros(int(self.lib.tt_inbatch_ce_sync_bytes(B)), dtype=torch.uint8, device=self.dev)
        if self.local_fast:
            self.dy_parts = int(self.lib.tt_inbatch_ce_bwd_nparts(B, B, self.H, self.prec))
            self.ce_ws = torch.empty(int(self.lib.tt_inbatch_ce_fwd_ex_workspace(B, B)), dtype=torch.uint8, device=self.dev)
        elif self.global_fast:
            Bg = B * self.world
            self.dy_parts = int(self.lib.tt_inbatch_ce_bwd_nparts_ex(B, Bg, B, Bg, self.H))
            # one all-gather moves [Q_r | D_r] (bf16) of every rank; the loss kernels index it block-wise in place
            self.yg_bf16 = torch.empty(self.world * 2 * B, self.H, dtype=torch.bfloat16, device=self.dev)
            self.lse_g = torch.empty(Bg, **f32)
            self.ce_ws = torch.empty(int(self.lib.tt_inbatch_ce_fwd_ex_workspace(B, Bg)), dtype=torch.uint8, device=self.dev)
        # the loss backward can also finish the normalise backward itself (CTA pairs exchange accumulator halves through
        # distributed shared memory): it then emits dz (bf16) + column sums and the [parts, R, H] fp32 slices never exist
        Bg = B * self.world
        self.ce_fused = bool(os.environ.get("TT_CE_FUSED", "1") != "0" and
                             fast and (self.local_fast or self.global_fast) and B % 32 == 0 and self.passes == 2 and
                             self.lib.tt_inbatch_ce_bwd_fused_ok(B, Bg if self.global_fast else B, B,
                                                                 Bg if self.global_fast else B, self.H))
        # one-pass step: unit-norm tower outputs bound every logit by 1/temperature, so the loss forward and the query
        # gradient share ONE pass over S (tt_inbatch_ce_fwd_dq) and the document gradient is the only other loss launch
        self.onepass = bool(os.environ.get("TT_CE_ONEPASS", "1") != "0" and os.environ.get("TT_CE_FUSED", "1") != "0" and fast and (self.local_fast or self.global_fast) and
                            B % 32 == 0 and self.passes == 2 and
                            self.lib.tt_inbatch_ce_onepass_ok(B, Bg if self.global_fast else B, self.H, 1.0 / self.temperature))
        # both loss launches as ONE kernel with a grid-wide barrier (tt_inbatch_ce_onepass): bitwise the same results, measured
        # neutral (105.5 vs 104.5 us per step) -- the second launch's boundary is not what the step pays for -- so off by default
        self.onelaunch = bool(self.onepass and self.local_fast and os.environ.get("TT_CE_ONELAUNCH", "0") == "1")
        if self.onepass:
            self.ce_fused = True
            self.onepass_sync = torch.zeros(int(self.lib.tt_inbatch_ce_onepass_sync_bytes(B)), dtype=torch.uint8, device=self.dev)
        # stored-E form (per-rank negatives, B x B bf16 small enough to stay in L2): the first loss launch also stores its E
        # tiles and the document gradient becomes one plain product (tt_inbatch_ce_dd_stash) -- S is formed once per step
        self.ce_stash = None
        if (self.onepass and self.local_fast and not self.onelaunch and os.environ.get("TT_CE_STASH", "1") != "0" and
                self.lib.tt_inbatch_ce_stash_ok(B, B, self.H)):
            self.ce_stash = torch.empty(int(self.lib.tt_inbatch_ce_stash_bytes(B, B, self.H)), dtype=torch.uint8, device=self.dev)
        if self.ce_fused:
            self.dz_bf16 = torch.empty(R, self.H, dtype=torch.bfloat16, device=self.dev)
            self.dz_colsum = torch.empty(R // 32, self.H, **f32)
        # multi-GPU exchanges over NVLink peer memory (one kernel each) instead of NCCL: tower outputs, lse, gradients
        if p2p is None:
            p2p = os.environ.get("TT_P2P", "1") != "0"
        self.p2p = bool(p2p and self.world > 1 and self.world <= 8 and self.flat_grad.numel() % 4 == 0 and
                        (not self.global_fast or (2 * B * self.H * 2) % 256 == 0))
        # the peer-memory gradient exchange pushes every rank's full fp32 gradient to every peer (world x the bytes of a
        # ring all-reduce, 2 x world x n x 4 bytes of buffer): right for the KB..MB gradients of the char towers where the
        # collective's fixed cost dominates, wrong for a 480 MB word-embedding table -- those go through NCCL
        self.p2p_grad = bool(self.p2p and self.flat_grad.numel() * 4 <= int(os.environ.get("TT_P2P_GRAD_MAX_BYTES", 64 << 20)))
        self.x_grad = None
        if self.p2p:
            try:
                if self.p2p_grad:
                    self.x_grad = parallel.P2PExchange(self.flat_grad.numel() * 4, self.group, self.dev, double_buffered=True)
                if self.global_fast:
                    # documents and queries travel separately: the loss forward needs only D, so the Q exchange (needed by
                    # the dD pass of the backward) runs on a side stream underneath the forward kernel
                    self.x_d = parallel.P2PExchange(B * self.H * 2, self.group, self.dev)
                    self.x_q = parallel.P2PExchange(B * self.H * 2, self.group, self.dev)
                    self.x_lse = parallel.P2PExchange(B * 4, self.group, self.dev)
                    self._side = torch.cuda.Stream(device=self.dev)
            except RuntimeError as e:                       # raised on ALL ranks together (see P2PExchange): NCCL path
                import warnings
                warnings.warn(f"FusedTrainer: {e}; using NCCL collectives")
                self.p2p = self.p2p_grad = False
        # the loss forward consumes the gathered documents block by block as they land (tt_inbatch_ce_fwd_dq_p2p)
        self.gated = bool(self.p2p and self.global_fast and self.onepass and B % 512 == 0 and (B * self.H * 2) % 256 == 0 and
                          os.environ.get("TT_CE_GATED", "1") != "0")
        if self.p2p:
            if self.global_fast:
                self.dg_bf16 = self.x_d.gathered(torch.bfloat16, (B, self.H)).view(self.world * B, self.H)
                self.qg_bf16 = self.x_q.gathered(torch.bfloat16, (B, self.H)).view(self.world * B, self.H)
                self.lse_g = self.x_lse.gathered(torch.float32, (B,)).reshape(self.world * B) \
                    if (B * 4) % 256 == 0 else None
                if self.lse_g is None or not self.lse_g.is_contiguous():
                    self.lse_g = torch.empty(self.world * B, **f32)       # odd B: gather into slots, then pack (one small copy)
                    self._lse_pack = True
                else:
                    self._lse_pack = False
        self.dy_part_stride = R * self.H if self.dy_parts > 1 else 0
        self.dy_all = torch.empty(1 if self.ce_fused else self.dy_parts, R, self.H, **f32)
        self.dy = self.dy_all[0]
        self.dpooled = torch.empty(R, self.E, **f32)
        self.saved: List[Dict[str, torch.Tensor]] = []
        for tower, r0, nr in self.groups:
            if isinstance(tower, MeanPoolingTower):
                self.saved.append(dict(h1=torch.empty(nr, self.H, **f32), z=torch.empty(nr, self.H, **f32)))
            elif isinstance(tower, AveragePoolingTower):
                if tower.has_projection:
                    self.saved.append(dict(a=torch.empty(nr, self.H, **f32), stats=torch.empty(nr, 2, **f32),
                                           z=torch.empty(nr, self.H, **f32)))
                else:
                    self.saved.append({})
            else:
                raise TypeError(f"FusedTrainer: unsupported tower {type(tower).__name__}")
        bf = self.prec == _lib.TT_PREC_BF16
        # E % 8 != 0 (the word tower's E = 300): bf16 rows of x / W1 need a 16-byte pitch for TMA, so the tower calls
        # convert them into padded rows themselves and no contiguous shadow is kept here
        self.e_shadow = bool(bf and self.E % 8 == 0)
        self.pooled_bf16 = torch.empty(R, self.E, dtype=torch.bfloat16, device=self.dev) if self.e_shadow else None
        self.inv_norm = torch.empty(R, **f32) if bf else None
        self.y_bf16 = torch.empty(R, self.H, dtype=torch.bfloat16, device=self.dev) if bf else None
        # bf16 shadow of the flat parameter buffer, refreshed by the AdamW kernel -> the tensor-core GEMMs
        # read weights without any per-step conversion kernel
        self.flat_bf16 = ops.cast_bf16(self.flat) if bf else None
        # small vocabularies on the tensor-core path: the forward also emits the pooling matrix P (x = P table) and the
        # tower backward forms P^T da1 once -- no dx, no separate embedding backward (see tt_mlp_embed_t)
        self.embed_fused = bool(bf and self.train_table and self.V <= 1024 and self.V % 8 == 0 and self.E % 4 == 0 and
                                all(isinstance(t, MeanPoolingTower) for t, _, _ in self.groups))
        self.pool_bf16 = torch.empty(R, self.V, dtype=torch.bfloat16, device=self.dev) if self.embed_fused else None
        # ... and for char-sized vocabularies the tower kernel forms x = P table itself: the gather kernel shrinks to a
        # histogram and the pooled activations never reach HBM
        self.embed_in_tower = bool(self.embed_fused and os.environ.get("TT_EMBED_IN_TOWER", "1") != "0" and
                                   self._shadow(self.table) is not None and
                                   self.lib.tt_mlp_fwd_embed_ok(self.E, self.H, self.V))
        # ... and could build the histogram too (tt_mlp_embed_t.ids: one launch less per step, bitwise the same P).  Measured
        # SLOWER (tower kernel 10.0 -> 17.7 us with a warp per row, 21.9 us with a thread per row, against 3.4 us for the
        # stand-alone histogram launch that spreads the rows over 8192 warps): 64 CTAs x 8 warps cannot hide the per-row
        # chain of shared-memory updates.  Kept for reference, off by default.
        self.pool_in_tower = bool(self.embed_in_tower and self.L <= 255 and self.V <= 128 * (self.E // 64) and
                                  os.environ.get("TT_POOL_IN_TOWER", "0") == "1")
        self.embed_ws = (torch.empty(int(max(self.lib.tt_mlp_embed_workspace(self.V, self.H, nr) for _, _, nr in self.groups)),
                                     dtype=torch.uint8, device=self.dev) if self.embed_fused else None)
        self.h1_bf16 = [torch.empty(nr, self.H, dtype=torch.bfloat16, device=self.dev) if bf else None
                        for _, _, nr in self.groups]
        self.loss = torch.zeros((), **f32)
        self.lse = torch.empty(B, **f32)
        self.pos_mean = torch.zeros((), **f32)
        self.neg_mean = torch.zeros((), **f32)
        self.sims = torch.empty(3 * B, **f32)
        self.grad_scale = torch.full((), 1.0 / self.world, **f32)
        lib = self.lib
        nb = max(lib.tt_mlp_workspace(R, self.E, self.H, self.prec), lib.tt_proj_ln_workspace(R, self.E, self.H, self.prec),
                 lib.tt_embed_pool_bwd_workspace(R, L, self.V, self.E),
                 lib.tt_inbatch_ce_workspace(B * self.world, B * self.world, self.H, self.prec))
        self.ws = torch.empty(int(nb), dtype=torch.uint8, device=self.dev)
        # untied towers: the two towers' forward and backward launches are independent chains (3 launches each in the backward),
        # so they run concurrently on two streams inside the captured step.  The second tower then needs its own workspaces
        # and -- both towers share ONE embedding table (twotower/train.py:120-139) -- its own table-gradient buffer, which the
        # optimizer launch adds to the first tower's (tt_adamw_step_extra: no launch for the sum, same order of additions as
        # the serial schedule).  With a process group the sum would have to precede the gradient exchange: serial there.
        self.par_towers = bool(len(self.groups) == 2 and bf and os.environ.get("TT_TOWER_PAR", "1") != "0" and
                               all(isinstance(t, MeanPoolingTower) for t, _, _ in self.groups) and
                               (self.world == 1 or not self.embed_fused) and self.n_params % 4 == 0)
        if self.par_towers:
            self.ws2 = torch.empty_like(self.ws)
            self.embed_ws2 = torch.empty_like(self.embed_ws) if self.embed_ws is not None else None
            self.table_grad2 = torch.zeros((self.table.numel() + 3) // 4 * 4, **f32) if self.embed_fused else None
            self._tower_stream = torch.cuda.Stream(device=self.dev)
        self._loss_ring = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]   # read_loss_async()
        self._loss_ring_g = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]  # written by graph 0 / 1
        self._graphs: List[Optional[torch.cuda.CUDAGraph]] = [None, None]
        self._last_gi = -1                                  # graph of the last pipelined step (-1: explicit-ids step)
        self._loss_ev = [torch.cuda.Event() for _ in range(2)]
        self._loss_slot = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._stage = None
        self._staged = False
        self.steps_done = 0

    # ---------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _shadow(self, param):
        """bf16 view of `param` inside the flat shadow buffer (None in fp32 mode / frozen params)."""
        if self.flat_bf16 is None or id(param) not in self.offsets:
            return None
        off = self.offsets[id(param)]
        return self.flat_bf16[off:off + param.numel()]

    def _dropout_cfg(self, tower, gi: int):
        """(p, training, seed) of an avg_pool tower's Dropout (encoders.py:102): the reference loop runs model.train(), so
        Dropout(p) is active.  The mode is read when the step is recorded (CUDA graph): changing model.train()/eval()
        afterwards needs a new trainer.  The seed mixes the tower instance, its row group and the rank; the per-step
        variation comes from the device-side step counter (tt_proj_ln_fwd seed_step)."""
        train = int(bool(self.model.training) and tower.dropout_p > 0)
        seed = (tower._seed + 0xA24BAED4963EE407 * (gi + 1) + 0x9FB21C651E98DF25 * (self.rank + 1)) % (1 << 64)
        return float(tower.dropout_p), train, seed

    def _tower_fwd(self, gi: int, second: bool = False):
        """second: this call runs on the second tower stream (par_towers) and uses the second set of workspaces."""
        tower, r0, nr = self.groups[gi]
        lib, s, sv = self.lib, self._stream(), self.saved[gi]
        ws = self.ws2 if second else self.ws
        x, y = self.pooled[r0:r0 + nr], self.y[r0:r0 + nr]
        yb = self.y_bf16[r0:r0 + nr] if self.y_bf16 is not None else None
        xb = self.pooled_bf16[r0:r0 + nr] if self.pooled_bf16 is not None else None
        if isinstance(tower, MeanPoolingTower):
            l1, l2 = tower.feed_forward[0], tower.feed_forward[2]
            pure = (self.dy_parts > 1 or self.global_fast or self.ce_fused) and yb is not None
            y_ptr = None if pure else y                     # fp32 y unused on the pure bf16 path
            # pure bf16 path: the normalise step is saved as (y_bf16, 1/|z|); the fp32 pre-normalise tensor is never written
            inv = self.inv_norm[r0:r0 + nr] if (pure and self.H <= 512) else None
            z_ptr = None if inv is not None else sv["z"]
            emb = None
            if self.embed_in_tower:
                emb = _lib.MlpEmbed(self.pool_bf16[r0:r0 + nr].data_ptr(), self.V, self.table.data_ptr(),
                                    self._shadow(self.table).data_ptr(), None, 0, None, 0)
                if self.pool_in_tower:                      # the tower kernel builds P from the ids itself: no histogram launch
                    emb.ids = self._ids_cur[r0:r0 + nr].data_ptr()
                    emb.id_bytes = 8 if self.ids.dtype == torch.int64 else 4
                    emb.L = self.L
                    emb.inv_len = self.inv_len[r0:r0 + nr].data_ptr()
            check(lib.tt_mlp_fwd(_p(x), _p(l1.weight), _p(l1.bias), _p(l2.weight), _p(l2.bias), nr, self.E, self.H,
                                 _p(sv["h1"]), _p(z_ptr), _p(y_ptr), _p(yb), _p(xb),
                                 _p(self._shadow(l1.weight) if self.e_shadow else None),
                                 _p(self._shadow(l2.weight)), _p(self.h1_bf16[gi]), _p(inv),
                                 C.byref(emb) if emb is not None else None, self.prec, _p(ws),
                                 ws.numel(), s), "tt_mlp_fwd")
        elif tower.has_projection:
            lin, ln = tower.projection[0], tower.projection[2]
            p_drop, train, seed = self._dropout_cfg(tower, gi)
            check(lib.tt_proj_ln_fwd(_p(x), _p(lin.weight), _p(lin.bias), _p(ln.weight), _p(ln.bias), nr, self.E,
                                     self.H, 1, p_drop, train, seed, _p(self.step_count), _p(sv["a"]), _p(sv["stats"]),
                                     _p(sv["z"]), _p(y), self.prec, _p(self.ws), self.ws.numel(), s), "tt_proj_ln_fwd")
        else:
            check(lib.tt_proj_ln_fwd(_p(x), None, None, None, None, nr, self.E, self.H, 0, 0.0, 0, 0, None, None, None,
                                     None, _p(y), self.prec, None, 0, s), "tt_proj_ln_fwd")

    def _tower_bwd(self, gi: int, second: bool = False):
        tower, r0, nr = self.groups[gi]
        lib, s, sv = self.lib, self._stream(), self.saved[gi]
        ws = self.ws2 if second else self.ws
        x, dy = self.pooled[r0:r0 + nr], self.dy[r0:r0 + nr]          # slice 0; further slices dy_part_stride apart
        dx = self.dpooled[r0:r0 + nr] if self.train_table else None
        xb = self.pooled_bf16[r0:r0 + nr] if self.pooled_bf16 is not None else None
        emb = None
        if self.embed_fused:
            if second:                                      # own table-gradient buffer, summed inside the optimizer launch
                emb = _lib.MlpEmbed(self.pool_bf16[r0:r0 + nr].data_ptr(), self.V, self.table.data_ptr(), None,
                                    self.table_grad2.data_ptr(), 0, self.embed_ws2.data_ptr(), self.embed_ws2.numel())
            else:
                emb = _lib.MlpEmbed(self.pool_bf16[r0:r0 + nr].data_ptr(), self.V, self.table.data_ptr(), None,
                                    self.table.grad.data_ptr(), 1 if gi > 0 else 0, self.embed_ws.data_ptr(), self.embed_ws.numel())
            dx = None
        if isinstance(tower, MeanPoolingTower):
            l1, l2 = tower.feed_forward[0], tower.feed_forward[2]
            yb = self.y_bf16[r0:r0 + nr] if self.y_bf16 is not None else None
            pure = (self.dy_parts > 1 or self.global_fast or self.ce_fused) and yb is not None and self.H <= 512
            inv = self.inv_norm[r0:r0 + nr] if pure else None
            dzb = self.dz_bf16[r0:r0 + nr] if self.ce_fused else None
            dzc = self.dz_colsum[r0 // 32:(r0 + nr) // 32] if self.ce_fused else None
            check(lib.tt_mlp_bwd(_p(dy), _p(x), _p(l1.weight), _p(l2.weight), _p(sv["h1"]), _p(None if pure else sv["z"]), nr, self.E,
                                 self.H, _p(dx), _p(l1.weight.grad), _p(l1.bias.grad), _p(l2.weight.grad),
                                 _p(l2.bias.grad), _p(xb), _p(self._shadow(l1.weight) if self.e_shadow else None),
                                 _p(self._shadow(l2.weight)),
                                 _p(self.h1_bf16[gi]), self.dy_parts, self.dy_part_stride,
                                 C.byref(emb) if emb is not None else None, _p(yb if pure else None), _p(inv),
                                 _p(dzb), _p(dzc), self.prec, _p(ws), ws.numel(), s), "tt_mlp_bwd")
        elif tower.has_projection:
            lin, ln = tower.projection[0], tower.projection[2]
            p_drop, train, seed = self._dropout_cfg(tower, gi)
            check(lib.tt_proj_ln_bwd(_p(dy), _p(x), _p(lin.weight), _p(ln.weight), _p(sv["a"]), _p(sv["stats"]),
                                     _p(sv["z"]), nr, self.E, self.H, 1, p_drop, train, seed, _p(self.step_count), _p(dx),
                                     _p(lin.weight.grad), _p(lin.bias.grad), _p(ln.weight.grad), _p(ln.bias.grad),
                                     self.prec, _p(self.ws), self.ws.numel(), s), "tt_proj_ln_bwd")
        elif dx is not None:
            check(lib.tt_proj_ln_bwd(_p(dy), _p(x), None, None, None, None, None, nr, self.E, self.H, 0, 0.0, 0, 0, None,
                                     _p(dx), None, None, None, None, self.prec, None, 0, s), "tt_proj_ln_bwd")

    def _mark(self, name: str) -> None:
        """Developer aid (tools/step_timeline.py): phase boundaries of the step.  With `self._trace` set to a list an eager
        step leaves one CUDA event per boundary; with `self._stop_after` set the step ends at that boundary
        (`step_prefix`), so that prefixes of the step can be captured and timed as graphs of their own."""
        tr = getattr(self, "_trace", None)
        if tr is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream())
            tr.append((name, ev))
        if getattr(self, "_stop_after", None) == name:
            if getattr(self, "_side_open", False):          # join the forked exchange stream before the prefix ends
                torch.cuda.current_stream().wait_stream(self._side)
                self._side_open = False
            raise _StopStep()

    def step_prefix(self, stop_after: str) -> None:
        """Run the step's launches up to and including phase `stop_after` (names as in tools/step_timeline.py).  Collective
        with a process group.  Leaves parameters untouched unless the prefix reaches the optimizer."""
        self._stop_after = stop_after
        try:
            self._step_impl()
        except _StopStep:
            pass
        finally:
            self._stop_after = None

    def _step_impl(self):
        self._mark("start")
        lib, B, H, P = self.lib, self.B, self.H, self.passes
        R = P * B
        idb = 8 if self.ids.dtype == torch.int64 else 4
        s = self._stream()
        tower_pools = self.embed_in_tower                 # histogram only: the tower kernel multiplies P by the table
        if not self.pool_in_tower:                        # ... or nothing at all: the tower kernel also builds P from the ids
            check(lib.tt_embed_pool_fwd(_p(self._ids_cur), idb, _p(self.table), R, self.L, self.V, self.E,
                                        None if tower_pools else _p(self.pooled), _p(self.inv_len),
                                        None if tower_pools else _p(self.pooled_bf16), _p(self.pool_bf16), s), "tt_embed_pool_fwd")
        self._mark("embed_pool_fwd")
        if self.par_towers:
            self._towers_parallel(self._tower_fwd)
        else:
            for gi in range(len(self.groups)):
                self._tower_fwd(gi)
        self._mark("tower_fwd")
        q, d = self.y[:B], self.y[B:2 * B]
        dq, dd = self.dy[:B], self.dy[B:2 * B]
        if self.loss_name == "in_batch":
            inv_t = 1.0 / self.temperature
            if self.global_fast:
                self._global_inbatch_fast(s, inv_t)
            elif self.global_negatives:
                loss, lse, d_glob = parallel.global_inbatch_fwd(q, d, self.temperature, ops, self.group, self.prec)
                self.loss.copy_(loss)
                g_dq, g_dd = parallel.global_inbatch_bwd(q, d, d_glob, lse, self.temperature, ops, self.group, self.prec)
                dq.copy_(g_dq)
                dd.copy_(g_dd)
            elif self.onepass:
                self._local_loss_onepass(s)
            else:
                self._local_loss_fwd(s)
                self._local_loss_bwd(s)
        else:
            n, dn = self.y[2 * B:3 * B], self.dy[2 * B:3 * B]
            check(lib.tt_triplet_fwd(_p(q), _p(d), _p(n), B, H, self.margin, _p(self.loss), _p(self.sims),
                                     _p(self.pos_mean), _p(self.neg_mean), s), "tt_triplet_fwd")
            check(lib.tt_triplet_bwd(_p(q), _p(d), _p(n), _p(self.sims), B, H, self.margin,
                                     _p(self.grad_scale) if self.world > 1 else None, _p(dq), _p(dd), _p(dn), s),
                  "tt_triplet_bwd")
        self._mark("loss")
        if self.par_towers:
            self._towers_parallel(self._tower_bwd)
        else:
            for gi in range(len(self.groups)):
                self._tower_bwd(gi)
        if self.train_table and not self.embed_fused:
            check(lib.tt_embed_pool_bwd(_p(self._ids_cur), idb, _p(self.inv_len), _p(self.dpooled), R, self.L, self.V,
                                        self.E, _p(self.table.grad), _p(self.ws), self.ws.numel(), s),
                  "tt_embed_pool_bwd")
        self._mark("tower_bwd")
        pub = self._loss_dst                                # pipelined mode: the loss reaches the host (mapped pinned slot) from inside the step's last launch
        if self.p2p_grad:
            # every rank's gradients over NVLink; the optimizer launch adds the slots in rank order as it reads them
            # (bitwise identical parameters on all ranks, no reduction launch) and leaves the sum in flat_grad
            self.x_grad.allgather(self.flat_grad)
            self._mark("grad_exchange")
            check(lib.tt_adamw_step_p2p(_p(self.flat), _p(self.flat_grad), C.byref(self.x_grad.desc), _p(self.exp_avg),
                                        _p(self.exp_avg_sq), self.n_params, self.lr, self.betas[0], self.betas[1], self.eps,
                                        self.weight_decay, _p(self.step_count), _p(self.flat_bf16),
                                        _p(self.loss) if pub is not None else None, _p(pub), s), "tt_adamw_step_p2p")
        else:
            if self.world > 1:
                parallel.allreduce_sum_(self.flat_grad, self.group)
            self._mark("grad_exchange")
            if self.par_towers and self.embed_fused:
                # the second tower's table gradient joins here: grad[table] + table_grad2, written back to the table's .grad
                check(lib.tt_adamw_step_extra(_p(self.flat), _p(self.flat_grad), _p(self.table_grad2), self.offsets[id(self.table)],
                                              self.table_grad2.numel(), _p(self.exp_avg), _p(self.exp_avg_sq), self.n_params,
                                              self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                              _p(self.step_count), _p(self.flat_bf16), _p(self.loss) if pub is not None else None,
                                              _p(pub), s), "tt_adamw_step_extra")
            else:
                check(lib.tt_adamw_step_publish(_p(self.flat), _p(self.flat_grad), _p(self.exp_avg), _p(self.exp_avg_sq),
                                                self.n_params, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                                _p(self.step_count), _p(self.flat_bf16), _p(self.loss) if pub is not None else None,
                                                _p(pub), s), "tt_adamw_step")
        self._mark("adamw")

    def _towers_parallel(self, fn) -> None:
        """Both towers' launches of one phase (forward or backward) as two concurrent chains: tower 0 on the step's stream,
        tower 1 on the second tower stream, forked behind everything launched so far and joined before anything that follows
        (capturable: the fork / join become graph edges)."""
        cur = torch.cuda.current_stream()
        self._tower_stream.wait_stream(cur)
        fn(0)
        with torch.cuda.stream(self._tower_stream):
            fn(1, second=True)
        cur.wait_stream(self._tower_stream)

    def _local_loss_fwd(self, s):
        """In-batch loss forward against the local documents (one launch on the tensor-core path)."""
        lib, B, H = self.lib, self.B, self.H
        q, d = self.y[:B], self.y[B:2 * B]
        inv_t, scale = 1.0 / self.temperature, 1.0 / (B * self.world)
        qb = self.y_bf16[:B] if self.y_bf16 is not None else None
        db = self.y_bf16[B:2 * B] if self.y_bf16 is not None else None
        if self.local_fast:
            check(lib.tt_inbatch_ce_fwd_ex(_p(qb), B, _p(db), B, B, B, 0, 0, H, inv_t, 0, scale, _p(self.loss),
                                           _p(self.lse), _p(self.pos_mean), _p(self.ce_ws), self.ce_ws.numel(),
                                           _p(self.ce_sync), s), "tt_inbatch_ce_fwd_ex")
        else:
            check(lib.tt_inbatch_ce_fwd(_p(q), _p(d), _p(qb), _p(db), B, B, H, inv_t, 0, scale, _p(self.loss),
                                        _p(self.lse), _p(self.pos_mean), self.prec, _p(self.ws), self.ws.numel(), s),
                  "tt_inbatch_ce_fwd")

    def _local_loss_onepass(self, s):
        """Two loss launches: forward + query gradient in one pass over S, then the document gradient; both finish the
        normalise backward themselves (dz + column sums for the tower backward)."""
        lib, B, H = self.lib, self.B, self.H
        inv_t, scale = 1.0 / self.temperature, 1.0 / (B * self.world)
        qb, db = self.y_bf16[:B], self.y_bf16[B:2 * B]
        vp = lambda t: None if t is None else t.data_ptr()
        nb = B // 32
        qp = _lib.CePass(vp(qb), B, vp(db), B, B, B, 0, 0, None, 0, None, 0,
                         vp(self.dz_bf16[:B]), vp(self.dz_colsum[:nb]), vp(self.inv_norm[:B]))
        dp = _lib.CePass(vp(db), B, vp(qb), B, B, B, 0, 0, vp(self.lse), 0, None, 0,
                         vp(self.dz_bf16[B:2 * B]), vp(self.dz_colsum[nb:2 * nb]), vp(self.inv_norm[B:2 * B]))
        if self.onelaunch:
            # both launches as one kernel with a grid-wide barrier between the phases (square single-process case)
            rc = lib.tt_inbatch_ce_onepass(C.byref(qp), C.byref(dp), H, inv_t, inv_t, scale, None, _p(self.loss), _p(self.lse),
                                           _p(self.pos_mean), _p(self.onepass_sync), s)
            if rc == 0:
                return
            if rc != _lib.TT_ERR_UNSUPPORTED:
                check(rc, "tt_inbatch_ce_onepass")
            self.onelaunch = False                          # shapes / device do not allow it: two launches from now on
        if self.ce_stash is not None:
            check(lib.tt_inbatch_ce_fwd_dq_stash(C.byref(qp), H, inv_t, inv_t, scale, None, _p(self.loss), _p(self.lse),
                                                 _p(self.pos_mean), _p(self.onepass_sync), _p(self.ce_stash), s),
                  "tt_inbatch_ce_fwd_dq_stash")
            check(lib.tt_inbatch_ce_dd_stash(C.byref(dp), H, inv_t, scale, None, _p(self.ce_stash), s), "tt_inbatch_ce_dd_stash")
            return
        check(lib.tt_inbatch_ce_fwd_dq(C.byref(qp), H, inv_t, inv_t, scale, None, _p(self.loss), _p(self.lse),
                                       _p(self.pos_mean), _p(self.onepass_sync), s), "tt_inbatch_ce_fwd_dq")
        check(lib.tt_inbatch_ce_dd(C.byref(dp), H, inv_t, scale, None, s), "tt_inbatch_ce_dd")

    def _local_loss_bwd(self, s):
        """Both loss gradients in one launch: fused with the normalise backward (dz + column sums), or as per-split
        slices the tower backward sums, or (fp32 / unsupported shapes) as plain dq, dd."""
        lib, B, H = self.lib, self.B, self.H
        q, d = self.y[:B], self.y[B:2 * B]
        dq, dd = self.dy[:B], self.dy[B:2 * B]
        inv_t, scale = 1.0 / self.temperature, 1.0 / (B * self.world)
        qb = self.y_bf16[:B] if self.y_bf16 is not None else None
        db = self.y_bf16[B:2 * B] if self.y_bf16 is not None else None
        if self.ce_fused:
            vp = lambda t: None if t is None else t.data_ptr()
            nb = B // 32
            qp = _lib.CePass(vp(qb), B, vp(db), B, B, B, 0, 0, vp(self.lse), 0, None, 0,
                             vp(self.dz_bf16[:B]), vp(self.dz_colsum[:nb]), vp(self.inv_norm[:B]))
            dp = _lib.CePass(vp(db), B, vp(qb), B, B, B, 0, 0, vp(self.lse), 0, None, 0,
                             vp(self.dz_bf16[B:2 * B]), vp(self.dz_colsum[nb:2 * nb]), vp(self.inv_norm[B:2 * B]))
            check(lib.tt_inbatch_ce_bwd_parts_ex(C.byref(qp), C.byref(dp), H, inv_t, scale, None, self.dy_parts, s),
                  "tt_inbatch_ce_bwd_parts_ex")
        elif self.dy_parts > 1:
            check(lib.tt_inbatch_ce_bwd_parts(_p(qb), _p(db), _p(self.lse), B, B, H, inv_t, 0, scale, None,
                                              _p(dq), self.dy_part_stride, _p(dd), self.dy_part_stride, s),
                  "tt_inbatch_ce_bwd_parts")
        else:
            check(lib.tt_inbatch_ce_bwd(_p(q), _p(d), _p(qb), _p(db), _p(self.lse), B, B, H, inv_t, 0, scale,
                                        None, _p(dq), _p(dd), self.prec, _p(self.ws), self.ws.numel(), s),
                  "tt_inbatch_ce_bwd")

    def _global_inbatch_fast(self, s, inv_t):
        """Global in-batch negatives on the tensor-core path.  Peer-memory exchanges: D on the main stream (the forward
        needs nothing else), Q on a side stream underneath the forward kernel (only the backward's dD pass reads it), lse
        after the forward; both gradients come from ONE launch.  NCCL fallback: one all-gather of the [Q_r | D_r] blocks,
        which the loss kernels index in place (block-interleaved rows), and one of lse."""
        lib, B, H, W = self.lib, self.B, self.H, self.world
        Bg = B * W
        vp = lambda t: None if t is None else t.data_ptr()
        if self.p2p:
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)                    # tower outputs are ready
            self._side_open = True
            with torch.cuda.stream(self._side):
                self.x_q.allgather(self.y_bf16[:B])         # consumed by the backward's dD pass only: hidden under the forward
            self.x_d.allgather(self.y_bf16[B:2 * B])
            d_all, d_rows, d_blk, d_stride, d_off = self.dg_bf16, Bg, Bg, 0, 0          # one contiguous [Bg, H] matrix each
            q_all, q_off = self.qg_bf16, 0
        else:
            dist.all_gather_into_tensor(self.yg_bf16, self.y_bf16[:2 * B], group=self.group)
            d_all, d_rows, d_blk, d_stride, d_off = self.yg_bf16, W * 2 * B, B, 2 * B, B  # [Q_r | D_r] blocks, indexed in place
            q_all, q_off = self.yg_bf16, 0
        scale = 1.0 / Bg
        nb = B // 32
        self._mark("loss/gather_D")
        if self.onepass:                                    # forward + query gradient in one pass over S
            qp = _lib.CePass(vp(self.y_bf16[:B]), B, vp(d_all), Bg, d_rows, d_blk, d_stride, d_off, None, self.rank * B, None, 0,
                             vp(self.dz_bf16[:B]), vp(self.dz_colsum[:nb]), vp(self.inv_norm[:B]))
            if self.p2p and self.gated:
                # fused with the exchange: launched behind the all-gather kernel without waiting for it to retire; each
                # rank's block of D is consumed as its arrival counter reaches this round's target (own block first)
                check(lib.tt_inbatch_ce_fwd_dq_p2p(C.byref(qp), H, inv_t, inv_t, scale, None, _p(self.loss), _p(self.lse),
                                                   _p(self.pos_mean), _p(self.onepass_sync), C.byref(self.x_d.desc),
                                                   _p(self.y_bf16[B:2 * B]), s),
                      "tt_inbatch_ce_fwd_dq_p2p")
            else:
                check(lib.tt_inbatch_ce_fwd_dq(C.byref(qp), H, inv_t, inv_t, scale, None, _p(self.loss), _p(self.lse),
                                               _p(self.pos_mean), _p(self.onepass_sync), s), "tt_inbatch_ce_fwd_dq")
        else:
            check(lib.tt_inbatch_ce_fwd_ex(_p(self.y_bf16[:B]), B, _p(d_all), Bg, d_rows, d_blk, d_stride, d_off, H, inv_t,
                                           self.rank * B, scale, _p(self.loss), _p(self.lse), _p(self.pos_mean),
                                           _p(self.ce_ws), self.ce_ws.numel(), _p(self.ce_sync), s), "tt_inbatch_ce_fwd_ex")
        self._mark("loss/forward(+dQ)")
        if self.p2p:
            self.x_lse.allgather(self.lse)
            if self._lse_pack:
                self.lse_g.view(W, B).copy_(self.x_lse.gathered(torch.float32, (B,)))
            torch.cuda.current_stream().wait_stream(self._side)          # Q of every rank has arrived
            self._side_open = False
        else:
            dist.all_gather_into_tensor(self.lse_g, self.lse, group=self.group)
        self._mark("loss/gather_lse(+Q)")
        if self.onepass:
            dp = _lib.CePass(vp(self.y_bf16[B:2 * B]), B, vp(q_all), Bg, d_rows, d_blk, d_stride, q_off, vp(self.lse_g),
                             -self.rank * B, None, 0,
                             vp(self.dz_bf16[B:2 * B]), vp(self.dz_colsum[nb:2 * nb]), vp(self.inv_norm[B:2 * B]))
            check(lib.tt_inbatch_ce_dd(C.byref(dp), H, inv_t, scale, None, s), "tt_inbatch_ce_dd")
            return
        fz = (lambda a, b, c: (vp(a), vp(b), vp(c))) if self.ce_fused else (lambda a, b, c: (None, None, None))
        qp = _lib.CePass(vp(self.y_bf16[:B]), B, vp(d_all), Bg, d_rows, d_blk, d_stride, d_off, vp(self.lse), self.rank * B,
                         vp(self.dy[:B]), self.dy_part_stride,
                         *(fz(self.dz_bf16[:B], self.dz_colsum[:nb], self.inv_norm[:B]) if self.ce_fused else (None, None, None)))
        dp = _lib.CePass(vp(self.y_bf16[B:2 * B]), B, vp(q_all), Bg, d_rows, d_blk, d_stride, q_off, vp(self.lse_g),
                         -self.rank * B, vp(self.dy[B:2 * B]), self.dy_part_stride,
                         *(fz(self.dz_bf16[B:2 * B], self.dz_colsum[nb:2 * nb], self.inv_norm[B:2 * B]) if self.ce_fused else (None, None, None)))
        check(lib.tt_inbatch_ce_bwd_parts_ex(C.byref(qp), C.byref(dp), H, inv_t, scale, None, self.dy_parts, s),
              "tt_inbatch_ce_bwd_parts_ex")

    # ---------------------------------------------------------------------------------------
    def load_batch(self, q_ids: torch.Tensor, d_ids: torch.Tensor, n_ids: Optional[torch.Tensor] = None):
        """Copy a batch of token ids (host or device, [B,L]) into the static input buffer."""
        B = self.B
        parts = [q_ids, d_ids] + ([n_ids] if self.passes == 3 else [])
        if self.passes == 3 and n_ids is None:
            raise ValueError("triplet loss needs negative ids")
        for i, t in enumerate(parts):
            if tuple(t.shape) != (B, self.L):
                raise ValueError(f"expected ids of shape {(B, self.L)}, got {tuple(t.shape)}")
            self.ids[i * B:(i + 1) * B].copy_(t, non_blocking=True)

    # ---- input pipelining: the next batch's host->device copy overlaps the current step ------------------
    def prefetch(self, q_ids: torch.Tensor, d_ids: torch.Tensor, n_ids: Optional[torch.Tensor] = None) -> None:
        """Start copying the NEXT batch (pinned host or device tensors, [B,L]) on a separate copy stream, straight into
        the static id buffer the next `step()` (no arguments) will read: two buffers alternate, each with its own captured
        graph, so the step's own stream carries no staging copy.  One batch may be in flight."""
        if self._ids_bufs[1] is None:
            self._ids_bufs[1] = torch.zeros_like(self.ids)
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._copy_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._buf_free = [torch.cuda.Event(), torch.cuda.Event()]
            for ev in self._buf_free:
                ev.record(torch.cuda.current_stream())
            self._pf_idx = 1                                # buffer 0 doubles as the explicit-ids buffer: start with 1
        B = self.B
        parts = [q_ids, d_ids] + ([n_ids] if self.passes == 3 else [])
        if self.passes == 3 and n_ids is None:
            raise ValueError("triplet loss needs negative ids")
        if self._staged is not False:
            raise RuntimeError("prefetch(): the previous batch has not been consumed by step() yet")
        t = self._pf_idx
        buf = self._ids_bufs[t]
        cs = self._copy_stream
        cs.wait_event(self._buf_free[t])                    # the last replay that read this buffer has finished
        with torch.cuda.stream(cs):
            for i, x in enumerate(parts):
                if tuple(x.shape) != (B, self.L):
                    raise ValueError(f"expected ids of shape {(B, self.L)}, got {tuple(x.shape)}")
                buf[i * B:(i + 1) * B].copy_(x, non_blocking=True)
        self._copy_done[t].record(cs)
        self._staged = t
        self._pf_idx = t ^ 1

    def _consume_prefetched(self) -> int:
        if self._staged is False:
            raise RuntimeError("step() without arguments needs a prefetch() first")
        t = self._staged
        torch.cuda.current_stream().wait_event(self._copy_done[t])
        self._staged = False
        return t

    def run(self, gi: int = 0) -> torch.Tensor:
        """One optimizer step on the batch currently in static id buffer `gi` (0: the buffer load_batch() fills); returns
        the loss (device scalar)."""
        self._ids_cur = self._ids_bufs[gi]
        self._loss_dst = self._loss_ring_g[gi]
        if not self.use_graph:
            self._step_impl()
        elif self._graphs[gi] is None:
            if self._graphs[0] is None and self._graphs[1] is None:
                # warm-up outside capture (module loading, workspace sizing, NCCL communicators) ...
                st = self._snapshot()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._step_impl()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                # ... restore the state the warm-up step changed, then capture and replay once
                self._restore(st)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_impl()
            self._graphs[gi] = g
            self.graph = self._graphs[0] if self._graphs[0] is not None else g
            g.replay()
        else:
            self._graphs[gi].replay()
        self._ids_cur = self.ids
        self._loss_dst = None
        self.steps_done += 1
        return self.loss

    def step(self, q_ids=None, d_ids=None, n_ids=None) -> torch.Tensor:
        """One optimizer step.  With ids: copy them in and run.  Without: consume the batch started by prefetch()."""
        if q_ids is None:
            t = self._consume_prefetched()
            out = self.run(t)
            self._buf_free[t].record(torch.cuda.current_stream())
            self._last_gi = t
            return out
        self.load_batch(q_ids, d_ids, n_ids)
        self._last_gi = -1
        return self.run()

    def read_loss_async(self):
        """Queue a device->host copy of the loss of the step just issued into a pinned two-slot ring and return a
        zero-argument callable that waits for THAT copy and returns the Python float.  Lets a training loop read every
        step's loss (as the reference does with loss.item()) one step late, so the host never stalls the GPU:

            pending = None
            for batch in loader:
                trainer.step(*batch); nxt = trainer.read_loss_async()
                if pending is not None: log(pending())
                pending = nxt
        """
        if self._last_gi >= 0:
            # pipelined step: its replay already carried the copy into the slot of its graph; that slot is written again
            # two steps later, after the loop above has consumed it
            k = self._last_gi
            ring = self._loss_ring_g
        else:
            k = self._loss_slot
            self._loss_slot ^= 1
            ring = self._loss_ring
            ring[k].copy_(self.loss.reshape(1), non_blocking=True)      # stream-ordered before the next step overwrites it
        self._loss_ev[k].record(torch.cuda.current_stream())
        def wait(k=k, ring=ring):
            self._loss_ev[k].synchronize()
            _lib.raise_on_bad_ids()                                     # the step's gather kernels have completed
            return float(ring[k])
        return wait

    def kernels_per_step(self) -> int:
        """Launches of libtt_b200 kernels in one step, measured on an eager step whose effect on the parameters, the
        Adam moments and the step counter is undone.  Collective: with a process group every rank must call it."""
        st = self._snapshot()
        before = _lib.launch_count()
        self._step_impl()
        n = _lib.launch_count() - before
        torch.cuda.current_stream().synchronize()
        self._restore(st)
        return n

    # ---- state: external weight loads, checkpoint / resume (twotower/utils.py save_checkpoint stores
    # model.state_dict() and optimizer.state_dict(); train.py:359 torch.optim.AdamW) ---------------------------
    def _snapshot(self):
        return {k: v.clone() for k, v in (("flat", self.flat), ("m", self.exp_avg), ("v", self.exp_avg_sq),
                                          ("t", self.step_count))}

    def _restore(self, st) -> None:
        self.flat.copy_(st["flat"]); self.exp_avg.copy_(st["m"]); self.exp_avg_sq.copy_(st["v"])
        self.step_count.copy_(st["t"])
        self.sync_from_model()

    def sync_from_model(self) -> None:
        """Call after anything OUTSIDE the trainer wrote the parameters (``model.load_state_dict(...)``, manual
        ``p.data`` edits): refreshes the bf16 weight shadow the tensor-core kernels read.  The fp32 parameters
        themselves are views of the flat buffer, so in-place loads land there directly."""
        for p in self._params:
            off = self.offsets[id(p)]
            if p.data.data_ptr() != self.flat.data_ptr() + off * 4:          # someone re-pointed p.data: adopt the values
                self.flat[off:off + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + p.numel()].view(p.shape)
        if self.flat_bf16 is not None:
            ops.cast_bf16(self.flat, self.flat_bf16)

    def state_dict(self) -> Dict:
        """torch.optim.AdamW layout: {'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]} with the
        parameters numbered in ``model.parameters()`` order (trainable ones), so it loads into a reference optimizer."""
        step = float(self.step_count[0].item())
        state = {}
        for i, p in enumerate(self._params):
            off, n = self.offsets[id(p)], p.numel()
            state[i] = {"step": torch.tensor(step), "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": True, "params": list(range(len(self._params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: Dict) -> None:
        """Resume from ``state_dict()`` or from a reference ``torch.optim.AdamW.state_dict()`` of the same model."""
        ids = sd["param_groups"][0]["params"]
        if len(ids) != len(self._params):
            raise ValueError(f"optimizer state has {len(ids)} parameters, the model has {len(self._params)}")
        steps = set()
        for i, p in zip(ids, self._params):
            st = sd["state"].get(i)
            off, n = self.offsets[id(p)], p.numel()
            if st is None:
                self.exp_avg[off:off + n].zero_(); self.exp_avg_sq[off:off + n].zero_()
                continue
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ; the fused AdamW keeps one counter")
        g = sd["param_groups"][0]
        hp = (float(g["lr"]), tuple(g["betas"]), float(g["eps"]), float(g["weight_decay"]))
        if self.graph is not None and hp != (self.lr, tuple(self.betas), self.eps, self.weight_decay):
            raise ValueError("hyper-parameters are baked into the recorded CUDA graph; create the trainer with the checkpoint's values")
        self.lr, self.betas, self.eps, self.weight_decay = hp
        self.step_count.zero_()
        self.step_count[0] = steps.pop() if steps else 0
        self.sync_from_model()

    def check(self) -> None:
        """Synchronise and surface asynchronous failures: a token id outside the table (IndexError, as nn.Embedding raises)
        or a peer-memory exchange that timed out (RuntimeError)."""
        torch.cuda.current_stream().synchronize()
        _lib.raise_on_bad_ids()
        for x in (getattr(self, "x_grad", None), getattr(self, "x_d", None), getattr(self, "x_q", None), getattr(self, "x_lse", None)):
            if x is not None:
                x.check()
