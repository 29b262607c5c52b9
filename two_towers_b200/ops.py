"""Thin tensor-level wrappers over the C ABI (``include/tt_b200.h``) + autograd Functions.

PyTorch is used here for device memory (``torch.empty``), the current CUDA stream and the
autograd tape only -- every arithmetic operation on the hot path is a hand-written sm_100a
kernel inside ``libtt_b200.so``.  CPU tensors are rejected: there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import TT_PREC_BF16, TT_PREC_FP32, TT_TOPK_MAX, check, raise_on_bad_ids  # noqa: F401

_PRECISION = {"fp32": TT_PREC_FP32, "bf16": TT_PREC_BF16, TT_PREC_FP32: TT_PREC_FP32, TT_PREC_BF16: TT_PREC_BF16}
_default_precision = TT_PREC_FP32


def set_default_precision(p) -> None:
    """'fp32' (CUDA-core parity mode, rel 1e-5) or 'bf16' (tcgen05 performance mode, rel 2e-2)."""
    global _default_precision
    _default_precision = _PRECISION[p]


def get_default_precision() -> int:
    return _default_precision


def resolve_precision(p) -> int:
    return _default_precision if p is None else _PRECISION[p]


# --------------------------------------------------------------------------------------
# plumbing
# --------------------------------------------------------------------------------------
def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "two_towers_b200: tensor is on %s -- the hot path runs only on a B200 (sm_100a); "
                "there is no CPU fallback" % t.device)


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32 tensor, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ids(t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    if t.dtype == torch.int64:
        b = 8
    elif t.dtype == torch.int32:
        b = 4
    else:
        raise TypeError(f"token ids must be int64 or int32, got {t.dtype}")
    return (t if t.is_contiguous() else t.contiguous()), b


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _lib_():
    return _lib.load()


# --------------------------------------------------------------------------------------
# K1 / K2
# --------------------------------------------------------------------------------------
def embed_gather(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    _need_cuda(ids, table)
    ids_c, idb = _ids(ids)
    table = _f32(table)
    V, E = table.shape
    out = torch.empty(*ids.shape, E, dtype=torch.float32, device=table.device)
    check(_lib_().tt_embed_gather(_p(ids_c), idb, _p(table), ids_c.numel(), V, E, _p(out), _stream()),
          "tt_embed_gather")
    return out


def embed_pool_fwd(ids: torch.Tensor, table: torch.Tensor, want_bf16: bool = False):
    """ids [R,L] -> (pooled [R,E] fp32, inv_len [R], pooled_bf16 | None)."""
    _need_cuda(ids, table)
    if ids.dim() != 2:
        raise ValueError("ids must be [rows, L]")
    ids_c, idb = _ids(ids)
    table = _f32(table)
    V, E = table.shape
    R, L = ids_c.shape
    pooled = torch.empty(R, E, dtype=torch.float32, device=table.device)
    inv_len = torch.empty(R, dtype=torch.float32, device=table.device)
    pooled_bf16 = torch.empty(R, E, dtype=torch.bfloat16, device=table.device) if want_bf16 else None
    check(_lib_().tt_embed_pool_fwd(_p(ids_c), idb, _p(table), R, L, V, E, _p(pooled), _p(inv_len),
                                    _p(pooled_bf16), None, _stream()), "tt_embed_pool_fwd")
    return pooled, inv_len, pooled_bf16


def embed_pool_bwd(ids: torch.Tensor, inv_len: torch.Tensor, d_pooled: torch.Tensor, V: int) -> torch.Tensor:
    _need_cuda(ids, inv_len, d_pooled)
    ids_c, idb = _ids(ids)
    d_pooled = _f32(d_pooled)
    R, L = ids_c.shape
    E = d_pooled.shape[1]
    d_table = torch.empty(V, E, dtype=torch.float32, device=d_pooled.device)
    nb = _lib_().tt_embed_pool_bwd_workspace(R, L, V, E)
    ws = _workspace(nb, d_pooled.device)
    check(_lib_().tt_embed_pool_bwd(_p(ids_c), idb, _p(inv_len), _p(d_pooled), R, L, V, E, _p(d_table),
                                    _p(ws), ws.numel(), _stream()), "tt_embed_pool_bwd")
    return d_table


# --------------------------------------------------------------------------------------
# K3 / K3'
# --------------------------------------------------------------------------------------
def mlp_fwd(x, w1, b1, w2, b2, precision=None, want_bf16: bool = False):
    """-> (y [R,H], h1, z, y_bf16|None)."""
    _need_cuda(x, w1, b1, w2, b2)
    prec = resolve_precision(precision)
    x, w1, b1, w2, b2 = map(_f32, (x, w1, b1, w2, b2))
    R, E = x.shape
    H = w1.shape[0]
    dev = x.device
    h1 = torch.empty(R, H, dtype=torch.float32, device=dev)
    z = torch.empty(R, H, dtype=torch.float32, device=dev)
    y = torch.empty(R, H, dtype=torch.float32, device=dev)
    y_bf16 = torch.empty(R, H, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    ws = _workspace(_lib_().tt_mlp_workspace(R, E, H, prec), dev)
    check(_lib_().tt_mlp_fwd(_p(x), _p(w1), _p(b1), _p(w2), _p(b2), R, E, H, _p(h1), _p(z), _p(y), _p(y_bf16),
                             None, None, None, None, None, None, prec, _p(ws), ws.numel(), _stream()), "tt_mlp_fwd")
    return y, h1, z, y_bf16


def mlp_bwd(dy, x, w1, w2, h1, z, need_dx: bool = True, precision=None):
    """-> (dx|None, dw1, db1, dw2, db2)."""
    _need_cuda(dy, x, w1, w2, h1, z)
    prec = resolve_precision(precision)
    dy, x, w1, w2, h1, z = map(_f32, (dy, x, w1, w2, h1, z))
    R, E = x.shape
    H = w1.shape[0]
    dev = x.device
    dx = torch.empty(R, E, dtype=torch.float32, device=dev) if need_dx else None
    dw1 = torch.empty(H, E, dtype=torch.float32, device=dev)
    db1 = torch.empty(H, dtype=torch.float32, device=dev)
    dw2 = torch.empty(H, H, dtype=torch.float32, device=dev)
    db2 = torch.empty(H, dtype=torch.float32, device=dev)
    ws = _workspace(_lib_().tt_mlp_workspace(R, E, H, prec), dev)
    check(_lib_().tt_mlp_bwd(_p(dy), _p(x), _p(w1), _p(w2), _p(h1), _p(z), R, E, H, _p(dx), _p(dw1), _p(db1),
                             _p(dw2), _p(db2), None, None, None, None, 1, 0, None, None, None, None, None, prec, _p(ws), ws.numel(),
                             _stream()),
          "tt_mlp_bwd")
    return dx, dw1, db1, dw2, db2


def dropout_keep_mask(seed: int, rows: int, cols: int, p: float, seed_step: Optional[int] = None):
    """Host restatement (numpy uint64) of the counter-based keep-mask of tt_proj_ln_fwd/bwd (include/tt_b200.h): used by
    the parity tests to apply the SAME mask to the CPU restatement of the reference.  -> bool [rows, cols]."""
    import numpy as np
    m64 = (1 << 64) - 1
    if seed_step is not None:
        seed = (seed + (seed_step + 1) * 0xD1B54A32D192ED03) & m64
    with np.errstate(over="ignore"):
        idx = np.arange(rows * cols, dtype=np.uint64)
        x = np.uint64(seed & m64) + idx * np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    u = (x >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (u >= np.float32(p)).reshape(rows, cols)


def proj_ln_fwd(x, w, b, gamma, beta, has_projection: bool, dropout_p: float, training: bool, seed: int,
                precision=None, seed_step: Optional[torch.Tensor] = None):
    """-> (y, a, stats, z); a/stats/z are None without projection."""
    _need_cuda(x, w, b, gamma, beta)
    prec = resolve_precision(precision)
    x = _f32(x)
    R, E = x.shape
    dev = x.device
    if not has_projection:
        y = torch.empty(R, E, dtype=torch.float32, device=dev)
        check(_lib_().tt_proj_ln_fwd(_p(x), None, None, None, None, R, E, E, 0, 0.0, 0, 0, None, None, None, None,
                                     _p(y), prec, None, 0, _stream()), "tt_proj_ln_fwd")
        return y, None, None, None
    w, b, gamma, beta = map(_f32, (w, b, gamma, beta))
    H = w.shape[0]
    a = torch.empty(R, H, dtype=torch.float32, device=dev)
    stats = torch.empty(R, 2, dtype=torch.float32, device=dev)
    z = torch.empty(R, H, dtype=torch.float32, device=dev)
    y = torch.empty(R, H, dtype=torch.float32, device=dev)
    ws = _workspace(_lib_().tt_proj_ln_workspace(R, E, H, prec), dev)
    check(_lib_().tt_proj_ln_fwd(_p(x), _p(w), _p(b), _p(gamma), _p(beta), R, E, H, 1, float(dropout_p),
                                 int(bool(training)), int(seed) & (2 ** 64 - 1), _p(seed_step), _p(a), _p(stats), _p(z),
                                 _p(y), prec, _p(ws), ws.numel(), _stream()), "tt_proj_ln_fwd")
    return y, a, stats, z


def proj_ln_bwd(dy, x, w, gamma, a, stats, z, has_projection: bool, dropout_p: float, training: bool, seed: int,
                need_dx: bool = True, precision=None, seed_step: Optional[torch.Tensor] = None):
    """-> (dx|None, dw, db, dgamma, dbeta)."""
    _need_cuda(dy, x)
    prec = resolve_precision(precision)
    dy, x = _f32(dy), _f32(x)
    R, E = x.shape
    dev = x.device
    if not has_projection:
        dx = torch.empty(R, E, dtype=torch.float32, device=dev)
        check(_lib_().tt_proj_ln_bwd(_p(dy), _p(x), None, None, None, None, None, R, E, E, 0, 0.0, 0, 0, None, _p(dx),
                                     None, None, None, None, prec, None, 0, _stream()), "tt_proj_ln_bwd")
        return dx, None, None, None, None
    H = w.shape[0]
    dx = torch.empty(R, E, dtype=torch.float32, device=dev) if need_dx else None
    dw = torch.empty(H, E, dtype=torch.float32, device=dev)
    db = torch.empty(H, dtype=torch.float32, device=dev)
    dgamma = torch.empty(H, dtype=torch.float32, device=dev)
    dbeta = torch.empty(H, dtype=torch.float32, device=dev)
    ws = _workspace(_lib_().tt_proj_ln_workspace(R, E, H, prec), dev)
    check(_lib_().tt_proj_ln_bwd(_p(dy), _p(x), _p(_f32(w)), _p(_f32(gamma)), _p(a), _p(stats), _p(z), R, E, H, 1,
                                 float(dropout_p), int(bool(training)), int(seed) & (2 ** 64 - 1), _p(seed_step), _p(dx),
                                 _p(dw), _p(db), _p(dgamma), _p(dbeta), prec, _p(ws), ws.numel(), _stream()),
          "tt_proj_ln_bwd")
    return dx, dw, db, dgamma, dbeta


# --------------------------------------------------------------------------------------
# K4
# --------------------------------------------------------------------------------------
def inbatch_ce_fwd(q, d, temperature: float, label_offset: int = 0, loss_scale: Optional[float] = None,
                   precision=None, q_bf16=None, d_bf16=None, want_pos_mean: bool = False):
    """-> (loss scalar tensor, lse [Bq], pos_mean | None)."""
    _need_cuda(q, d)
    prec = resolve_precision(precision)
    q, d = _f32(q), _f32(d)
    Bq, H = q.shape
    Bd = d.shape[0]
    dev = q.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    lse = torch.empty(Bq, dtype=torch.float32, device=dev)
    pos_mean = torch.empty((), dtype=torch.float32, device=dev) if want_pos_mean else None
    scale = 1.0 / Bq if loss_scale is None else float(loss_scale)
    ws = _workspace(_lib_().tt_inbatch_ce_workspace(Bq, Bd, H, prec), dev)
    check(_lib_().tt_inbatch_ce_fwd(_p(q), _p(d), _p(q_bf16), _p(d_bf16), Bq, Bd, H, 1.0 / float(temperature),
                                    int(label_offset), scale, _p(loss), _p(lse), _p(pos_mean), prec, _p(ws),
                                    ws.numel(), _stream()), "tt_inbatch_ce_fwd")
    return loss, lse, pos_mean


def inbatch_ce_bwd(q, d, lse, temperature: float, label_offset: int = 0, loss_scale: Optional[float] = None,
                   grad_out: Optional[torch.Tensor] = None, need_dq: bool = True, need_dd: bool = True,
                   precision=None, q_bf16=None, d_bf16=None):
    _need_cuda(q, d, lse, grad_out)
    prec = resolve_precision(precision)
    q, d = _f32(q), _f32(d)
    Bq, H = q.shape
    Bd = d.shape[0]
    dev = q.device
    dq = torch.empty_like(q) if need_dq else None
    dd = torch.empty_like(d) if need_dd else None
    scale = 1.0 / Bq if loss_scale is None else float(loss_scale)
    if grad_out is not None:
        grad_out = _f32(grad_out)
    ws = _workspace(_lib_().tt_inbatch_ce_workspace(Bq, Bd, H, prec), dev)
    check(_lib_().tt_inbatch_ce_bwd(_p(q), _p(d), _p(q_bf16), _p(d_bf16), _p(lse), Bq, Bd, H,
                                    1.0 / float(temperature), int(label_offset), scale, _p(grad_out), _p(dq),
                                    _p(dd), prec, _p(ws), ws.numel(), _stream()), "tt_inbatch_ce_bwd")
    return dq, dd


_ONEPASS_SYNC = {}


def inbatch_ce_onepass(q_bf16: torch.Tensor, d_bf16: torch.Tensor, temperature: float, label_offset: int = 0,
                       loss_scale: Optional[float] = None, grad_out: Optional[torch.Tensor] = None,
                       logit_bound: Optional[float] = None, stash: Optional[bool] = None):
    """Loss forward and both gradients of the in-batch softmax (twotower/losses.py:107-116) in TWO launches on unit-norm
    bf16 rows: ``tt_inbatch_ce_fwd_dq`` (forward + dq, S formed once, fixed softmax shift) and ``tt_inbatch_ce_dd``.
    ``stash`` (None: whenever ``tt_inbatch_ce_stash_ok``): the first launch also stores its E tiles and the second is the plain
    product of ``tt_inbatch_ce_dd_stash`` -- nothing is recomputed.  Returns (loss, lse, pos_mean, dq, dd) with fp32 gradients."""
    _need_cuda(q_bf16, d_bf16, grad_out)
    assert q_bf16.dtype == torch.bfloat16 and d_bf16.dtype == torch.bfloat16
    lib = _lib_()
    Bq, H = q_bf16.shape
    Bd = d_bf16.shape[0]
    dev = q_bf16.device
    inv_t = 1.0 / float(temperature)
    bound = inv_t if logit_bound is None else float(logit_bound)
    if not lib.tt_inbatch_ce_onepass_ok(Bq, Bd, H, bound):
        raise RuntimeError(f"inbatch_ce_onepass: unsupported (H={H}, logit bound {bound})")
    scale = 1.0 / Bq if loss_scale is None else float(loss_scale)
    key = (dev.index, Bq)
    if key not in _ONEPASS_SYNC:
        _ONEPASS_SYNC[key] = torch.zeros(int(lib.tt_inbatch_ce_onepass_sync_bytes(Bq)), dtype=torch.uint8, device=dev)
    sync = _ONEPASS_SYNC[key]
    f32 = dict(dtype=torch.float32, device=dev)
    loss, lse, pm = torch.empty((), **f32), torch.empty(Bq, **f32), torch.empty((), **f32)
    dq = torch.empty(Bq, H, **f32)
    if grad_out is not None:
        grad_out = _f32(grad_out)
    qp = _lib.CePass(q_bf16.data_ptr(), Bq, d_bf16.data_ptr(), Bd, Bd, Bd, 0, 0, None, int(label_offset), dq.data_ptr(), 0,
                     None, None, None)
    stash_ok = bool(lib.tt_inbatch_ce_stash_ok(Bq, Bd, H))
    if stash and not stash_ok:
        raise RuntimeError(f"inbatch_ce_onepass: the stored-E form is not available for Bq={Bq} Bd={Bd} H={H}")
    if stash or (stash is None and stash_ok):
        buf = _workspace(lib.tt_inbatch_ce_stash_bytes(Bq, Bd, H), dev)
        check(lib.tt_inbatch_ce_fwd_dq_stash(C.byref(qp), H, inv_t, bound, scale, _p(grad_out), _p(loss), _p(lse), _p(pm),
                                             _p(sync), _p(buf), _stream()), "tt_inbatch_ce_fwd_dq_stash")
        dd = torch.empty(Bd, H, **f32)
        dp = _lib.CePass(d_bf16.data_ptr(), Bd, q_bf16.data_ptr(), Bq, Bq, Bq, 0, 0, None, int(label_offset),
                         dd.data_ptr(), 0, None, None, None)
        check(lib.tt_inbatch_ce_dd_stash(C.byref(dp), H, inv_t, scale, _p(grad_out), _p(buf), _stream()), "tt_inbatch_ce_dd_stash")
        return loss, lse, pm, dq, dd
    check(lib.tt_inbatch_ce_fwd_dq(C.byref(qp), H, inv_t, bound, scale, _p(grad_out), _p(loss), _p(lse), _p(pm), _p(sync),
                                   _stream()), "tt_inbatch_ce_fwd_dq")
    n = int(lib.tt_inbatch_ce_dd_nparts(Bd, Bq, H))
    parts = torch.empty(n, Bd, H, **f32)
    # document j is the positive of query j - label_offset (tt_ce_pass_t: "positive at row == col + off")
    dp = _lib.CePass(d_bf16.data_ptr(), Bd, q_bf16.data_ptr(), Bq, Bq, Bq, 0, 0, lse.data_ptr(), int(label_offset),
                     parts.data_ptr(), Bd * H, None, None, None)
    check(lib.tt_inbatch_ce_dd(C.byref(dp), H, inv_t, scale, _p(grad_out), _stream()), "tt_inbatch_ce_dd")
    return loss, lse, pm, dq, parts.sum(0) if n > 1 else parts[0]


# --------------------------------------------------------------------------------------
# K5 / K6
# --------------------------------------------------------------------------------------
def triplet_fwd(q, p, n, margin: float, want_sims: bool = False):
    _need_cuda(q, p, n)
    q, p, n = map(_f32, (q, p, n))
    B, H = q.shape
    dev = q.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    sims = torch.empty(3 * B, dtype=torch.float32, device=dev)
    pm = torch.empty((), dtype=torch.float32, device=dev) if want_sims else None
    nm = torch.empty((), dtype=torch.float32, device=dev) if want_sims else None
    check(_lib_().tt_triplet_fwd(_p(q), _p(p), _p(n), B, H, float(margin), _p(loss), _p(sims), _p(pm), _p(nm),
                                 _stream()), "tt_triplet_fwd")
    return loss, sims, pm, nm


def triplet_bwd(q, p, n, sims, margin: float, grad_out=None):
    _need_cuda(q, p, n, sims)
    q, p, n = map(_f32, (q, p, n))
    B, H = q.shape
    dq, dp, dn = torch.empty_like(q), torch.empty_like(p), torch.empty_like(n)
    if grad_out is not None:
        grad_out = _f32(grad_out)
    check(_lib_().tt_triplet_bwd(_p(q), _p(p), _p(n), _p(sims), B, H, float(margin), _p(grad_out), _p(dq), _p(dp),
                                 _p(dn), _stream()), "tt_triplet_bwd")
    return dq, dp, dn


def multineg_fwd(q, p, negs, temperature: float):
    _need_cuda(q, p, negs)
    q, p, negs = map(_f32, (q, p, negs))
    B, N, H = negs.shape
    dev = q.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    probs = torch.empty(B * (N + 1) + B, dtype=torch.float32, device=dev)
    check(_lib_().tt_multineg_fwd(_p(q), _p(p), _p(negs), B, N, H, 1.0 / float(temperature), _p(loss), _p(probs),
                                  _stream()), "tt_multineg_fwd")
    return loss, probs


def multineg_bwd(q, p, negs, probs, temperature: float, grad_out=None):
    _need_cuda(q, p, negs, probs)
    q, p, negs = map(_f32, (q, p, negs))
    B, N, H = negs.shape
    dq, dp, dnegs = torch.empty_like(q), torch.empty_like(p), torch.empty_like(negs)
    if grad_out is not None:
        grad_out = _f32(grad_out)
    check(_lib_().tt_multineg_bwd(_p(q), _p(p), _p(negs), _p(probs), B, N, H, 1.0 / float(temperature),
                                  _p(grad_out), _p(dq), _p(dp), _p(dnegs), _stream()), "tt_multineg_bwd")
    return dq, dp, dnegs


# --------------------------------------------------------------------------------------
# K7
# --------------------------------------------------------------------------------------
def topk_scan(index: torch.Tensor, queries: torch.Tensor, k: int, cosine: bool = True, id_offset: int = 0,
              workspace: Optional[torch.Tensor] = None, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """index [N,H] fp32|bf16, queries [nq,H] fp32 -> (scores [nq,k] fp32, ids [nq,k] int64), sorted
    descending, ties -> lower id.  k is clamped by the caller (k <= N)."""
    _need_cuda(index, queries)
    if index.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("index must be float32 or bfloat16")
    if not index.is_contiguous():
        raise ValueError("index must be contiguous [N,H]")
    queries = _f32(queries)
    N, H = index.shape
    nq = queries.shape[0]
    dev = index.device
    if out is None:
        scores = torch.empty(nq, k, dtype=torch.float32, device=dev)
        ids = torch.empty(nq, k, dtype=torch.int64, device=dev)
    else:
        scores, ids = out
    if workspace is None:
        workspace = _workspace(_lib_().tt_topk_scan_workspace(N, H, nq, k), dev)
    check(_lib_().tt_topk_scan(_p(index), int(index.dtype == torch.bfloat16), _p(queries), N, H, nq, int(k),
                               int(bool(cosine)), int(id_offset), _p(scores), _p(ids), _p(workspace),
                               workspace.numel(), _stream()), "tt_topk_scan")
    return scores, ids


def topk_scan_batched_ok(index: torch.Tensor, k: int) -> bool:
    """True when the batched tensor-core scan supports this index (bf16, H % 64 == 0, H <= 256) and k (<= 128)."""
    return bool(index.dtype == torch.bfloat16 and index.dim() == 2 and _lib_().tt_topk_scan_batched_ok(index.shape[1], int(k)))


def index_row_inv_norms(index: torch.Tensor) -> torch.Tensor:
    """1 / max(|d_n|, 1e-8) for every row of a bf16 index (cosine scores in topk_scan_batched); compute once per index."""
    _need_cuda(index)
    out = torch.empty(index.shape[0], dtype=torch.float32, device=index.device)
    check(_lib_().tt_index_row_inv_norms(_p(index), index.shape[0], index.shape[1], _p(out), _stream()), "tt_index_row_inv_norms")
    return out


def topk_scan_batched(index: torch.Tensor, queries: torch.Tensor, k: int, id_offset: int = 0,
                      row_inv_norms: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
    """Batched scan on the tcgen05 tensor cores: bf16 index [N,H], fp32 queries [nq,H] -> (scores [nq,k], ids [nq,k]),
    the index is read once per 128 queries.  row_inv_norms given: cosine scores, else raw dot products."""
    _need_cuda(index, queries, row_inv_norms)
    if index.dtype != torch.bfloat16 or not index.is_contiguous():
        raise TypeError("topk_scan_batched needs a contiguous bfloat16 index")
    queries = _f32(queries)
    N, H = index.shape
    nq = queries.shape[0]
    scores = torch.empty(nq, k, dtype=torch.float32, device=index.device)
    ids = torch.empty(nq, k, dtype=torch.int64, device=index.device)
    if workspace is None:
        workspace = _workspace(_lib_().tt_topk_scan_batched_workspace(N, H, nq), index.device)
    check(_lib_().tt_topk_scan_batched(_p(index), _p(queries), N, H, nq, int(k), _p(row_inv_norms), int(id_offset), _p(scores),
                                       _p(ids), _p(workspace), workspace.numel(), _stream()), "tt_topk_scan_batched")
    return scores, ids


def topk_scan_workspace_bytes(N: int, H: int, nq: int, k: int) -> int:
    return int(_lib_().tt_topk_scan_workspace(N, H, nq, k))


def topk_scan_p2p_ok(xch, nq: int, k: int) -> bool:
    """True when `xch` (parallel.P2PExchange) can carry the fused sharded search for nq queries / top-k."""
    return bool(_lib_().tt_topk_scan_p2p_ok(C.byref(xch.desc), int(nq), int(k)))


def topk_scan_p2p(index: torch.Tensor, queries: torch.Tensor, k: int, xch, cosine: bool = True, id_offset: int = 0,
                  workspace: Optional[torch.Tensor] = None):
    """Row-sharded exact top-k in two launches (tt_topk_scan_p2p): this rank's scan, then one kernel that merges the block
    lists, exchanges the shard's k candidates with every rank over NVLink peer memory and selects the global top-k.
    Collective over xch's group; returns (scores [nq,k], GLOBAL ids [nq,k]), identical on every rank."""
    _need_cuda(index, queries)
    if index.dtype not in (torch.float32, torch.bfloat16) or not index.is_contiguous():
        raise TypeError("index must be contiguous float32 or bfloat16 [N,H]")
    queries = _f32(queries)
    N, H = index.shape
    nq = queries.shape[0]
    dev = index.device
    scores = torch.empty(nq, k, dtype=torch.float32, device=dev)
    ids = torch.empty(nq, k, dtype=torch.int64, device=dev)
    if workspace is None:
        workspace = _workspace(_lib_().tt_topk_scan_workspace(N, H, nq, k), dev)
    check(_lib_().tt_topk_scan_p2p(_p(index), int(index.dtype == torch.bfloat16), _p(queries), N, H, nq, int(k),
                                   int(bool(cosine)), int(id_offset), C.byref(xch.desc), _p(scores), _p(ids), _p(workspace),
                                   workspace.numel(), _stream()), "tt_topk_scan_p2p")
    return scores, ids


def topk_merge(scores: torch.Tensor, ids: torch.Tensor):
    """scores/ids [R,nq,k] -> merged ([nq,k], [nq,k])."""
    _need_cuda(scores, ids)
    scores = _f32(scores)
    ids = ids.contiguous()
    R, nq, k = scores.shape
    out_s = torch.empty(nq, k, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(nq, k, dtype=torch.int64, device=scores.device)
    check(_lib_().tt_topk_merge(_p(scores), _p(ids), R, nq, k, 0, 0, _p(out_s), _p(out_i), _stream()), "tt_topk_merge")
    return out_s, out_i


def packed_topk_buffer(nq: int, k: int, device, ranks: int = 1):
    """One byte buffer per rank holding [nq,k] fp32 scores followed by [nq,k] int64 ids (8-byte aligned), so the
    sharded search needs ONE all-gather.  Returns (buffer [ranks, nbytes] uint8, nbytes, id byte offset)."""
    sbytes = (nq * k * 4 + 7) // 8 * 8
    nbytes = sbytes + nq * k * 8
    return torch.empty(ranks, nbytes, dtype=torch.uint8, device=device), nbytes, sbytes


def packed_views(buf_row: torch.Tensor, nq: int, k: int, id_off: int):
    s = buf_row[:nq * k * 4].view(torch.float32).view(nq, k)
    i = buf_row[id_off:id_off + nq * k * 8].view(torch.int64).view(nq, k)
    return s, i


def topk_merge_packed(buf: torch.Tensor, nq: int, k: int, id_off: int):
    """Merge R packed candidate records (see packed_topk_buffer) -> ([nq,k], [nq,k])."""
    _need_cuda(buf)
    R = buf.shape[0]
    nbytes = buf.stride(0) if R > 1 else buf.shape[1]          # records may sit in wider slots (peer-memory exchange buffer)
    if nbytes % 8 != 0 or buf.stride(1) != 1:
        raise ValueError("packed records need an 8-byte aligned rank pitch")
    out_s = torch.empty(nq, k, dtype=torch.float32, device=buf.device)
    out_i = torch.empty(nq, k, dtype=torch.int64, device=buf.device)
    base = buf.data_ptr()
    check(_lib_().tt_topk_merge(C.c_void_p(base), C.c_void_p(base + id_off), R, nq, k, nbytes // 4, nbytes // 8,
                                _p(out_s), _p(out_i), _stream()), "tt_topk_merge")
    return out_s, out_i


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(src)
    src = _f32(src)
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    check(_lib_().tt_cast_f32_to_bf16(_p(src), _p(dst), src.numel(), _stream()), "tt_cast_f32_to_bf16")
    return dst


# --------------------------------------------------------------------------------------
# optimizer
# --------------------------------------------------------------------------------------
def adamw_step(param, grad, exp_avg, exp_avg_sq, step_count, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
               weight_decay=0.01, param_bf16=None):
    _need_cuda(param, grad, exp_avg, exp_avg_sq, step_count)
    check(_lib_().tt_adamw_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), float(lr),
                                float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                _p(step_count), _p(param_bf16), _stream()), "tt_adamw_step")


# --------------------------------------------------------------------------------------
# autograd glue (so the reference's `loss.backward(); optimizer.step()` loop works unchanged)
# --------------------------------------------------------------------------------------
class EmbedPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table):
        pooled, inv_len, _ = embed_pool_fwd(ids, table)
        ctx.save_for_backward(ids, inv_len)
        ctx.V = table.shape[0]
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        ids, inv_len = ctx.saved_tensors
        d_table = embed_pool_bwd(ids, inv_len, d_pooled, ctx.V) if ctx.needs_input_grad[1] else None
        return None, d_table


class EmbedGatherFn(torch.autograd.Function):
    """[B,L] -> [B,L,E] (API compatibility with LookupEmbedding.forward).  The backward treats
    every token as its own row of length 1 so the same deterministic scatter kernel applies."""

    @staticmethod
    def forward(ctx, ids, table):
        ctx.save_for_backward(ids)
        ctx.V = table.shape[0]
        return embed_gather(ids, table)

    @staticmethod
    def backward(ctx, d_out):
        (ids,) = ctx.saved_tensors
        if not ctx.needs_input_grad[1]:
            return None, None
        flat = ids.reshape(-1, 1)
        ones = torch.ones(flat.shape[0], dtype=torch.float32, device=d_out.device)
        return None, embed_pool_bwd(flat, ones, d_out.reshape(flat.shape[0], -1), ctx.V)


class MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, precision):
        y, h1, z, _ = mlp_fwd(x, w1, b1, w2, b2, precision)
        ctx.save_for_backward(x, w1, w2, h1, z)
        ctx.precision = precision
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, h1, z = ctx.saved_tensors
        dx, dw1, db1, dw2, db2 = mlp_bwd(dy, x, w1, w2, h1, z, ctx.needs_input_grad[0], ctx.precision)
        return dx, dw1, db1, dw2, db2, None


class ProjLnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, has_projection, dropout_p, training, seed, precision=None):
        y, a, stats, z = proj_ln_fwd(x, w, b, gamma, beta, has_projection, dropout_p, training, seed, precision)
        if has_projection:
            ctx.save_for_backward(x, w, gamma, a, stats, z)
        else:
            ctx.save_for_backward(x)
        ctx.cfg = (has_projection, dropout_p, training, seed, precision)
        return y

    @staticmethod
    def backward(ctx, dy):
        has_projection, dropout_p, training, seed, precision = ctx.cfg
        if has_projection:
            x, w, gamma, a, stats, z = ctx.saved_tensors
            dx, dw, db, dg, dbt = proj_ln_bwd(dy, x, w, gamma, a, stats, z, True, dropout_p, training, seed,
                                              ctx.needs_input_grad[0], precision)
            return dx, dw, db, dg, dbt, None, None, None, None, None
        (x,) = ctx.saved_tensors
        dx, *_ = proj_ln_bwd(dy, x, None, None, None, None, None, False, 0.0, False, 0)
        return dx, None, None, None, None, None, None, None, None, None


class InBatchLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, d, temperature, label_offset, loss_scale, precision):
        loss, lse, _ = inbatch_ce_fwd(q, d, temperature, label_offset, loss_scale, precision)
        ctx.save_for_backward(q, d, lse)
        ctx.cfg = (temperature, label_offset, loss_scale, precision)
        return loss

    @staticmethod
    def backward(ctx, g):
        q, d, lse = ctx.saved_tensors
        temperature, label_offset, loss_scale, precision = ctx.cfg
        dq, dd = inbatch_ce_bwd(q, d, lse, temperature, label_offset, loss_scale, g.contiguous(),
                                ctx.needs_input_grad[0], ctx.needs_input_grad[1], precision)
        return dq, dd, None, None, None, None


class TripletLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, p, n, margin):
        loss, sims, _, _ = triplet_fwd(q, p, n, margin)
        ctx.save_for_backward(q, p, n, sims)
        ctx.margin = margin
        return loss

    @staticmethod
    def backward(ctx, g):
        q, p, n, sims = ctx.saved_tensors
        dq, dp, dn = triplet_bwd(q, p, n, sims, ctx.margin, g.contiguous())
        return dq, dp, dn, None


class MultiNegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, p, negs, temperature):
        loss, probs = multineg_fwd(q, p, negs, temperature)
        ctx.save_for_backward(q, p, negs, probs)
        ctx.temperature = temperature
        return loss

    @staticmethod
    def backward(ctx, g):
        q, p, negs, probs = ctx.saved_tensors
        dq, dp, dnegs = multineg_bwd(q, p, negs, probs, ctx.temperature, g.contiguous())
        return dq, dp, dnegs, None


# --------------------------------------------------------------------------------------
# tensor-core self-test hook (tests only)
# --------------------------------------------------------------------------------------
def selftest_tc_gemm(a_bf16: torch.Tensor, a_mn_major: bool, b_bf16: torch.Tensor, b_mn_major: bool, M: int, N: int,
                     K: int, splits: int = 1) -> torch.Tensor:
    _need_cuda(a_bf16, b_bf16)
    c = torch.empty(M, N, dtype=torch.float32, device=a_bf16.device)
    partial = torch.empty(splits * M * N, dtype=torch.float32, device=a_bf16.device) if splits > 1 else None
    check(_lib_().tt_selftest_tc_gemm(_p(a_bf16.contiguous()), int(a_mn_major), _p(b_bf16.contiguous()), int(b_mn_major),
                                      M, N, K, _p(c), splits, _p(partial), _stream()), "tt_selftest_tc_gemm")
    return c
