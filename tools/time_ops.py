#!/usr/bin/env python
"""Per-op device time of the fused train step at the BASELINE shape: each C-ABI call is captured in its own
CUDA graph and replayed between L2 flushes with CUDA events around it (no host launch gaps, no profiler)."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import two_towers_b200 as tt
from two_towers_b200 import _lib
from two_towers_b200.train import _p


FLUSH_MODE = "write"
FLUSH_BUF2 = None
FLUSH_SINK = []


def timed(fn, flush, reps=43):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for i in range(reps):
        if FLUSH_MODE != "none":
            flush.add_(1)
        if FLUSH_MODE == "write_read":          # read a second buffer: dirty lines are written back, L2 is cold AND clean
            FLUSH_SINK.append(float(0) if FLUSH_BUF2 is None else 0.0)
            torch.sum(FLUSH_BUF2, dtype=torch.float32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.mean(ts))             # event resolution is 2.048 us: the mean over 40 replays resolves finer steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--flush", default="write", choices=["write", "write_read", "none"])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    global FLUSH_MODE, FLUSH_BUF2
    FLUSH_MODE = a.flush
    if a.flush == "write_read":
        FLUSH_BUF2 = torch.ones(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    torch.manual_seed(0)
    emb = tt.embeddings.build("lookup", 128, embedding_dim=64)
    model = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to(dev)
    tr = tt.FusedTrainer(model, loss="in_batch", batch_size=a.B, max_len=64, precision=a.precision, use_cuda_graph=False)
    g = torch.Generator().manual_seed(1)
    q = torch.randint(1, 128, (a.B, 64), generator=g); d = torch.randint(1, 128, (a.B, 64), generator=g)
    tr.step(q, d); torch.cuda.synchronize()
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    lib, B, H, R = tr.lib, tr.B, tr.H, 2 * tr.B
    s = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    from two_towers_b200._lib import check
    qb = tr.y_bf16[:B] if tr.y_bf16 is not None else None
    db = tr.y_bf16[B:2 * B] if tr.y_bf16 is not None else None
    qf, df = tr.y[:B], tr.y[B:2 * B]
    ops = {
        "embed_pool_fwd": lambda: check(lib.tt_embed_pool_fwd(_p(tr.ids), 8, _p(tr.table), R, tr.L, tr.V, tr.E, None if tr.embed_in_tower else _p(tr.pooled), _p(tr.inv_len), None if tr.embed_in_tower else _p(tr.pooled_bf16), _p(tr.pool_bf16), s()), "x"),
        "tower_fwd": lambda: tr._tower_fwd(0),
        "ce_fwd": lambda: tr._local_loss_fwd(s()),
        "ce_bwd": lambda: tr._local_loss_bwd(s()),
        "tower_bwd": lambda: tr._tower_bwd(0),
        "embed_pool_bwd": lambda: check(lib.tt_embed_pool_bwd(_p(tr.ids), 8, _p(tr.inv_len), _p(tr.dpooled), R, tr.L, tr.V, tr.E, _p(tr.table.grad), _p(tr.ws), tr.ws.numel(), s()), "x"),
        "adamw": lambda: check(lib.tt_adamw_step(_p(tr.flat), _p(tr.flat_grad), _p(tr.exp_avg), _p(tr.exp_avg_sq), tr.n_params, 1e-3, 0.9, 0.999, 1e-8, 0.01, _p(tr.step_count), _p(tr.flat_bf16), s()), "x"),
        "whole_step": lambda: tr._step_impl(),
    }
    if tr.onepass:                            # forward + dQ in one launch, dD in the other
        vp = lambda t: None if t is None else t.data_ptr()
        nb = B // 32
        qp = _lib.CePass(vp(qb), B, vp(db), B, B, B, 0, 0, None, 0, None, 0, vp(tr.dz_bf16[:B]), vp(tr.dz_colsum[:nb]), vp(tr.inv_norm[:B]))
        dp = _lib.CePass(vp(db), B, vp(qb), B, B, B, 0, 0, vp(tr.lse), 0, None, 0, vp(tr.dz_bf16[B:2 * B]), vp(tr.dz_colsum[nb:2 * nb]), vp(tr.inv_norm[B:2 * B]))
        inv_t = 1.0 / tr.temperature
        ops["ce_fwd"] = lambda: check(lib.tt_inbatch_ce_fwd_dq(C.byref(qp), H, inv_t, inv_t, 1.0 / B, None, _p(tr.loss), _p(tr.lse), _p(tr.pos_mean), _p(tr.onepass_sync), s()), "x")
        ops["ce_bwd"] = lambda: check(lib.tt_inbatch_ce_dd(C.byref(dp), H, inv_t, 1.0 / B, None, s()), "x")
        if getattr(tr, "ce_stash", None) is not None:   # stored-E form: the first launch stores E, the second is a plain product
            ops["ce_fwd"] = lambda: check(lib.tt_inbatch_ce_fwd_dq_stash(C.byref(qp), H, inv_t, inv_t, 1.0 / B, None, _p(tr.loss), _p(tr.lse), _p(tr.pos_mean), _p(tr.onepass_sync), _p(tr.ce_stash), s()), "x")
            ops["ce_bwd"] = lambda: check(lib.tt_inbatch_ce_dd_stash(C.byref(dp), H, inv_t, 1.0 / B, None, _p(tr.ce_stash), s()), "x")
            print("loss: stored-E form (tt_inbatch_ce_fwd_dq_stash + tt_inbatch_ce_dd_stash)")
        ops = {("ce_fwd_dq" if k == "ce_fwd" else "ce_dd" if k == "ce_bwd" else k): v for k, v in ops.items()}
        if tr.onelaunch:                          # both as one kernel
            ops = {k: v for k, v in ops.items() if k != "ce_dd"}
            ops = {("ce_onepass" if k == "ce_fwd_dq" else k): (lambda: tr._local_loss_onepass(s())) if k == "ce_fwd_dq" else v for k, v in ops.items()}
    if getattr(tr, "pool_in_tower", False):
        del ops["embed_pool_fwd"]            # the tower kernel builds the pooling matrix itself
    if tr.embed_fused:
        del ops["embed_pool_bwd"]            # folded into the tower backward (tt_mlp_embed_t)
    total = 0.0
    for name, fn in ops.items():
        before = _lib.launch_count()
        fn()
        nl = _lib.launch_count() - before
        t = timed(fn, flush)
        if name != "whole_step":
            total += t
        print(f"{name:16s} {t:9.1f} us   ({nl} launches)")
    print(f"{'sum of ops':16s} {total:9.1f} us")
    # cumulative prefixes of the step in ONE graph: the differences are each op's cost inside the pipeline
    # (launch gaps included, the single-graph replay floor cancelled)
    names = [n for n in ops if n != "whole_step"]
    prev = 0.0
    for k in range(1, len(names) + 1):
        def prefix(k=k):
            for n in names[:k]:
                ops[n]()
        t = timed(prefix, flush)
        print(f"prefix..{names[k - 1]:16s} {t:9.1f} us   (+{t - prev:6.1f})")
        prev = t
    # back-to-back repeats of one op inside one graph: (T10 - T1) / 9 = kernel duration + launch boundary, without the
    # single-graph replay floor
    for name in names:
        t1 = timed(ops[name], flush)
        def rep10(name=name):
            for _ in range(10):
                ops[name]()
        t10 = timed(rep10, flush)
        print(f"repeat {name:16s} T1 {t1:7.1f}  T10 {t10:7.1f}  per launch-group {(t10 - t1) / 9:6.1f} us")
    # launch-boundary cost: 32 back-to-back AdamW launches over a tiny slice
    def chain():
        for _ in range(32):
            check(lib.tt_adamw_step(_p(tr.flat), _p(tr.flat_grad), _p(tr.exp_avg), _p(tr.exp_avg_sq), 1024, 1e-3, 0.9, 0.999, 1e-8, 0.01, _p(tr.step_count), _p(tr.flat_bf16), s()), "x")
    t = timed(chain, flush)
    print(f"32 tiny launches {t:9.1f} us   ({t / 32:.2f} us per boundary, TT_PDL={os.environ.get('TT_PDL', '1')})")


if __name__ == "__main__":
    main()
