#!/usr/bin/env python
"""Is the CE forward bound by operand streaming / MMA (scales with H) or by its epilogue (does not)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import two_towers_b200 as tt
from two_towers_b200 import _lib
lib = _lib.load()
dev = "cuda"
flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
def timed(fn, reps=23):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    ts = []
    for i in range(reps):
        flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3 / 10)
    return float(np.mean(ts))
for B, Bd, H in [(4096, 4096, 256), (4096, 4096, 128), (4096, 4096, 64), (4096, 32768, 256), (4096, 32768, 128)]:
    q = tt.ops.cast_bf16(torch.nn.functional.normalize(torch.randn(B, H, device=dev), dim=-1))
    d = tt.ops.cast_bf16(torch.nn.functional.normalize(torch.randn(Bd, H, device=dev), dim=-1))
    ws = torch.empty(int(lib.tt_inbatch_ce_fwd_ex_workspace(B, Bd)), dtype=torch.uint8, device=dev)
    sync = torch.zeros(int(lib.tt_inbatch_ce_sync_bytes(B)), dtype=torch.uint8, device=dev)
    loss = torch.zeros((), device=dev); lse = torch.zeros(B, device=dev); pm = torch.zeros((), device=dev)
    def fwd():
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.tt_inbatch_ce_fwd_ex(q.data_ptr(), B, d.data_ptr(), Bd, Bd, Bd, 0, 0, H, 10.0, 0, 1.0 / B, loss.data_ptr(),
                                            lse.data_ptr(), pm.data_ptr(), ws.data_ptr(), ws.numel(), sync.data_ptr(), s), "fwd")
    t = timed(fwd)
    print(f"ce_fwd B={B} Bd={Bd} H={H}: {t:7.1f} us per launch   ({2.0 * B * Bd * H / t / 1e6:7.1f} TFLOP/s)")
