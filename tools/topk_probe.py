#!/usr/bin/env python
"""Batched tensor-core scan: time vs nq / N / k (what bounds tt_topk_scan_batched)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import two_towers_b200 as tt

dev = torch.device("cuda", 0)
H = 256
def index(N):
    D = torch.empty(N, H, device=dev)
    for s in range(0, N, 1_000_000):
        e = min(N, s + 1_000_000)
        D[s:e] = torch.nn.functional.normalize(torch.randn(e - s, H, device=dev), dim=-1)
    return tt.ops.cast_bf16(D)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for N in (int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000,):
    idx = index(N)
    q = torch.nn.functional.normalize(torch.randn(256, H, device=dev), dim=-1)
    ws = torch.empty(tt.ops._lib_().tt_topk_scan_batched_workspace(N, H, 128), dtype=torch.uint8, device=dev)
    for k in (1, 10, 100):
        for nq in (1, 16, 64, 128):
            ms = t(lambda: tt.ops.topk_scan_batched(idx, q[:nq].contiguous(), k, workspace=ws))
            print(f"N={N} k={k:3d} nq={nq:3d}: {ms:8.3f} ms   ({N * H * 2 / ms / 1e6:7.1f} GB/s)")
    ms = t(lambda: tt.ops.topk_scan(idx, q[:1].contiguous(), 100, cosine=False))
    print(f"N={N} single-query scan kernel: {ms:.3f} ms")
