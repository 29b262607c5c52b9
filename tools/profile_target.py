#!/usr/bin/env python
"""Short, graph-free workload for ncu: a few eager fused train steps at the BASELINE shape
(B=4096, L=64, E=64, d=256) and a few top-100 scans over a 10M x 256 index.

    python tools/profile_target.py --what train --precision bf16
    python tools/profile_target.py --what search --dtype fp32
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import two_towers_b200 as tt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="train", choices=["train", "search"])
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if a.what == "train":
        emb = tt.embeddings.build("lookup", 128, embedding_dim=64)
        model = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to(dev)
        tr = tt.FusedTrainer(model, loss="in_batch", batch_size=4096, max_len=64, precision=a.precision,
                             use_cuda_graph=False)
        g = torch.Generator().manual_seed(1)
        q = torch.randint(1, 128, (4096, 64), generator=g)
        d = torch.randint(1, 128, (4096, 64), generator=g)
        for _ in range(a.iters):
            loss = tr.step(q, d)
        torch.cuda.synchronize()
        print("train ok, loss", loss.item(), "launches/step", tr.kernels_per_step())
    else:
        N, H = a.rows, 256
        D = torch.empty(N, H, device=dev)
        for s in range(0, N, 1_000_000):
            e = min(N, s + 1_000_000)
            D[s:e] = torch.nn.functional.normalize(torch.randn(e - s, H, device=dev), dim=-1)
        idx = D if a.dtype == "fp32" else tt.ops.cast_bf16(D)
        q = torch.nn.functional.normalize(torch.randn(1, H, device=dev), dim=-1)
        for _ in range(a.iters):
            s, i = tt.ops.topk_scan(idx, q, 100, cosine=False)
        torch.cuda.synchronize()
        print("search ok, best", s[0, 0].item(), i[0, 0].item())


if __name__ == "__main__":
    main()
