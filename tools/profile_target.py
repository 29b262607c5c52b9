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
    ap.add_argument("--what", default="train", choices=["train", "search", "word", "search_batched", "msmarco"])
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
    elif a.what in ("word", "msmarco"):
        # configs[2] (word tower: V = 400k, E = 300, L = 32, trainable table, in-batch) / configs[3] (untied, H = 128)
        import bench
        if a.what == "word":
            c = bench.WORD
            emb = tt.embeddings.build("lookup", c["V"], embedding_dim=c["E"])
            model = tt.build_two_tower("mean", emb, hidden_dim=c["H"], tied_weights=True).to(dev)
            q, d = bench.zipf_ids(c["B"], c["L"], c["V"], 1).to(dev), bench.zipf_ids(c["B"], c["L"], c["V"], 2).to(dev)
        else:
            c = bench.MSM
            emb = tt.embeddings.build("lookup", c["V"], embedding_dim=c["E"])
            model = tt.build_two_tower("mean", emb, hidden_dim=c["H"], tied_weights=False).to(dev)
            q, d = (bench.repeated_ids(c["B"], c["L"], c["V"], s_, c["repeat"]).to(dev) for s_ in (1, 2))
        tr = tt.FusedTrainer(model, loss="in_batch", batch_size=c["B"], max_len=c["L"], precision=a.precision,
                             use_cuda_graph=False, id_dtype=torch.int32)
        for _ in range(a.iters):
            loss = tr.step(q, d)
        torch.cuda.synchronize()
        print(a.what, "ok, loss", loss.item(), "launches/step", tr.kernels_per_step())
    elif a.what == "search_batched":
        N, H = a.rows, 256
        D = torch.empty(N, H, device=dev)
        for s in range(0, N, 1_000_000):
            e = min(N, s + 1_000_000)
            D[s:e] = torch.nn.functional.normalize(torch.randn(e - s, H, device=dev), dim=-1)
        idx = tt.ops.cast_bf16(D)
        del D
        q = torch.nn.functional.normalize(torch.randn(64, H, device=dev), dim=-1)
        for _ in range(a.iters):
            s, i = tt.ops.topk_scan_batched(idx, q, 100)
        torch.cuda.synchronize()
        print("batched search ok, best", s[0, 0].item(), i[0, 0].item())
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
