#!/usr/bin/env python
"""Stress of the scan's dynamic round scheduling (fp32 x 256 columns): many back-to-back launches, eager and inside a CUDA
graph, one and two queries per pass; every answer must equal the answer of a second pass over the same queries, and a
sample of them the fp64 oracle's."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import two_towers_b200 as tt

dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, H, k = 3_000_001, 256, 100
D = torch.nn.functional.normalize(torch.randn(N, H, device=dev), dim=-1)
qs = torch.nn.functional.normalize(torch.randn(512, H, device=dev), dim=-1)
ws = torch.empty(tt.ops.topk_scan_workspace_bytes(N, H, 2, k), dtype=torch.uint8, device=dev)
for nq in (1, 2):
    first = []
    for i in range(0, 512, nq):
        s, ids = tt.ops.topk_scan(D, qs[i:i + nq], k, cosine=False, workspace=ws)
        first.append((s.clone(), ids.clone()))
    torch.cuda.synchronize()
    for j, i in enumerate(range(0, 512, nq)):
        s, ids = tt.ops.topk_scan(D, qs[i:i + nq], k, cosine=False, workspace=ws)
        assert torch.equal(ids, first[j][1]) and torch.equal(s, first[j][0]), (nq, i)
    # a few against fp64 on the host
    for i in (0, 100, 510):
        sc = (qs[i:i + 1].double() @ D.double().T)[0]
        top = torch.topk(sc, k).indices
        assert set(top.tolist()) == set(first[i // nq][1][i % nq].tolist()), (nq, i)
    print(f"nq={nq}: {512 // nq * 2} eager launches consistent")
# graph replays
q1 = qs[:1].clone()
out = tt.ops.topk_scan(D, q1, k, cosine=False, workspace=ws)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = tt.ops.topk_scan(D, q1, k, cosine=False, workspace=ws)
for i in range(300):
    q1.copy_(qs[i % 512:i % 512 + 1])
    g.replay()
    if i % 50 == 0:
        torch.cuda.synchronize()
        ref = tt.ops.topk_scan(D, qs[i % 512:i % 512 + 1], k, cosine=False)
        assert torch.equal(out[1], ref[1]) and torch.equal(out[0], ref[0]), i
torch.cuda.synchronize()
print("300 graph replays consistent")
