#!/usr/bin/env python
"""Sharded-search timing under torchrun (one rank per GPU): per-query device time of the graph-replayed ShardedTopK with
the fused merge + exchange + merge kernel (tt_topk_scan_p2p) and with the three-launch chain it replaces, next to this
shard's scan alone.  --rows = rows PER SHARD (default: the 8-shard share of the 10 M x 256 index).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/search_probe.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import two_towers_b200 as tt
from two_towers_b200 import parallel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_250_000)
    ap.add_argument("--reps", type=int, default=200)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    H, k = 256, 100
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    D = torch.nn.functional.normalize(torch.randn(a.rows, H, device=dev, generator=gen), dim=-1)
    qs = torch.nn.functional.normalize(torch.randn(64, H, device=dev, generator=torch.Generator(device=dev).manual_seed(11)), dim=-1)
    lo = rank * a.rows

    def timed(fn):
        for i in range(5):
            fn(i)
        blocks = []
        for rep in range(5):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.reps):
                fn(i)
            e1.record(); torch.cuda.synchronize()
            blocks.append(e0.elapsed_time(e1) * 1e3 / a.reps)
        t = torch.tensor(blocks, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(np.median(t.cpu().numpy()))

    for dt in ("fp32", "bf16"):
        index = D if dt == "fp32" else tt.ops.cast_bf16(D)
        ws = torch.empty(tt.ops.topk_scan_workspace_bytes(a.rows, H, 1, k), dtype=torch.uint8, device=dev)
        t_scan = timed(lambda i: tt.ops.topk_scan(index, qs[i % 64:i % 64 + 1], k, cosine=False, id_offset=lo, workspace=ws))
        res, outs = {}, {}
        for fused in ("1", "0"):
            os.environ["TT_SEARCH_FUSED"] = fused
            st = parallel.ShardedTopK(index, k, lo, tt.ops, None, cosine=False, nq=1)
            assert st.fused == (fused == "1"), (st.fused, fused)
            res[fused] = timed(lambda i: st(qs[i % 64:i % 64 + 1]))
            outs[fused] = [tuple(t.clone() for t in st(qs[i:i + 1])) for i in range(8)]
        os.environ.pop("TT_SEARCH_FUSED")
        torch.cuda.synchronize()
        for (s1, i1), (s0, i0) in zip(outs["1"], outs["0"]):     # same answers from both forms, on every rank
            assert torch.equal(s1, s0) and torch.equal(i1, i0)
            ref = [torch.empty_like(i1) for _ in range(world)]
            dist.all_gather(ref, i1)
            assert all(torch.equal(r, i1) for r in ref)
        if rank == 0:
            byts = a.rows * H * (4 if dt == "fp32" else 2)
            print(f"{dt}: world {world}, {a.rows} rows/shard: local scan + block merge (eager, 2 launches) {t_scan:7.1f} us "
                  f"({byts / t_scan / 1e3:.0f} GB/s) | sharded, fused kernel {res['1']:7.1f} us | sharded, three-launch chain {res['0']:7.1f} us", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
