#!/usr/bin/env python
"""Per-phase device timeline of the fused train step at the BASELINE shape on N ranks (torchrun), eager (no CUDA graph):
one CUDA event per phase boundary on the step's stream, averaged over the timed steps, reported as mean and max over
ranks.  Phases that end in a peer-memory exchange include the wait for the slowest rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/step_timeline.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import two_towers_b200 as tt


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    torch.manual_seed(0)
    B, L, V = 4096, 64, 128
    emb = tt.embeddings.build("lookup", V, embedding_dim=64)
    model = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to(dev)
    tr = tt.FusedTrainer(model, loss="in_batch", batch_size=B, max_len=L, precision="bf16", process_group=pg,
                         global_negatives=True, id_dtype=torch.int32, use_cuda_graph=False)
    g = torch.Generator().manual_seed(1 + rank)
    q = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).to(dev)
    d = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).to(dev)
    tr.load_batch(q, d)
    for _ in range(10):
        tr.run()
    torch.cuda.synchronize()
    # phase boundaries of one eager step (names only)
    tr._trace = []
    tr.run()
    torch.cuda.synchronize()
    names = [n for n, _ in tr._trace[1:]]
    tr._trace = None
    # prefix graphs: the step's launches up to each boundary, captured and replayed on their own; the differences of the
    # replay times are each phase's cost inside the pipeline (launch gaps as in the real step, no host in the way)
    st = tr._snapshot()
    times = []
    for n in names:
        tr.step_prefix(n)                                    # warm (and keep every rank in step)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            tr.step_prefix(n)
        ts = []
        for i in range(25):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(e0.elapsed_time(e1) * 1e3)
        times.append(sum(ts) / len(ts))
        del g
    tr._restore(st)
    acc = torch.tensor(times, dtype=torch.float64, device=dev)
    mx, mean = acc.clone(), acc.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mean, op=dist.ReduceOp.SUM)
        mean /= world
    if rank == 0:
        print(f"[step timeline] {world} rank(s), B={B}/rank, global negatives, bf16, onepass={tr.onepass}, p2p={getattr(tr, "p2p", False)}, gated={getattr(tr, "gated", False)}")
        print("  prefix graph ending after phase      replay us (mean / max over ranks)   + this phase (max)")
        prev = 0.0
        for n, a, b in zip(names, mean.tolist(), mx.tolist()):
            print(f"  {n:34s} {a:8.1f} {b:8.1f}      +{b - prev:7.1f}")
            prev = b
        print("  (a replay of a one-kernel graph costs ~8-10 us: the first row carries that floor; no L2 flush between replays)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
