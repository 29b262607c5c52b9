#!/usr/bin/env python
"""2+ ranks: all-gather of D over peer memory followed by the loss forward + dQ, gated (consumes blocks as they land) vs
plain (waits for the exchange kernel to retire); graph replays timed with CUDA events."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import two_towers_b200 as tt
from two_towers_b200 import _lib, parallel


def main():
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, H = 4096, 256
    Bg = B * world
    torch.manual_seed(rank)
    q = tt.ops.cast_bf16(torch.nn.functional.normalize(torch.randn(B, H, device=dev), dim=-1))
    d = tt.ops.cast_bf16(torch.nn.functional.normalize(torch.randn(B, H, device=dev), dim=-1))
    x = parallel.P2PExchange(B * H * 2, None, dev)
    dg = x.gathered(torch.bfloat16, (B, H)).view(Bg, H)
    vp = lambda t: None if t is None else t.data_ptr()
    dz = torch.zeros(B, H, dtype=torch.bfloat16, device=dev); cs = torch.zeros(B // 32, H, device=dev); inv = torch.ones(B, device=dev)
    sync = torch.zeros(int(lib.tt_inbatch_ce_onepass_sync_bytes(B)), dtype=torch.uint8, device=dev)
    loss = torch.zeros((), device=dev); lse = torch.zeros(B, device=dev)
    qp = _lib.CePass(vp(q), B, vp(dg), Bg, Bg, Bg, 0, 0, None, rank * B, None, 0, vp(dz), vp(cs), vp(inv))
    res = {}
    for mode in ("plain", "gated", "exchange_only", "loss_only"):
        def body():
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            if mode != "loss_only":
                x.allgather(d, count=False)
            if mode == "gated":
                _lib.check(lib.tt_inbatch_ce_fwd_dq_p2p(C.byref(qp), H, 10.0, 10.0, 1.0 / Bg, None, vp(loss), vp(lse), None, vp(sync), C.byref(x.desc), vp(d), s), "gated")
            elif mode in ("plain", "loss_only"):
                _lib.check(lib.tt_inbatch_ce_fwd_dq(C.byref(qp), H, 10.0, 10.0, 1.0 / Bg, None, vp(loss), vp(lse), None, vp(sync), s), "plain")
        body(); torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        ts = []
        for i in range(30):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            if i >= 5:
                ts.append(e0.elapsed_time(e1) * 1e3)
        t = torch.tensor(sum(ts) / len(ts), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # eager, 50 iterations back to back (no graph): average device time per iteration
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            body()
        e1.record(); torch.cuda.synchronize()
        te = torch.tensor(e0.elapsed_time(e1) * 1e3 / 50, device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # graph of 10 iterations
        g10 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g10):
            for _ in range(10):
                body()
        t10 = []
        for i in range(12):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g10.replay(); e1.record(); torch.cuda.synchronize()
            if i >= 2:
                t10.append(e0.elapsed_time(e1) * 1e3 / 10)
        tg = torch.tensor(sum(t10) / len(t10), device=dev, dtype=torch.float64)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        res[mode] = (t.item(), loss.item(), float(lse.double().sum().item()), te.item(), tg.item())
    if rank == 0:
        for k, v in res.items():
            print(f"  {k:14s} graph x1 {v[0]:7.1f} us   eager x50 {v[3]:7.1f} us/iter   graph x10 {v[4]:7.1f} us/iter   loss {v[1]:.6f}  sum(lse) {v[2]:.4f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
