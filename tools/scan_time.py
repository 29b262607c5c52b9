import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import two_towers_b200 as tt
dev = torch.device("cuda", 0)
H, k = 256, 100
for rows in (1_250_000, 10_000_000):
    D = torch.empty(rows, H, device=dev)
    for s in range(0, rows, 1_000_000):
        e = min(rows, s + 1_000_000)
        D[s:e] = torch.nn.functional.normalize(torch.randn(e - s, H, device=dev), dim=-1)
    qs = torch.nn.functional.normalize(torch.randn(64, H, device=dev), dim=-1)
    for dt in ("fp32", "bf16"):
        idx = D if dt == "fp32" else tt.ops.cast_bf16(D)
        ws = torch.empty(tt.ops.topk_scan_workspace_bytes(rows, H, 1, k), dtype=torch.uint8, device=dev)
        for i in range(3): tt.ops.topk_scan(idx, qs[i:i+1], k, cosine=False, workspace=ws)
        torch.cuda.synchronize()
        ts = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(50): tt.ops.topk_scan(idx, qs[i % 64:i % 64 + 1], k, cosine=False, workspace=ws)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 50)
        t = float(np.median(ts)); b = rows * H * (4 if dt == "fp32" else 2)
        print(f"rows {rows} {dt}: {t:8.1f} us/query  {b / t / 1e3:6.0f} GB/s", flush=True)
        del idx
    del D
