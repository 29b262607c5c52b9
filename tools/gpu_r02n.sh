#!/bin/bash
# how much of the step is cold-cache cost: per-op times with and without the L2 flush between replays; e2e loop speed
cd "$(dirname "$0")/.."
TAG=${1:-r02n}
O=gpurun_out; mkdir -p $O
timeout 300 python tools/time_ops.py --flush none > $O/${TAG}_time_ops_noflush.log 2>&1; echo "noflush exit $?"; head -16 $O/${TAG}_time_ops_noflush.log
timeout 300 python tools/time_ops.py > $O/${TAG}_time_ops.log 2>&1; echo "flush exit $?"; head -16 $O/${TAG}_time_ops.log
timeout 300 python - > $O/${TAG}_loop.log 2>&1 <<'PY'
import time, torch, two_towers_b200 as tt
dev = torch.device("cuda", 0)
emb = tt.embeddings.build("lookup", 128, embedding_dim=64)
model = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to(dev)
tr = tt.FusedTrainer(model, loss="in_batch", batch_size=4096, max_len=64, precision="bf16", id_dtype=torch.int32)
g = torch.Generator().manual_seed(1)
q = torch.randint(1, 128, (4096, 64), generator=g, dtype=torch.int32).to(dev); d = torch.randint(1, 128, (4096, 64), generator=g, dtype=torch.int32).to(dev)
tr.load_batch(q, d)
for _ in range(20): tr.run()
torch.cuda.synchronize()
for n in (50, 200, 1000):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): tr.run()
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{n} back-to-back graph replays, no flush: {e0.elapsed_time(e1) * 1e3 / n:.1f} us/step device, {(t1 - t0) * 1e6 / n:.1f} us/step wall")
PY
cat $O/${TAG}_loop.log | tail -5
