#!/bin/bash
# cold (in-step) profile of the loss kernels: eager steps with the in-kernel stamps on
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TT_CE_DEBUG=1 timeout 200 python - > $O/r03p_cold_timeline.log 2>&1 <<'PY'
import torch, two_towers_b200 as tt
dev = torch.device("cuda", 0)
emb = tt.embeddings.build("lookup", 128, embedding_dim=64)
model = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to(dev)
tr = tt.FusedTrainer(model, loss="in_batch", batch_size=4096, max_len=64, precision="bf16", id_dtype=torch.int32, use_cuda_graph=False)
g = torch.Generator().manual_seed(1)
q = torch.randint(1, 128, (4096, 64), generator=g, dtype=torch.int32).to(dev); d = torch.randint(1, 128, (4096, 64), generator=g, dtype=torch.int32).to(dev)
x = torch.randn(8192, 8192, device=dev)
for _ in range(20): y = x @ x
for _ in range(6):
    tr.step(q, d)
torch.cuda.synchronize()
PY
echo "exit $?"
awk '/ce_fwd_dq per-CTA/{n++} n==6' $O/r03p_cold_timeline.log | head -12
grep "ce_fwd_dq tail" $O/r03p_cold_timeline.log | tail -2
awk '/ce_bwd per-CTA/{n++} n==6' $O/r03p_cold_timeline.log | head -8
grep "ce_bwd tail" $O/r03p_cold_timeline.log | tail -2
