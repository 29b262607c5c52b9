#!/bin/bash
# per-CTA timeline of the fused CE backward as the trainer runs it
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TT_CE_DEBUG=${CE_PASS:-0} timeout 300 python - > $O/ce_timeline_fused.log 2>&1 <<'PY'
import ctypes as C, torch, two_towers_b200 as tt
torch.manual_seed(0)
emb = tt.embeddings.build("lookup", 128, embedding_dim=64)
model = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to("cuda")
tr = tt.FusedTrainer(model, loss="in_batch", batch_size=4096, max_len=64, precision="bf16", use_cuda_graph=False)
g = torch.Generator().manual_seed(1)
q = torch.randint(1, 128, (4096, 64), generator=g); d = torch.randint(1, 128, (4096, 64), generator=g)
tr.step(q, d); torch.cuda.synchronize()
print("==== second step")
tr.step(q, d); torch.cuda.synchronize()
PY
awk '/==== second step/{f=1} f' $O/ce_timeline_fused.log | grep -A18 "timeline, cycles" | head -20
awk '/==== second step/{f=1} f' $O/ce_timeline_fused.log | grep -A200 "per-CTA" | awk "NR%16==2" | head -10
