#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TT_CE_DEBUG=${CE_PASS:-0} timeout 300 python - > $O/ce_timeline.log 2>&1 <<'PY'
import torch, two_towers_b200 as tt
B,H=4096,256
q=torch.nn.functional.normalize(torch.randn(B,H,device='cuda'),dim=-1); d=torch.nn.functional.normalize(torch.randn(B,H,device='cuda'),dim=-1)
loss,lse,_=tt.ops.inbatch_ce_fwd(q,d,0.1,precision='bf16')
for i in range(2):
    dq,dd=tt.ops.inbatch_ce_bwd(q,d,lse,0.1,precision='bf16')
torch.cuda.synchronize()
PY
tail -40 $O/ce_timeline.log
