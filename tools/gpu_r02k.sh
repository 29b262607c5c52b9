#!/bin/bash
# one-pass loss: parity tests, per-op times, in-kernel timeline
cd "$(dirname "$0")/.."
TAG=${1:-r02k}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_onepass.py -m gpu -q -s --maxfail=40 > $O/${TAG}_pytest_onepass.log 2>&1; echo "pytest onepass exit $?"
grep -E "^FAILED|passed|failed|error" $O/${TAG}_pytest_onepass.log | tail -30 | cut -c1-250
timeout 300 python tools/time_ops.py > $O/${TAG}_time_ops.log 2>&1; echo "time_ops exit $?"; head -12 $O/${TAG}_time_ops.log
TT_CE_ONEPASS=0 timeout 300 python tools/time_ops.py > $O/${TAG}_time_ops_legacy.log 2>&1; echo "time_ops legacy exit $?"; head -12 $O/${TAG}_time_ops_legacy.log
TT_CE_DEBUG=1 timeout 120 python - > $O/${TAG}_ce_timeline.log 2>&1 <<'PY'
import torch, two_towers_b200 as tt
q = tt.ops.cast_bf16(torch.nn.functional.normalize(torch.randn(4096, 256, device="cuda"), dim=-1))
d = tt.ops.cast_bf16(torch.nn.functional.normalize(torch.randn(4096, 256, device="cuda"), dim=-1))
import ctypes as C
from two_towers_b200 import _lib
lib = _lib.load()
B, H = 4096, 256
vp = lambda t: None if t is None else t.data_ptr()
dz = torch.zeros(2 * B, H, dtype=torch.bfloat16, device="cuda"); cs = torch.zeros(2 * B // 32, H, device="cuda"); inv = torch.ones(2 * B, device="cuda")
sync = torch.zeros(int(lib.tt_inbatch_ce_onepass_sync_bytes(B)), dtype=torch.uint8, device="cuda")
loss = torch.zeros((), device="cuda"); lse = torch.zeros(B, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
qp = _lib.CePass(vp(q), B, vp(d), B, B, B, 0, 0, None, 0, None, 0, vp(dz[:B]), vp(cs[:B // 32]), vp(inv[:B]))
dp = _lib.CePass(vp(d), B, vp(q), B, B, B, 0, 0, vp(lse), 0, None, 0, vp(dz[B:]), vp(cs[B // 32:]), vp(inv[B:]))
x = torch.randn(8192, 8192, device="cuda")
for _ in range(30):
    y = x @ x                      # clocks up
for _ in range(3):
    _lib.check(lib.tt_inbatch_ce_fwd_dq(C.byref(qp), H, 10.0, 10.0, 1.0 / B, None, vp(loss), vp(lse), None, vp(sync), s), "fwd_dq")
    _lib.check(lib.tt_inbatch_ce_dd(C.byref(dp), H, 10.0, 1.0 / B, None, s), "dd")
torch.cuda.synchronize()
PY
echo "timeline exit $?"; grep -A12 "tail, cycles" $O/${TAG}_ce_timeline.log | tail -60 | cut -c1-200
