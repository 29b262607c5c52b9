#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TT_GEMM_DEBUG=1 timeout 300 python tools/profile_target.py --what train --precision bf16 --iters 2 > $O/gemm_timeline.log 2>&1
awk '/tc_gemm</{n++} n>=3' $O/gemm_timeline.log | grep -A12 "last end"
