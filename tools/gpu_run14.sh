#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/profile_target.py --what train --precision bf16 --iters 2 > $O/plain_train_bf16_2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 1 -f -o $O/prof_gemm_da1_b python tools/profile_target.py --what train --precision bf16 --iters 2 > $O/ncu_gemm.log 2>&1
tail -3 $O/ncu_gemm.log
