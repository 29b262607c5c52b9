#!/bin/bash
# launch list of the eager bf16 train step with warm caches (per-kernel durations, kernel timed alone)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/profile_target.py --what train --precision bf16 --iters 3 > $O/plain_train_bf16_3.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/launches_train_bf16_warm.csv \
     python tools/profile_target.py --what train --precision bf16 --iters 3 > $O/ncu_train_bf16_warm.log 2>&1
tail -2 $O/plain_train_bf16_3.log; tail -2 $O/ncu_train_bf16_warm.log
