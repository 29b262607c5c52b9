#!/bin/bash
# profiles for the round: launch list of the eager bf16 train step + full ncu capture of the loss backward
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/profile_target.py --what train --precision bf16 --iters 3 > $O/plain_train_bf16_3.log 2>&1 || { tail -5 $O/plain_train_bf16_3.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_train_bf16_final.csv \
   python tools/profile_target.py --what train --precision bf16 --iters 3 > $O/ncu_train_bf16_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_ce_bwd_kernel -s 2 -c 1 -f -o $O/prof_ce_bwd_final \
   python tools/profile_target.py --what train --precision bf16 --iters 3 > $O/ncu_ce_bwd_final.log 2>&1
tail -2 $O/plain_train_bf16_3.log; tail -2 $O/ncu_train_bf16_final.log; tail -2 $O/ncu_ce_bwd_final.log
