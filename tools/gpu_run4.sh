#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor_core.py -q --timeout 600 -p no:cacheprovider -x > $O/pytest_all.log 2>&1; echo "exit $?" >> $O/pytest_all.log
timeout 900 python bench.py --steps 30 --warmup 5 --precision bf16 --no-search --no-cpu-baseline > $O/bench_bf16_c.log 2>&1; echo "exit $?" >> $O/bench_bf16_c.log
timeout 300 python tools/profile_target.py --what train --precision bf16 > $O/plain_train_bf16.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_train_bf16_c.csv python tools/profile_target.py --what train --precision bf16 > $O/ncu_train_bf16.log 2>&1
timeout 300 python tools/profile_target.py --what train --precision bf16 --iters 2 > $O/plain_train_bf16_2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_ce_bwd -c 2 -o $O/prof_ce_bwd python tools/profile_target.py --what train --precision bf16 --iters 2 > $O/ncu_ce_bwd.log 2>&1
tail -5 $O/pytest_all.log; tail -c 700 $O/bench_bf16_c.log
