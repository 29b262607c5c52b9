#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tensor_core.py -q -k test_mlp_bf16_vs_oracle --timeout 300 -p no:cacheprovider > $O/pytest_tc_mlp.log 2>&1; echo "exit $?" >> $O/pytest_tc_mlp.log
timeout 300 python tools/profile_target.py --what train --precision bf16 > $O/plain_train_bf16.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_train_bf16.csv python tools/profile_target.py --what train --precision bf16 > $O/ncu_train_bf16.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 --precision bf16 --no-search --no-cpu-baseline > $O/bench_bf16_b.log 2>&1; echo "exit $?" >> $O/bench_bf16_b.log
tail -3 $O/pytest_tc_mlp.log; tail -c 900 $O/bench_bf16_b.log
