#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tensor_core.py -q --timeout 600 -p no:cacheprovider -x > $O/pytest_tc.log 2>&1; echo "exit $?" >> $O/pytest_tc.log
timeout 600 python tools/time_ops.py --precision bf16 > $O/time_ops_bf16.log 2>&1; echo "exit $?" >> $O/time_ops_bf16.log
timeout 900 python bench.py --steps 30 --warmup 5 --precision bf16 --no-search --no-cpu-baseline > $O/bench_bf16_c.log 2>&1; echo "exit $?" >> $O/bench_bf16_c.log
timeout 300 python tools/profile_target.py --what train --precision bf16 --iters 2 > $O/plain_train_bf16_2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 3 -c 1 -o $O/prof_gemm_da1 python tools/profile_target.py --what train --precision bf16 --iters 2 > $O/ncu_gemm.log 2>&1
tail -3 $O/pytest_tc.log; cat $O/time_ops_bf16.log; tail -c 400 $O/bench_bf16_c.log
