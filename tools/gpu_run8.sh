#!/bin/bash
# PDL on/off comparison: parity tests, per-op timings, train bench.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor_core.py -q --timeout 600 -p no:cacheprovider > $O/pytest_all.log 2>&1; echo "exit $?" >> $O/pytest_all.log
timeout 300 python tools/time_ops.py --precision bf16 > $O/time_ops_bf16.log 2>&1; echo "exit $?" >> $O/time_ops_bf16.log
timeout 600 python bench.py --steps 30 --warmup 5 --precision bf16 --no-search --no-cpu-baseline > $O/bench_bf16_pdl.log 2>&1; echo "exit $?" >> $O/bench_bf16_pdl.log
TT_PDL=0 timeout 600 python bench.py --steps 30 --warmup 5 --precision bf16 --no-search --no-cpu-baseline > $O/bench_bf16_nopdl.log 2>&1; echo "exit $?" >> $O/bench_bf16_nopdl.log
tail -3 $O/pytest_all.log; cat $O/time_ops_bf16.log; tail -c 400 $O/bench_bf16_pdl.log; tail -c 400 $O/bench_bf16_nopdl.log
