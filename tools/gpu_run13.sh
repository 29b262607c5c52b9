#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor_core.py -q --timeout 600 -p no:cacheprovider -x > $O/pytest_all.log 2>&1; echo "exit $?" >> $O/pytest_all.log
bash tools/gpu_run12.sh > $O/gemm_tl_summary.log 2>&1
timeout 300 python tools/time_ops.py --precision bf16 > $O/time_ops_bf16.log 2>&1; echo "exit $?" >> $O/time_ops_bf16.log
tail -3 $O/pytest_all.log; grep -E "last end|c[0-3]:|cta   0" $O/gemm_tl_summary.log | head -24; cat $O/time_ops_bf16.log
