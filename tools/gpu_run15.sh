#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -q --timeout 120 -p no:cacheprovider -x -k "fused_normalise" > $O/pytest_fused.log 2>&1; echo "exit $?" >> $O/pytest_fused.log
tail -30 $O/pytest_fused.log
