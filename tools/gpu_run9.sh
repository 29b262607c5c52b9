#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/time_ops.py --precision bf16 --flush write_read > $O/time_ops_wr.log 2>&1; echo "exit $?" >> $O/time_ops_wr.log
cat $O/time_ops_wr.log
