#!/bin/bash
# round-end style check: smoke, default bench, reference arm
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "exit $?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_default.log 2>&1; echo "exit $?" >> $O/bench_default.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.log 2>&1; echo "exit $?" >> $O/bench_reference.log
tail -3 $O/smoke.log; tail -c 6000 $O/bench_default.log; tail -c 1500 $O/bench_reference.log
