#!/bin/bash
# round-end rehearsal on one GPU: full gpu test-suite, smoke, default bench (+ reference arm), per-op times
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu --timeout 600 -p no:cacheprovider > $O/pytest_gpu_final.log 2>&1; echo "exit $?" >> $O/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "exit $?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench_default.log 2>&1; echo "exit $?" >> $O/bench_default.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference.log 2>&1; echo "exit $?" >> $O/bench_reference.log
timeout 300 python tools/time_ops.py --precision bf16 > $O/time_ops_bf16_final.log 2>&1; echo "exit $?" >> $O/time_ops_bf16_final.log
tail -3 $O/pytest_gpu_final.log; tail -2 $O/smoke.log; tail -c 1200 $O/bench_reference.log | head -c 400; echo; cat $O/time_ops_bf16_final.log | grep -E "whole|prefix|repeat"
python - <<'PY'
import json
for l in open('gpurun_out/bench_default.log'):
    if l.startswith('{'):
        j=json.loads(l)
        print('value', round(j['value']/1e6,2), 'M pairs/s', round(j['ms_per_step']*1e3,1), 'us | e2e', round(j['e2e']['value']/1e6,2), '| roofline', round(j['roofline']['frac'],3), round(j['roofline']['ms']*1e3,1), 'us', j['roofline']['traffic'], '| cpu', round(j['cpu_baseline']['value']), '| search', round(j['search']['fp32']['qps']), round(j['search']['bf16']['qps']), 'e2e', round(j['search']['e2e']['value']), '| clocks', j['clocks'])
PY
