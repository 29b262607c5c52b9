#!/bin/bash
# One-GPU check used with `gpurun`: GPU test-suite, smoke(), default bench + reference arm.  Logs -> gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh r02a [pytest-args...]'
cd "$(dirname "$0")/.."
TAG=${1:-check}; shift
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
# one pytest process per file: a kernel fault (sticky CUDA error) in one file cannot poison the others
: > $O/${TAG}_pytest.log
for f in tests/test_gpu_parity.py tests/test_gpu_tensor_core.py tests/test_gpu_round2.py tests/test_gpu_onepass.py tests/test_plumbing.py tests/test_gpu_multi.py; do
  timeout 900 python -m pytest $f -m gpu -q -s --maxfail=40 "$@" >> $O/${TAG}_pytest.log 2>&1; echo "pytest $f exit $?" | tee -a $O/${TAG}_pytest.log
done
grep -E "^FAILED|passed|failed|error" $O/${TAG}_pytest.log | tail -30 | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"; tail -3 $O/${TAG}_bench.err | cut -c1-300
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref exit $?"
python - <<PY
import json
try:
    l = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.3g pairs/s  %.1f us/step  e2e %.3g  roofline %.3f  launches/step %s" % (l["value"], l["ms_per_step"] * 1e3, l["e2e"]["value"], l["roofline"]["frac"], l["gpu_launches_per_step"]))
    for k in ("word_tower", "msmarco", "search"):
        if k in l: print(k, json.dumps(l[k])[:900])
except Exception as e:
    print("no bench line:", e)
PY
