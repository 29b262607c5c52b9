#!/bin/bash
# Multi-GPU check used with `gpurun --gpus N`: the 2-rank parity tests, then bench.py on N ranks.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_multi.sh r02h 2'
cd "$(dirname "$0")/.."
TAG=${1:-multi}; N=${2:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/${TAG}_gpus.txt 2>&1
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -4 $O/${TAG}_pytest_multi.log | cut -c1-300
cp $O/multi_parity.log $O/${TAG}_multi_parity.log 2>/dev/null
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "bench N=$N exit $?"; tail -3 $O/${TAG}_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    l = json.loads(open("$O/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=%d value %.3g pairs/s  %.1f us/step  e2e %.3g  roofline %.3f  local_negatives %.3g" % (l["n_gpus"], l["value"], l["ms_per_step"] * 1e3, l["e2e"]["value"], l["roofline"]["frac"], l.get("local_negatives", {}).get("value", 0)))
    s = l.get("search", {})
    print("search fp32 %.0f QPS  bf16 %.0f QPS" % (s.get("fp32", {}).get("qps", 0), s.get("bf16", {}).get("qps", 0)))
    for k in ("word_tower", "msmarco"):
        if k in l: print(k, json.dumps(l[k])[:600])
except Exception as e:
    print("no bench line:", e)
PY
