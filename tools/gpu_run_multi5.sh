#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_multi.py -q --timeout 280 -p no:cacheprovider > $O/pytest_multi.log 2>&1; echo "exit $?" >> $O/pytest_multi.log
tail -4 $O/pytest_multi.log | cut -c1-300
for P in 0 1; do
TT_P2P=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$P bench.py --gpus $N --steps 30 --warmup 5 > $O/bench_n${N}_p2p$P.log 2>&1; echo "exit $?" >> $O/bench_n${N}_p2p$P.log
python - $O/bench_n${N}_p2p$P.log <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j=json.loads(l); ok=True
        print(sys.argv[1], 'value', round(j['value']/1e6,2), 'M pairs/s', round(j['ms_per_step']*1e3,1), 'us | e2e', round(j['e2e']['value']/1e6,2), '| local', round(j.get('local_negatives',{}).get('value',0)/1e6,2), '| roofline frac', round(j['roofline']['frac'],3), round(j['roofline']['ms']*1e3,1), 'us | search', round(j['search']['fp32']['qps']), round(j['search']['bf16']['qps']))
if not ok: print(open(sys.argv[1]).read()[-1500:])
PY
done
