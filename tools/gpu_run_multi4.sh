#!/bin/bash
# N-GPU bench, NCCL collectives vs NVLink peer-memory exchanges
cd "$(dirname "$0")/.."
N=${1:-8}
O=gpurun_out
mkdir -p $O
for P in 0 1; do
TT_P2P=$P timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$P bench.py --gpus $N --steps 30 --warmup 5 --no-search > $O/bench_n${N}_p2p$P.log 2>&1; echo "exit $?" >> $O/bench_n${N}_p2p$P.log
python - $O/bench_n${N}_p2p$P.log <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        j=json.loads(l); ok=True
        print(sys.argv[1], 'value', round(j['value']/1e6,2), 'M pairs/s', round(j['ms_per_step']*1e3,1), 'us | e2e', round(j['e2e']['value']/1e6,2), '| local', round(j.get('local_negatives',{}).get('value',0)/1e6,2), round(j.get('local_negatives',{}).get('ms_per_step',0)*1e3,1), 'us | launches', j['gpu_launches_per_step'])
if not ok: print(open(sys.argv[1]).read()[-1500:])
PY
done
