#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_multi.py -q --timeout 280 -p no:cacheprovider > $O/pytest_multi.log 2>&1; echo "exit $?" >> $O/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > $O/bench_n$N.log 2>&1; echo "exit $?" >> $O/bench_n$N.log
tail -5 $O/pytest_multi.log | cut -c1-300
tail -c 2500 $O/bench_n$N.log
