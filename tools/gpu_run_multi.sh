#!/bin/bash
# usage: bash tools/gpu_run_multi.sh N   (run under gpurun --gpus N)
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/multi_gpus.log 2>&1
timeout 400 python -m pytest tests/test_gpu_multi.py -q --timeout 380 -p no:cacheprovider > $O/pytest_multi.log 2>&1; echo "exit $?" >> $O/pytest_multi.log
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_n1.log 2>&1; echo "exit $?" >> $O/bench_n1.log
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 > $O/bench_n$n.log 2>&1; echo "exit $?" >> $O/bench_n$n.log
  fi
done
tail -5 $O/pytest_multi.log
for n in 1 $N; do tail -c 1500 $O/bench_n$n.log; echo; done
