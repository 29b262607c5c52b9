#!/bin/bash
# ncu evidence used with `gpurun` (one GPU): launch list of the eager train step + one `--set full` capture per kernel regex.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r02 tc_ce_bwd_kernel embed_pool_fwd_kernel'
cd "$(dirname "$0")/.."
TAG=${1:-prof}; shift
O=gpurun_out; mkdir -p $O
WHAT=${TT_PROFILE_WHAT:-train}
timeout 300 python tools/profile_target.py --what $WHAT --iters 3 > $O/${TAG}_plain_${WHAT}.log 2>&1 || { tail -5 $O/${TAG}_plain_${WHAT}.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_${WHAT}.csv \
   python tools/profile_target.py --what $WHAT --iters 3 > $O/${TAG}_ncu_launches_${WHAT}.log 2>&1
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o $O/${TAG}_${K} \
     python tools/profile_target.py --what $WHAT --iters 3 > $O/${TAG}_ncu_${K}.log 2>&1
  ncu -i $O/${TAG}_${K}.ncu-rep --page raw --csv > $O/${TAG}_${K}_raw.csv 2>/dev/null
  tail -1 $O/${TAG}_ncu_${K}.log
done
