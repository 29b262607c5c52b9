#!/bin/bash
# GPU session 2: parity (fp32), tensor-core tests (isolated processes), bench, ncu
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider > $O/pytest_parity.log 2>&1; echo "exit $?" >> $O/pytest_parity.log
for t in test_tc_gemm_selftest test_mlp_bf16_vs_oracle test_inbatch_ce_bf16_vs_oracle test_inbatch_bf16_full_size_known_answers test_fused_trainer_bf16_tracks_fp32; do
  timeout 600 python -m pytest tests/test_gpu_tensor_core.py -q -k $t --timeout 300 -p no:cacheprovider > $O/pytest_tc_$t.log 2>&1; echo "exit $?" >> $O/pytest_tc_$t.log
done
timeout 900 python bench.py --steps 30 --warmup 5 --precision fp32 > $O/bench_fp32.log 2>&1; echo "exit $?" >> $O/bench_fp32.log
timeout 900 python bench.py --steps 30 --warmup 5 --precision bf16 --no-search > $O/bench_bf16.log 2>&1; echo "exit $?" >> $O/bench_bf16.log
# ncu: launch list of an eager train step (fp32 path, known good) + full capture of the scan kernel
timeout 300 python tools/profile_target.py --what train --precision fp32 > $O/plain_train_fp32.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_train_fp32.csv python tools/profile_target.py --what train --precision fp32 > $O/ncu_train_fp32.log 2>&1
timeout 300 python tools/profile_target.py --what search --dtype fp32 --iters 2 > $O/plain_search.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -c 2 -o $O/prof_scan_fp32 python tools/profile_target.py --what search --dtype fp32 --iters 2 > $O/ncu_search.log 2>&1
for f in $O/pytest_*.log; do echo "== $f"; tail -4 $f; done
tail -c 600 $O/bench_bf16.log
