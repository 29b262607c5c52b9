cd /root/repo
O=gpurun_out
timeout 300 python tools/topk_probe.py 10000000 > $O/r02j_topk_probe.log 2>&1; tail -14 $O/r02j_topk_probe.log
bash tools/gpu_check.sh r02j
