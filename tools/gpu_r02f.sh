cd /root/repo
O=gpurun_out
timeout 300 python tools/topk_probe.py 10000000 > $O/r02f_topk_probe.log 2>&1; tail -14 $O/r02f_topk_probe.log
timeout 300 python tools/time_ops.py > $O/r02f_time_ops.log 2>&1; tail -22 $O/r02f_time_ops.log
bash tools/gpu_check.sh r02f
