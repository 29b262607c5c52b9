cd /root/repo
O=gpurun_out
timeout 300 python tools/topk_probe.py 10000000 > $O/r02g_topk_probe.log 2>&1; tail -14 $O/r02g_topk_probe.log
TT_CE_DEBUG=1 timeout 120 python tools/profile_target.py --what train --iters 2 > $O/r02g_ce_timeline_pass1.log 2>&1; grep "fused tail" $O/r02g_ce_timeline_pass1.log | tail -2
bash tools/gpu_check.sh r02g
