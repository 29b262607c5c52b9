cd /root/repo
O=gpurun_out
timeout 300 python tools/topk_probe.py 2000000 > $O/r02d_topk_probe.log 2>&1; tail -15 $O/r02d_topk_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_topk_kernel -s 20 -c 1 -f -o $O/r02d_tc_topk python tools/topk_probe.py 2000000 > $O/r02d_ncu_topk.log 2>&1
ncu -i $O/r02d_tc_topk.ncu-rep --page raw --csv > $O/r02d_tc_topk_raw.csv 2>/dev/null
ncu -i $O/r02d_tc_topk.ncu-rep --page source --csv > $O/r02d_tc_topk_source.csv 2>/dev/null
TT_CE_DEBUG=1 timeout 120 python tools/profile_target.py --what train --iters 2 > $O/r02d_ce_timeline_pass1.log 2>&1
TT_CE_DEBUG=0 timeout 120 python tools/profile_target.py --what train --iters 2 > $O/r02d_ce_timeline_pass0.log 2>&1
timeout 300 python tools/time_ops.py > $O/r02d_time_ops.log 2>&1; tail -25 $O/r02d_time_ops.log
