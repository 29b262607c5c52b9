cd /root/repo
O=gpurun_out
bash tools/gpu_profile.sh r02 tc_ce_bwd_kernel tc_ce_fwd_kernel tc_mlp_fwd_kernel
TT_PROFILE_WHAT=word bash tools/gpu_profile.sh r02 embed_pool_fwd_kernel seg_reduce_vec_kernel
TT_PROFILE_WHAT=search_batched bash tools/gpu_profile.sh r02 tc_topk_kernel
ncu -i $O/r02_tc_topk_kernel.ncu-rep --page source --csv > $O/r02_tc_topk_kernel_source.csv 2>/dev/null
TT_PROFILE_WHAT=search bash tools/gpu_profile.sh r02 scan_topk_kernel
ls -la $O | grep "r02_" | head -40
