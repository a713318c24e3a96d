#!/bin/bash
# ncu --set full captures of single kernels (one launch each, after warm-up launches) -> gpurun_out/r2_ncu_*.ncu-rep
set -u
mkdir -p gpurun_out
cap() { # name, kernel regex, skip, then the command
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -f -o gpurun_out/r2_ncu_$name "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  tail -2 gpurun_out/r2_ncu_$name.log
}
cap pw32        conv_pw_kernel          2 python scripts/ncu_targets.py pw 32 32 128 4
cap pw_up64     conv_pw_kernel          2 python scripts/ncu_targets.py up 64 32 64 4
NND_PW=0 cap igemm_pw32 conv_igemm_kernel 2 python scripts/ncu_targets.py pw 32 32 128 4
cap wgrad_tc128 conv_wgrad_tc_kernel    2 python scripts/ncu_targets.py block 128 128 32 4
cap conv_tc128  conv_tc_kernel          4 python scripts/ncu_targets.py block 128 128 32 4
cap norm_bwd_reduce norm_bwd_reduce     2 python scripts/ncu_targets.py block 32 32 128 4
cap norm_bwd_apply  norm_bwd_apply      2 python scripts/ncu_targets.py block 32 32 128 4
cap nms_mask    nms_mask_kernel         2 python scripts/ncu_targets.py nms 100000
cap nms_scan    nms_scan_kernel         2 python scripts/ncu_targets.py nms 100000
ls -la gpurun_out/*.ncu-rep
