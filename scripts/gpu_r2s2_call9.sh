#!/bin/bash
# round 2, session 2, call 9: evidence -- ncu --set full of the four TMA-fed kernels, the full default bench line, the step's launch list
set -u
mkdir -p gpurun_out
cap() { # name, kernel regex, skip, then the command
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -f -o gpurun_out/r2s2_ncu_$name "$@" > gpurun_out/r2s2_ncu_$name.log 2>&1
  tail -1 gpurun_out/r2s2_ncu_$name.log
}
cap wgrad_tma128   conv_wgrad_tma_kernel     2 python scripts/ncu_targets.py block 128 128 32 4
cap wgrad_tma64    conv_wgrad_tma_kernel     2 python scripts/ncu_targets.py block 64 64 64 4
cap conv_tct128    conv_tct_kernel           4 python scripts/ncu_targets.py block 128 128 32 4
cap conv_tct_s2_32 conv_tct_kernel           2 python scripts/ncu_targets.py block2 32 64 128 4
cap wgrad_tma_s2   conv_wgrad_tma_s2_kernel  2 python scripts/ncu_targets.py block2 32 64 128 4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s2c9_bench.json 2> gpurun_out/r2s2c9_bench.err; head -c 250 gpurun_out/r2s2c9_bench.json; echo
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1200 --csv --log-file gpurun_out/r2s2c9_launches.csv \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/r2s2c9_profile.log 2>&1
tail -2 gpurun_out/r2s2c9_profile.log; ls -la gpurun_out/*r2s2_ncu*.ncu-rep
