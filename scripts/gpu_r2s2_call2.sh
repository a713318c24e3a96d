#!/bin/bash
# round 2, session 2, call 2: where does the TMA wgrad kernel spend its time (component timing + ncu --set full)
set -u
mkdir -p gpurun_out
{
timeout 120 python scripts/time_wgrad_tma.py 128 128 32 4 0 1 9 17 33 25 41 2>&1 | grep wgrad
timeout 120 python scripts/time_wgrad_tma.py 64 64 64 4 0 1 9 17 33 25 41 2>&1 | grep wgrad
} | tee gpurun_out/r2s2c2_components.txt
REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tma -s 2 -c 1 -o gpurun_out/r2s2c2_wgrad_tma128 -f python scripts/time_wgrad_tma.py 128 128 32 4 1 > gpurun_out/r2s2c2_ncu128.log 2>&1
REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tma -s 2 -c 1 -o gpurun_out/r2s2c2_wgrad_tma64 -f python scripts/time_wgrad_tma.py 64 64 64 4 1 > gpurun_out/r2s2c2_ncu64.log 2>&1
ls -la gpurun_out | tail -5
