#!/bin/bash
# round 2, session 2, call 1: descriptor-model probe of the TMA-fed wgrad kernel (one variant per process, each under a timeout)
set -u
mkdir -p gpurun_out
for m in 3 1 7 5; do
  timeout 150 python scripts/probe_wgrad_tma.py $m 2>&1 | grep -v Warning | tail -20
  echo "exit mode $m: $?"
done | tee gpurun_out/r2s2c1_probe.txt
