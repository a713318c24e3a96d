#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pointwise_tma_gpu.py -q -x > gpurun_out/r2c6_pw_tests.log 2>&1; tail -3 gpurun_out/r2c6_pw_tests.log
for m in 0 1; do
  for args in "128 128 32 4" "64 64 64 4" "256 256 16 4"; do
    NND_WG_MAP=$m timeout 120 python scripts/profile_conv.py $args wgrad 2>&1 | tail -1 | sed "s/^/map=$m /"
  done
done | tee gpurun_out/r2c6_wgrad_map.txt
for m in 0 1; do
  NND_WG_MAP=$m timeout 200 ncu --metrics l1tex__m_xbar2l1tex_read_bytes.sum,gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv_wgrad_tc_kernel -s 3 -c 1 python scripts/profile_conv.py 128 128 32 4 wgrad 2>&1 | grep -E "xbar|duration|lts__|tensor" | sed "s/^/map=$m /"
done | tee -a gpurun_out/r2c6_wgrad_map.txt
NND_WG_MAP=0 timeout 120 python -m pytest tests/test_net_gpu.py -q -k "tcgen05_vs_mma_sync_single_layer or all_taps_wgrad" 2>&1 | tail -2
NND_WG_MAP=1 timeout 120 python -m pytest tests/test_net_gpu.py -q -k "tcgen05_vs_mma_sync_single_layer or all_taps_wgrad" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --trace-layers gpurun_out/r2c6_layers.csv > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err; head -c 300 gpurun_out/r2c6_bench.json; echo
python scripts/ncu_targets.py pw 32 32 128 4 > /dev/null
timeout 200 ncu --metrics gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_pw_kernel -s 2 -c 1 python scripts/ncu_targets.py pw 32 32 128 4 2>&1 | grep -E "duration|dram_thr"
timeout 200 ncu --metrics gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_pw_kernel -s 2 -c 1 python scripts/ncu_targets.py up 64 32 64 4 2>&1 | grep -E "duration|dram_thr"
