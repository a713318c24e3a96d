#!/bin/bash
set -u
mkdir -p gpurun_out
for m in 0 1; do
  for args in "32 32 128 4" "64 64 64 4" "32 64 64 4" "64 32 64 4"; do
    NND_TCS_MAP=$m timeout 120 python scripts/profile_conv.py $args fprop 2>&1 | tail -1 | sed "s/^/tcs_map=$m /"
  done
done | tee gpurun_out/r2c9_tcs_map.txt
NND_TCS_MAP=1 timeout 300 python -m pytest tests/test_net_gpu.py -q -k "streaming or conv_norm_relu or whole_network or network_forward" 2>&1 | tail -3
for m in 0 1; do
NND_TCS_MAP=$m timeout 200 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_tcs_kernel -s 3 -c 1 python scripts/profile_conv.py 32 32 128 4 fprop 2>&1 | grep -E "duration|throughput|tensor|xbar|wavefronts|writeback" | sed "s/^/tcs_map=$m /"
done | tee -a gpurun_out/r2c9_tcs_map.txt
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --experimental tcs_map > gpurun_out/r2c9_bench_tcsmap.json 2> gpurun_out/r2c9_bench.err; head -c 300 gpurun_out/r2c9_bench_tcsmap.json; echo
