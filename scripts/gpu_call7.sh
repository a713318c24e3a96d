#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pointwise_tma_gpu.py -q -x > gpurun_out/r2c7_pw_tests.log 2>&1; tail -3 gpurun_out/r2c7_pw_tests.log
for t in "pw 32 32 128 4" "up 64 32 64 4" "pw 64 64 64 4"; do
timeout 200 ncu --metrics gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum --clock-control none -k regex:conv_pw_kernel -s 2 -c 1 python scripts/ncu_targets.py $t 2>&1 | grep -E "duration|dram_thr|sectors" | sed "s/^/$t: /"
done | tee gpurun_out/r2c7_pw_ncu.txt
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2c7_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2c7_pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --trace-layers gpurun_out/r2c7_layers.csv > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; head -c 300 gpurun_out/r2c7_bench.json; echo
