#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pointwise_tma_gpu.py -q -x > gpurun_out/r2c5_pw_tests.log 2>&1; tail -4 gpurun_out/r2c5_pw_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c5_bench_full.json 2> gpurun_out/r2c5_bench.err; head -c 300 gpurun_out/r2c5_bench_full.json; echo
cap() { local name=$1 rx=$2 skip=$3; shift 3
  timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c 1 -f -o gpurun_out/r2_ncu_$name "$@" > gpurun_out/r2_ncu_$name.log 2>&1; tail -1 gpurun_out/r2_ncu_$name.log; }
cap pw32_v2    conv_pw_kernel 2 python scripts/ncu_targets.py pw 32 32 128 4
cap pw_up64_v2 conv_pw_kernel 2 python scripts/ncu_targets.py up 64 32 64 4
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c5_sanitizer.log 2>&1; tail -8 gpurun_out/r2c5_sanitizer.log
