#!/bin/bash
# Round-end check on one fresh B200: the whole GPU suite, smoke(), the default bench line, the reference arm, the ncu launch list.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; head -c 250 gpurun_out/r2f_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2>> gpurun_out/r2f_bench.err; head -c 250 gpurun_out/r2f_bench_ref.json; echo
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1100 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/r2f_profile.log 2>&1
tail -2 gpurun_out/r2f_profile.log; tail -3 gpurun_out/r2f_bench.err
