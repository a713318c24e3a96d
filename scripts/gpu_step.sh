mkdir -p gpurun_out
./scripts/mma_rate > gpurun_out/mma_rate3.txt 2>&1
timeout 800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_bench.log 2>&1
tail -6 gpurun_out/pytest_gpu.log; grep -E "^1 [0-3] +(32|96|128) " gpurun_out/mma_rate3.txt; cat gpurun_out/bench.json
