mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2900 -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
