mkdir -p gpurun_out
timeout 800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 200 python bench.py --config adam --no-cpu-baseline --steps 6 > gpurun_out/bench_adam.json 2> gpurun_out/bench_adam.err
timeout 200 python bench.py --config lidc --no-cpu-baseline --steps 6 > gpurun_out/bench_lidc.json 2> gpurun_out/bench_lidc.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2150 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_bench.log 2>&1
python -c "from __graft_entry__ import smoke; smoke()" > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cut -c1-200 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_adam.json; cut -c1-200 gpurun_out/bench_lidc.json; tail -2 gpurun_out/smoke.log
