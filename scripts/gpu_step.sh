mkdir -p gpurun_out
timeout 800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_refactor.json 2> gpurun_out/bench.err
tail -4 gpurun_out/pytest_gpu.log; cut -c1-200 gpurun_out/bench_refactor.json
