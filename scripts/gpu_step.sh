mkdir -p gpurun_out
timeout 800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
