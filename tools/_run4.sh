set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r01d.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/gpu_tests_r01d.log
timeout 600 python tools/time_batch.py 1024 148 2 2>&1 | grep -v "jacobi sweep" > gpurun_out/t_w32.log
cat gpurun_out/t_w32.log
timeout 900 python tools/configs_check.py > gpurun_out/configs_r01d.log 2>&1; tail -8 gpurun_out/configs_r01d.log
