set -x
timeout 600 python tools/bdc_check.py > gpurun_out/bdc_check.log 2>&1; echo "bdc_check rc=$?"
tail -45 gpurun_out/bdc_check.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_dc.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/gpu_tests_dc.log
LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py 1024 148 2 2>&1 | grep -v "jacobi sweep" > gpurun_out/t_dc.log
cat gpurun_out/t_dc.log
