set -x
LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py 1024 148 2 2>&1 | grep -v "jacobi sweep" > gpurun_out/t_ldh.log
cat gpurun_out/t_ldh.log
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or full_size or known_answer or tiny" > gpurun_out/gpu_tests_ldh.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests_ldh.log
