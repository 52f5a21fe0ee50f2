set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r01e.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/gpu_tests_r01e.log
timeout 1500 python tools/c2_full.py 100 > gpurun_out/c2_full.log 2>&1; echo "c2 rc=$?"; grep -v Warning gpurun_out/c2_full.log | tail -4
