set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r01f.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/gpu_tests_r01f.log
timeout 900 python bench.py > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; echo "bench rc=$?"
cat gpurun_out/bench_r01f.json | cut -c1-400
bash tools/ncu_pass.sh r01f full
