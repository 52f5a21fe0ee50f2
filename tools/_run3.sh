set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r01c.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/gpu_tests_r01c.log
timeout 900 python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; echo "bench rc=$?"
cat gpurun_out/bench_r01c.json; tail -3 gpurun_out/bench_r01c.err
bash tools/ncu_pass.sh r01c
