set -x
LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py 1024 148 1 2>&1 | grep -v "jacobi sweep" > gpurun_out/t_prof_aed.log
cat gpurun_out/t_prof_aed.log
