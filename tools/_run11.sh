set -x
LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py 512 1184 2 2>&1 | grep -v "jacobi sweep" > gpurun_out/t_m512.log
cat gpurun_out/t_m512.log
LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py 256 2368 2 2>&1 | grep -v "jacobi sweep" > gpurun_out/t_m256.log
cat gpurun_out/t_m256.log
