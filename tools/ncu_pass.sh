#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the three heaviest kernels.
set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c2"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:jacobi_step -s 40 -c 2 -f -o gpurun_out/prof_jacobi $CMD > gpurun_out/ncu_jacobi.log 2>&1
echo "jacobi rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hessenberg -c 1 -f -o gpurun_out/prof_hess $CMD > gpurun_out/ncu_hess.log 2>&1
echo "hess rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hqr_kernel -c 1 -f -o gpurun_out/prof_hqr $CMD > gpurun_out/ncu_hqr.log 2>&1
echo "hqr rc=$?"
ls -la gpurun_out/
