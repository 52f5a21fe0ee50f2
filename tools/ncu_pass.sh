#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the heaviest kernels.
set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c2"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hqr_kernel -c 1 -f -o gpurun_out/prof2_hqr $CMD > gpurun_out/ncu_hqr.log 2>&1
echo "hqr rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rjacobi_update -s 40 -c 1 -f -o gpurun_out/prof2_rupd $CMD > gpurun_out/ncu_rupd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rjacobi_eig -s 40 -c 1 -f -o gpurun_out/prof2_reig $CMD > gpurun_out/ncu_reig.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rjacobi_gram -s 40 -c 1 -f -o gpurun_out/prof2_rgram $CMD > gpurun_out/ncu_rgram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bidiag_panel -s 8 -c 1 -f -o gpurun_out/prof2_bidiag $CMD > gpurun_out/ncu_bidiag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hess_panel -s 8 -c 1 -f -o gpurun_out/prof2_hess $CMD > gpurun_out/ncu_hess.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:zgemm_batched_kernel<2" -c 1 -f -o gpurun_out/prof2_hankel $CMD > gpurun_out/ncu_hankel.log 2>&1
echo "done rc=$?"
ls -la gpurun_out/*.ncu-rep
