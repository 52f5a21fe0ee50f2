#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command (+ with "full": --set full captures of the heaviest kernels).
# usage: bash tools/ncu_pass.sh TAG [full]      (one GPU; never under torchrun)
set -x
TAG=${1:-r02a}
CMD="python bench.py --steps 1 --warmup 1 --members 148 --e2e-steps 1 --no-cpu-baseline --no-configs --no-weak"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "list rc=$?"
if [ "$2" = "full" ]; then
  cap() {   # name regex skip
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_${TAG}_$1 $CMD > gpurun_out/ncu_${TAG}_$1.log 2>&1
    echo "$1 rc=$?"
  }
  cap hqr hqr_kernel 0
  cap bidiag bidiag_panel 8
  cap hess hess_panel 8
  capd() {  # like cap, but the regex is matched against the demangled name (template arguments)
    timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_${TAG}_$1 $CMD > gpurun_out/ncu_${TAG}_$1.log 2>&1
    echo "$1 rc=$?"
  }
  capd hankel "zgemm_batched_kernel<.int.2" 0
  capd rankk "zgemm_rankk_kernel<.int.32, .bool.0" 20
  cap bdcgemm bdc_gemm 4
  ls -la gpurun_out/*_${TAG}_*.ncu-rep
  python tools/ncu_summary.py gpurun_out/ncu_summary_$TAG.md gpurun_out/prof_${TAG}_*.ncu-rep
fi
echo "done"
