set -x
TAG=r02b
CMD="python bench.py --steps 1 --warmup 1 --members 148 --e2e-steps 1 --no-cpu-baseline --no-configs --no-weak"
capd() { timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_${TAG}_$1 $CMD > gpurun_out/ncu_${TAG}_$1.log 2>&1; echo "$1 rc=$?"; }
capd hankel "zgemm_batched_kernel<.int.2" 0
capd rankk "zgemm_rankk_kernel<.int.32, .bool.0" 20
python tools/ncu_summary.py gpurun_out/ncu_summary_$TAG.md gpurun_out/prof_${TAG}_*.ncu-rep
