for cfg in "1024 148" "512 1184" "256 2368"; do
timeout 600 python tools/time_batch.py $cfg 2 2>&1 | grep "hqr=\|solves/s" | tail -2 > gpurun_out/t_leaf.log
echo "cfg=$cfg LEAF=32"; cat gpurun_out/t_leaf.log
done
timeout 300 python tools/bdc_check.py 2>&1 | tail -3
