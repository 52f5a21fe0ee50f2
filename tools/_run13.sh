for nb in 30 60 1000; do
for cfg in "1024 148" "512 1184" "256 2368"; do
LLCK_AED_NIBBLE=$nb LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py $cfg 2 2>&1 | grep "hqr phase\|hqr=" | tail -2 > gpurun_out/t_nib.log
echo "cfg=$cfg NIBBLE=$nb"; cat gpurun_out/t_nib.log
done
done
