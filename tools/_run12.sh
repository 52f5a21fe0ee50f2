for nw in 16 20; do
LLCK_AED_NW=$nw LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py 512 1184 2 2>&1 | grep "hqr phase\|hqr=" | tail -2 > gpurun_out/t_m512_nw$nw.log
echo "m=512 NW=$nw"; cat gpurun_out/t_m512_nw$nw.log
LLCK_AED_NW=$nw LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py 256 2368 2 2>&1 | grep "hqr phase\|hqr=" | tail -2 > gpurun_out/t_m256_nw$nw.log
echo "m=256 NW=$nw"; cat gpurun_out/t_m256_nw$nw.log
done
