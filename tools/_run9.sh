set -x
for nw in 32 24 16 40; do
LLCK_AED_NW=$nw LLCK_VERBOSE=1 timeout 600 python tools/time_batch.py 1024 148 1 2>&1 | grep "hqr phase\|solves/s\|hqr=" > gpurun_out/t_nw$nw.log
echo "NW=$nw"; cat gpurun_out/t_nw$nw.log
done
