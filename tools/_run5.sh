set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_r01d.json 2> gpurun_out/bench_2gpu_r01d.err; echo "rc=$?"
cat gpurun_out/bench_2gpu_r01d.json | cut -c1-1500; tail -3 gpurun_out/bench_2gpu_r01d.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_2gpu_r01d.json 2>&1; echo "rc=$?"
tail -2 gpurun_out/bench_ref_2gpu_r01d.json | cut -c1-800
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_check.py > gpurun_out/dist_check_r01d.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/dist_check_r01d.log
