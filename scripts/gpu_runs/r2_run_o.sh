#!/bin/bash
# round-2 GPU run O (8 GPUs): the C3 bench as the driver launches it, final code
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 5 > gpurun_out/r2_bench_8gpu_final.json 2> gpurun_out/r2_bench_8gpu_final.log
echo "bench 8 rc=$?"; tail -c 900 gpurun_out/r2_bench_8gpu_final.json
