#!/bin/bash
# round-2 GPU run E (8 GPUs): the C3 bench as the driver launches it (default partition), then node shards for comparison
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_box_e.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.log
echo "bench 8 rc=$?"; tail -c 1800 gpurun_out/r2_bench_8gpu.json; tail -5 gpurun_out/r2_bench_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 8 --steps 3 --warmup 2 --mode node_sharded --e2e-steps 3 --no-parity > gpurun_out/r2_bench_8gpu_node_sharded.json 2> gpurun_out/r2_bench_8gpu_node_sharded.log
echo "bench 8 node rc=$?"; tail -c 600 gpurun_out/r2_bench_8gpu_node_sharded.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 \
    bench.py --gpus 4 --steps 3 --warmup 2 --e2e-steps 3 --no-parity > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.log
echo "bench 4 rc=$?"; tail -c 600 gpurun_out/r2_bench_4gpu.json
