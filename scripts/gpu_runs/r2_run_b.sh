#!/bin/bash
# round-2 GPU run B (2 GPUs): sample-sharded / multi-device tests, then the C3 bench at N=2 in both partitions
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_box_b.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sample_sharded.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_gpu_tests_2gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_2gpu_final.log
tail -25 gpurun_out/r2_gpu_tests_2gpu_final.log
for mode in sample_sharded node_sharded; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 2 --warmup 2 --mode $mode --e2e-steps 3 --no-parity > gpurun_out/r2_bench_2gpu_final_$mode.json 2> gpurun_out/r2_bench_2gpu_final_$mode.log
  echo "bench $mode rc=$?"
  tail -c 1500 gpurun_out/r2_bench_2gpu_final_$mode.json
  tail -5 gpurun_out/r2_bench_2gpu_final_$mode.log
done
