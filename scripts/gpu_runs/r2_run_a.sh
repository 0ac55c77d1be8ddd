#!/bin/bash
# round-2 GPU run A: full GPU test-suite + a first 1-GPU bench line
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; } > gpurun_out/r2_box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_gpu_tests_a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_a.log
tail -40 gpurun_out/r2_gpu_tests_a.log
timeout 900 python bench.py --steps 2 --warmup 2 --verbose 1 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.log
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench_a.json
