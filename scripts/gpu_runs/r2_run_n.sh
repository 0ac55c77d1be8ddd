#!/bin/bash
# round-2 GPU run N (2 GPUs): the single-process multi-device tests after the objective-pass fix
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sample_sharded.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_gpu_tests_2gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_2gpu_final.log
tail -6 gpurun_out/r2_gpu_tests_2gpu_final.log
