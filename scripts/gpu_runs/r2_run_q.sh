#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline_parity.py -m gpu -q -s -p no:cacheprovider -k "default_settings or n200" > gpurun_out/r2_gpu_tests_q.log 2>&1
echo "pytest rc=$?"; grep -E "N=200|passed|failed" gpurun_out/r2_gpu_tests_q.log
