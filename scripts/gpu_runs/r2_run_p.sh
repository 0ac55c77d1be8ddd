#!/bin/bash
# round-2 GPU run P (1 GPU): compute-sanitizer memcheck over the round-2 device paths + final full suite
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/dev_sanitize.py > gpurun_out/r2_sanitizer.log 2>&1
echo "sanitizer rc=$?"; tail -6 gpurun_out/r2_sanitizer.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_gpu_tests_p.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_tests_p.log
