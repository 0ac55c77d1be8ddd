#!/bin/bash
# round-2 GPU run G (1 GPU): full GPU suite on the final code (with durations), smoke, C3 bench, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --durations=25 -p no:cacheprovider > gpurun_out/r2_gpu_tests_g.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_g.log
grep -E "passed|failed|rc=|FAILED|Error" gpurun_out/r2_gpu_tests_g.log | tail -12
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_g.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke_g.log
timeout 900 python bench.py --steps 3 --warmup 3 --verbose 1 > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.log
echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_g.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches_g.csv \
    python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_list_g.log 2>&1
echo "ncu list rc=$?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.log
echo "reference rc=$?"; tail -c 300 gpurun_out/r2_bench_reference.json
