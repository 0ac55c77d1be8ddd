#!/bin/bash
# round-2 GPU run F (1 GPU): full GPU suite on the final code, C3 bench, launch list, ncu full capture
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_gpu_tests_f.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_f.log
grep -E "passed|failed|rc=|FAILED|Error" gpurun_out/r2_gpu_tests_f.log | tail -20
timeout 900 python bench.py --steps 3 --warmup 3 --verbose 1 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.log
echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_f.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches_f.csv \
    python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_list_f.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_energy_pair_kernel|tc_grad_kernel" -s 8 -c 2 \
    -o gpurun_out/r2_c3_early -f python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_early.log 2>&1
echo "ncu early rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_energy_pair_kernel|tc_grad_kernel" -s 50 -c 2 \
    -o gpurun_out/r2_c3_late -f python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_late.log 2>&1
echo "ncu late rc=$?"
