#!/bin/bash
# round-2 GPU run D (1 GPU): full GPU suite with the rough level + default warm start, C3 bench variants, ncu captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_gpu_tests_d.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_d.log
grep -E "passed|failed|rc=|FAILED|Error" gpurun_out/r2_gpu_tests_d.log | tail -20
timeout 900 python bench.py --steps 3 --warmup 3 --verbose 1 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.log
echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_d.json
timeout 600 python bench.py --steps 3 --warmup 2 --verbose 1 --no-warm-start --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_bench_d_cold.json 2> gpurun_out/r2_bench_d_cold.log
echo "bench cold rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_energy_pair_kernel|tc_grad_kernel" -s 8 -c 2 \
    -o gpurun_out/r2_c3_rough -f python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_rough.log 2>&1
echo "ncu rough rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_energy_pair_kernel|tc_grad_kernel" -s 60 -c 2 \
    -o gpurun_out/r2_c3_coarse_b -f python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_coarse_b.log 2>&1
echo "ncu coarse rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_d.csv \
    python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_list_d.log 2>&1
echo "ncu list rc=$?"
