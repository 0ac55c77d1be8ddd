#!/bin/bash
# round-2 GPU run C (1 GPU): full GPU suite, C3 bench cold / warm start, the other configs, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_gpu_tests_c.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_c.log
grep -E "passed|failed|rc=|FAILED|Error" gpurun_out/r2_gpu_tests_c.log | tail -20
timeout 900 python bench.py --steps 3 --warmup 3 --verbose 1 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.log
echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_c.json
timeout 600 python bench.py --steps 3 --warmup 2 --verbose 1 --warm-start --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_bench_c_warm.json 2> gpurun_out/r2_bench_c_warm.log
echo "bench warm rc=$?"; tail -c 300 gpurun_out/r2_bench_c_warm.json
for cfg in c0 c1 c2 c4; do
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 2 > gpurun_out/r2_bench_$cfg.json 2> gpurun_out/r2_bench_$cfg.log
  echo "bench $cfg rc=$?"; tail -c 400 gpurun_out/r2_bench_$cfg.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 220 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_energy_pair_kernel|tc_grad_kernel" -s 6 -c 2 \
    -o gpurun_out/r2_c3_coarse -f python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 900 python scripts/bench_matrix_input.py 1e7 1000 f64 > gpurun_out/r2_matrix_input_f64.json 2> gpurun_out/r2_matrix_input_f64.log
echo "matrix rc=$?"; cat gpurun_out/r2_matrix_input_f64.json
