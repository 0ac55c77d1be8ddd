#!/bin/bash
# round-2 GPU run L (1 GPU): final code -- full GPU suite, smoke, C3 bench, the other configs, launch list, ncu full captures, matrix input
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s --durations=10 -p no:cacheprovider > gpurun_out/r2_gpu_tests_l.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_l.log
grep -E "passed|failed|rc=|FAILED|Error" gpurun_out/r2_gpu_tests_l.log | tail -12
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_l.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke_l.log
timeout 900 python bench.py --steps 3 --warmup 3 --verbose 1 > gpurun_out/r2_bench_l.json 2> gpurun_out/r2_bench_l.log
echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_l.json
for cfg in c0 c1 c2 c4; do
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 2 > gpurun_out/r2_bench_l_$cfg.json 2> gpurun_out/r2_bench_l_$cfg.log
  echo "bench $cfg rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches_l.csv \
    python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_list_l.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_energy_pair_kernel|tc_grad_kernel" -s 20 -c 2 \
    -o gpurun_out/r2_c3_final_coarse -f python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_final_coarse.log 2>&1
echo "ncu coarse rc=$?"
timeout 900 python scripts/bench_matrix_input.py 1e7 1000 f64 > gpurun_out/r2_matrix_input_l.json 2> gpurun_out/r2_matrix_input_l.log
echo "matrix rc=$?"; cat gpurun_out/r2_matrix_input_l.json
