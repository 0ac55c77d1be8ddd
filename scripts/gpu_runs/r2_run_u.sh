#!/bin/bash
# round-2 GPU run U (1 GPU): wave-aligned pair schedule -- parity tests, bench, ncu capture (DRAM bytes of the energy kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline_parity.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_gpu_tests_u.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_tests_u.log
timeout 600 python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_bench_u.json 2> gpurun_out/r2_bench_u.log; echo "bench rc=$?"
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r2_bench_u.json').read().strip().splitlines()[-1])
print({k:p[k] for k in ('value','ms_per_step','passes')}, p['roofline']['frac'], p['roofline']['avg_launch_ms'], p['roofline']['kernel_ms_share'], p['clocks'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_energy_pair_kernel|tc_grad_kernel" -s 20 -c 2 \
    -o gpurun_out/r2_c3_wave -f python bench.py --steps 1 --warmup 0 --skip-e2e --skip-cpu --no-parity > gpurun_out/r2_ncu_wave.log 2>&1
echo "ncu rc=$?"
