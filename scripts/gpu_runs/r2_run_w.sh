#!/bin/bash
# round-2 GPU run W (1 GPU): blocked Gauss-Jordan of the mean-field start -- warm-start tests + solver-only bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline_parity.py -m gpu -q -s -p no:cacheprovider -k "warm_start or n200 or c2_fixture" > gpurun_out/r2_gpu_tests_w.log 2>&1
echo "pytest rc=$?"; grep -E "rounds|passed|failed" gpurun_out/r2_gpu_tests_w.log | tail -12
timeout 600 python bench.py --steps 3 --warmup 2 --skip-e2e --skip-cpu --no-parity --verbose 1 > gpurun_out/r2_bench_w.json 2> gpurun_out/r2_bench_w.log; echo "bench rc=$?"
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r2_bench_w.json').read().strip().splitlines()[-1])
print({k:p[k] for k in ('value','ms_per_step','passes','gpu_launches')}, p['roofline']['frac'], p['max_abs_coupling_error_vs_truth'], p['clocks'])
PY
grep -E "warm|precision from|setup" gpurun_out/r2_bench_w.log | head -4
