#!/bin/bash
# round-2 GPU run H (1 GPU): solver-path tests after dropping the coarse->fine refresh pass + C3 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_headline_parity.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -s --durations=8 -p no:cacheprovider \
    > gpurun_out/r2_gpu_tests_k.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu_tests_k.log
grep -E "passed|failed|rc=|FAILED|Error|rounds" gpurun_out/r2_gpu_tests_k.log | tail -24
timeout 900 python bench.py --steps 3 --warmup 3 --verbose 1 --e2e-steps 4 > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.log
echo "bench rc=$?"; python - <<'PY'
import json
p=json.loads(open('gpurun_out/r2_bench_k.json').read().strip().splitlines()[-1])
print({k:p[k] for k in ('value','ms_per_step','passes','gpu_launches','n_stalled','max_residual')})
r=p['roofline']; print(r['kernel'], r['frac'], r['avg_launch_ms'], r['launches'], r['kernel_ms_share'])
print(p['e2e']['learn_seconds'], p['parity']['max_kkt_violation'], p['clocks'])
PY
grep -E "precision from|rounds so far" gpurun_out/r2_bench_k.log | head -3
