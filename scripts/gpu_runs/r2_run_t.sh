#!/bin/bash
# round-2 final check (1 GPU): what the driver runs at round end -- GPU tests, smoke, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_gpu_tests_t.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_tests_t.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r2_bench_t.json 2> gpurun_out/r2_bench_t.log; echo "bench rc=$?"
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r2_bench_t.json').read().strip().splitlines()[-1])
print({k:p[k] for k in ('value','ms_per_step','steps','warmup','passes')}, p['e2e']['learn_seconds'], p['e2e_from_matrix'].get('learn_seconds'), p['roofline']['frac'], p['parity']['max_kkt_violation'], p['clocks'])
PY
