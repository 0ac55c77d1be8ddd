#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_r.json 2> gpurun_out/r2_bench_r.log
echo "bench rc=$?"; python - <<'PY'
import json
p=json.loads(open('gpurun_out/r2_bench_r.json').read().strip().splitlines()[-1])
print({k:p[k] for k in ('value','ms_per_step','passes')}, p['e2e']['learn_seconds'], p['e2e_from_matrix'], p['clocks'])
PY
tail -3 gpurun_out/r2_bench_r.log
