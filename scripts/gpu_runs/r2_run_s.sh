#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_matrix_input.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_gpu_tests_s.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_gpu_tests_s.log
GML_B200_NO_FUSED_NEWTON=1 timeout 300 python bench.py --config c0 --steps 5 --warmup 3 | python -c "import json,sys; p=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c0 unfused', p['value'], p['e2e']['value'], p['gpu_launches'])"
for cfg in c0 c1; do timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r2_bench_s_$cfg.json; python -c "import json,sys; p=json.loads(open('gpurun_out/r2_bench_s_$cfg.json').read().strip().splitlines()[-1]); print('$cfg fused', p['value'], p['e2e']['value'], p['gpu_launches'], p['passes'])"; done
python -c "import __graft_entry__ as g; g.smoke()"
