#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 3 --warmup 2 --e2e-steps 3 --no-parity > gpurun_out/r2_bench_2gpu_v.json 2> gpurun_out/r2_bench_2gpu_v.log
echo "bench rc=$?"
python - <<'PY'
import json
p=json.loads(open('gpurun_out/r2_bench_2gpu_v.json').read().strip().splitlines()[-1])
print({k:p[k] for k in ('n_gpus','value','ms_per_step','passes')}, p['e2e']['learn_seconds'], p['multi_gpu_check'], p['roofline']['frac'], p['clocks'])
PY
