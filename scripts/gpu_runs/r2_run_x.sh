#!/bin/bash
# round-2 last check (1 GPU): full GPU suite + smoke on the final commit
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_gpu_tests_x.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_tests_x.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
