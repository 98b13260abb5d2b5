#!/bin/bash
# round 2, first GPU visit: opt-in validation, live reference parity, reference timing
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
bash tools/validate_optin.sh 2>&1 | tee gpurun_out/optin_summary.txt
echo "== vs reference (parity)"
GEMLITE_B200_EXPERIMENTAL=1 timeout 1700 python -m pytest tests/test_gpu_vs_reference.py -m gpu -q --timeout=1600 --tb=line > gpurun_out/vsref_pytest_all.log 2>&1
tail -40 gpurun_out/vsref_pytest_all.log | cut -c1-400
tail -30 gpurun_out/vs_reference_child.log | cut -c1-300
echo "== vs reference (timing)"
timeout 1500 python tools/bench_vs_reference.py --budget 800 > gpurun_out/vsref_timing.log 2>&1
tail -60 gpurun_out/vsref_timing.log | cut -c1-250
