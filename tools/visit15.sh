#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== mx dynamic tests (block-scaled kernel)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=600 --tb=short -k "mx_dynamic or fixtures or mx_activation" > gpurun_out/pytest_v15_new.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v15_new.log | cut -c1-400 | head -30
echo "== vs reference (mx cases)"
timeout 900 python -m pytest tests/test_gpu_vs_reference.py -m gpu -q --timeout=800 --tb=short -k "mxfp or nvfp" > gpurun_out/pytest_v15_vsref.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v15_vsref.log | cut -c1-400 | head -30
echo "== timing"
timeout 900 python tools/bench_vs_reference.py --only mxdyn --budget 400 --out gpurun_out/vs_reference_timing_v15.json 2>&1 | grep -v "^\[ours\]" | tail -30
