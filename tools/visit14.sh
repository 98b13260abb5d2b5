#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== block-scaled probe"
timeout 120 tools/_build/bs_probe > gpurun_out/bs_probe2.txt 2>&1; tail -14 gpurun_out/bs_probe2.txt
echo "== a16w8 tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=600 --tb=short -k "a16w8_weight_only" > gpurun_out/pytest_v14_new.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v14_new.log | cut -c1-330 | head -30
echo "== timing"
timeout 900 python tools/bench_vs_reference.py --only a16w8_int8_M64 --skip-reference --out gpurun_out/vs_reference_timing_v14.json 2>&1 | grep -v "^\[ours\]" | tail -8
