#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== new pair-kernel formats"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=600 --tb=short -k "a16w8_weight_only or packed_8_2_1 or fixtures" > gpurun_out/pytest_v12_new.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v12_new.log | cut -c1-330 | head -30
echo "== vs reference (a16w8 / hqq cases)"
timeout 900 python -m pytest tests/test_gpu_vs_reference.py -m gpu -q --timeout=800 --tb=short -k "a16w8 or hqq or channelwise or bitnet" > gpurun_out/pytest_v12_vsref.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v12_vsref.log | cut -c1-330 | head -30
echo "== timing"
timeout 900 python tools/bench_vs_reference.py --only a16w8,a16w2 --budget 500 --out gpurun_out/vs_reference_timing_v12.json 2>&1 | grep -v "^\[ours\]" | tail -40
