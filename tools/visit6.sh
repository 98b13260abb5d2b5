#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== mid tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=600 --tb=short -k "batched_decode" > gpurun_out/pytest_v6_new.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v6_new.log | cut -c1-330 | head -30
echo "== mid-M timing (ours)"
timeout 300 python tools/bench_vs_reference.py --only c2_M,c4_int8_M64 --skip-reference --out gpurun_out/v6_mid_timing.json 2>&1 | grep -E "graph|flush_median" | cut -c1-160
echo "== bench.py N=1"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; tail -c 6000 gpurun_out/bench_v6.json; tail -5 gpurun_out/bench_v6.err
echo "== bench.py --impl reference"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 | cut -c1-900
echo "== full suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short > gpurun_out/pytest_v6.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v6.log | cut -c1-330 | head -30
