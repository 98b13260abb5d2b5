#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== a8 mid tests + int8/fp8 tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=600 --tb=short -k "a8w8 or int8 or fp8 or chained" > gpurun_out/pytest_v5_new.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v5_new.log | cut -c1-330 | head -30
echo "== mid-M timing (ours)"
timeout 300 python tools/bench_vs_reference.py --only c2_M,c4_int8_M64 --skip-reference --out gpurun_out/v5_mid_timing.json 2>&1 | grep -E "graph|flush_median" | cut -c1-160
echo "== ncu mid kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_a16w4_mid -s 2 -c 1 -o gpurun_out/mid_w4_full -f python tools/run_mid.py 64 w4 > gpurun_out/ncu_mid_w4.log 2>&1; tail -2 gpurun_out/ncu_mid_w4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_a8w8_mid -s 2 -c 1 -o gpurun_out/mid_a8_full -f python tools/run_mid.py 64 a8 > gpurun_out/ncu_mid_a8.log 2>&1; tail -2 gpurun_out/ncu_mid_a8.log
echo "== vs reference + full suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short > gpurun_out/pytest_v5.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v5.log | cut -c1-330 | head -30
