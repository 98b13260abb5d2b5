#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest act quant / int8 / fp8 / fixtures"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -x -k "quant or int8 or fp8 or fixtures" > gpurun_out/pytest_a8.log 2>&1; tail -6 gpurun_out/pytest_a8.log | cut -c1-300
echo "== C4"
timeout 600 python tools/bench_configs.py --only=C4 2>&1 | tee gpurun_out/configs_c4.txt | grep -E "M=    1 |M=   64|M=  (256)|M= (1024|4096)|FAILED" | cut -c1-250
