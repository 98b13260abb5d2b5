#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest int8/fp8/fixtures"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -x -k "int8 or fp8 or fixtures" > gpurun_out/pytest_a8.log 2>&1; tail -8 gpurun_out/pytest_a8.log | cut -c1-300
echo "== C4"
timeout 600 python tools/bench_configs.py --only=C4 2>&1 | tee gpurun_out/configs_c4.txt | grep -E "M=  (256|512)|M= (1024|2048|4096)|FAILED" | cut -c1-250
