#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest mx / fixtures"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -x -k "mxfp or fixtures" > gpurun_out/pytest_mx.log 2>&1; tail -8 gpurun_out/pytest_mx.log | cut -c1-300
echo "== C5"
timeout 600 python tools/bench_configs.py --only=C5 2>&1 | tee gpurun_out/configs_c5.txt | cut -c1-250
