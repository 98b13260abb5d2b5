#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest all gpu"
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --tb=short > gpurun_out/pytest_gpu_full.log 2>&1; tail -6 gpurun_out/pytest_gpu_full.log | cut -c1-300
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== C4"
timeout 600 python tools/bench_configs.py --only=C4 2>&1 | tee gpurun_out/configs_c4.txt | grep -E "M=  (256|512)|M= (1024|2048|4096)|FAILED" | cut -c1-250
