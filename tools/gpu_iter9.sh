#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest all gpu"
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --tb=short > gpurun_out/pytest_gpu_full.log 2>&1; tail -12 gpurun_out/pytest_gpu_full.log | cut -c1-300
echo "== configs"
timeout 600 python tools/bench_configs.py --only=C3,Cx,C5 2>&1 | tee gpurun_out/configs2.txt | cut -c1-250
