#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest all gpu"
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --tb=short > gpurun_out/pytest_gpu_full.log 2>&1; tail -4 gpurun_out/pytest_gpu_full.log | cut -c1-300
echo "== sweep"
timeout 300 python tools/run_prefill.py 2>&1 | tail -5 | tee gpurun_out/prefill_sweep_epi8.log
timeout 300 python tools/bench_configs.py --only=C5 2>&1 | grep "M= 4096" | cut -c1-200 | tee gpurun_out/c5_epi8.log
