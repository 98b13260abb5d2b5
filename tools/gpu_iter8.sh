#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest prefill/mx/fixtures"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -x -k "prefill_pair or mxfp or fixtures" > gpurun_out/pytest_pair.log 2>&1; tail -8 gpurun_out/pytest_pair.log | cut -c1-300
echo "== sweep"
timeout 300 python tools/run_prefill.py 2>&1 | tail -6 | tee gpurun_out/prefill_sweep_v4.log
timeout 600 python tools/bench_configs.py --only=C5 2>&1 | tee gpurun_out/configs_c5.txt | cut -c1-250
