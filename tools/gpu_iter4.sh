#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest prefill (pair + fixtures)"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -x -k "prefill_pair or fixtures or a16w4" > gpurun_out/pytest_pair.log 2>&1; tail -3 gpurun_out/pytest_pair.log | cut -c1-400
echo "== prefill sweep (pair kernel)"
timeout 300 python tools/run_prefill.py 2>&1 | tail -6 | tee gpurun_out/prefill_sweep_v3.log
echo "== ncu full (pair prefill kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_a16w4_tc2 -c 1 -o gpurun_out/prefill2_full -f \
    python tools/run_prefill.py --quick > gpurun_out/ncu_prefill2.log 2>&1
tail -2 gpurun_out/ncu_prefill2.log
