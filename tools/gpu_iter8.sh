#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest prefill/mx/fixtures"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -x -k "prefill_pair or mxfp or fixtures" > gpurun_out/pytest_pair.log 2>&1; tail -4 gpurun_out/pytest_pair.log | cut -c1-300
for pf in 0 2 4 8; do
  echo "== sweep prefetch=$pf"
  GLB_PREFILL_PREFETCH=$pf timeout 300 python tools/run_prefill.py "--shapes=4096,8192,8192;4096,4096,4096;1024,8192,8192" 2>&1 | tail -3 | tee gpurun_out/prefill_sweep_pf$pf.log
done
