#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== a16w8 + vs-reference a16w8"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=600 --tb=short -k "a16w8_weight_only or mx_dynamic" > gpurun_out/pytest_v11_new.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v11_new.log | cut -c1-330 | head -20
echo "== tensor peaks"
timeout 300 python tools/measure_peaks.py > gpurun_out/tensor_peaks_r2.json 2>&1; cat gpurun_out/tensor_peaks_r2.json | tr -d '\n' | cut -c1-600; echo
echo "== bench N=1"
BENCH_SKIP_REFERENCE_TRITON=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v11.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_per_step_rank0_min_median_max','clocks')}, d['roofline']['frac'])
PY
echo "== full suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short > gpurun_out/pytest_v11.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v11.log | cut -c1-330 | head -30
