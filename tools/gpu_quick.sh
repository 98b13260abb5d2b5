#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --tb=short -x > gpurun_out/pytest_iter.log 2>&1; tail -3 gpurun_out/pytest_iter.log
: > gpurun_out/quick.log
run() { # variant pdl
  r=$(GLB_GEMV_VARIANT=$1 GLB_PDL=$2 BENCH_SKIP_EXTRAS=1 timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['roofline']['us_per_launch'],2), round(d['value']), round(d['e2e']['value']))
except Exception as e: print('ERR', e)")
  echo "decode variant=$1 pdl=$2 us/launch,GBs,e2e= $r" | tee -a gpurun_out/quick.log
}
run 52 1
run 42 1
run 3 1
for shape in "4096 8192 8192" "4096 4096 4096" "1024 8192 8192" "128 8192 8192" "32 4096 4096"; do
  set -- $shape
  PF_M=$1 PF_K=$2 PF_N=$3 timeout 300 python tools/run_prefill.py 2>&1 | tail -2 | tr '\n' ' ' | tee -a gpurun_out/quick.log; echo | tee -a gpurun_out/quick.log
done
