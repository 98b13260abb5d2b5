#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --tb=short -k "decode or fixtures or a16w4" > gpurun_out/pytest_iter.log 2>&1; tail -3 gpurun_out/pytest_iter.log
: > gpurun_out/decode_sweep.log
run() { # variant pdl debug alias
  r=$(GLB_GEMV_VARIANT=$1 GLB_PDL=$2 GLB_GEMV_DEBUG=$3 BENCH_ALIAS=$4 BENCH_SKIP_EXTRAS=1 timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['roofline']['us_per_launch'],2), round(d['value']), round(d['e2e']['value']))
except Exception as e: print('ERR', e)")
  echo "variant=$1 pdl=$2 debug=$3 alias=$4 us/launch,GBs,e2e= $r" | tee -a gpurun_out/decode_sweep.log
}
run 52 1 0 0
run 52 1 1 0
run 52 1 3 0
run 42 1 0 0
run 52 0 0 0
