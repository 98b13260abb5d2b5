#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/decode_sweep2.log
run() { # issue variant
  r=$(GLB_GEMV_ISSUE=$1 GLB_GEMV_VARIANT=$2 BENCH_SKIP_EXTRAS=1 timeout 300 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['roofline']['us_per_launch'],3), round(d['value']), round(d['e2e']['value']), round(d['e2e']['eager_value']), d['clocks'])
except Exception as e: print('ERR', e)")
  echo "issue=$1 variant=$2 us/launch,GBs,e2e,eager,clocks= $r" | tee -a gpurun_out/decode_sweep2.log
}
run 0 52
run 1 52
run 2 52
run 3 52
run 3 62
echo "== pytest decode (default issue mode)"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -k "decode or fixtures or a16w4_decode" > gpurun_out/pytest_dec.log 2>&1; tail -3 gpurun_out/pytest_dec.log | cut -c1-300
echo "== timeline issue=3"
GLB_LIB_PATH=$PWD/gemlite_b200/libgemlite_b200_timeline.so GLB_GEMV_VARIANT=52 timeout 200 python tools/decode_timeline.py > gpurun_out/timeline_v52_issue3.txt 2>&1; tail -9 gpurun_out/timeline_v52_issue3.txt
