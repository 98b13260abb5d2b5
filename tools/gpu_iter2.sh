#!/bin/bash
# GPU visit 2: new CTA-pair prefill kernel (parity + sweep + ncu), issue-rate micro-benchmarks, decode variant sweep.
set -u
mkdir -p gpurun_out
echo "== pytest prefill pair kernel"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -x -k "prefill_pair" > gpurun_out/pytest_pair.log 2>&1; tail -15 gpurun_out/pytest_pair.log | cut -c1-400
echo "== prefill sweep (pair kernel)"
timeout 300 python tools/run_prefill.py 2>&1 | tail -6 | tee gpurun_out/prefill_sweep_v2.log
echo "== prefill sweep bf16 + single-CTA kernel for comparison"
GLB_PREFILL_VARIANT=1 timeout 300 python tools/run_prefill.py --quick 2>&1 | tail -2 | tee gpurun_out/prefill_sweep_v1.log
echo "== microbench"
timeout 300 ./tools/microbench 2>&1 | tee gpurun_out/microbench.log | tail -50
echo "== decode sweep"
: > gpurun_out/decode_sweep.log
run() { # variant pdl
  r=$(GLB_GEMV_VARIANT=$1 GLB_PDL=$2 BENCH_SKIP_EXTRAS=1 timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['roofline']['us_per_launch'],2), round(d['value']), round(d['e2e']['value']))
except Exception as e: print('ERR', e)")
  echo "variant=$1 pdl=$2 us/launch,GBs,e2e= $r" | tee -a gpurun_out/decode_sweep.log
}
run 52 1
run 6 1
run 62 1
run 61 1
run 5 1
echo "== pytest decode with variant 6"
GLB_GEMV_VARIANT=6 timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -k "decode or fixtures" > gpurun_out/pytest_v6.log 2>&1; tail -5 gpurun_out/pytest_v6.log | cut -c1-300
echo "== ncu full (pair prefill kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_a16w4_tc2 -c 1 -o gpurun_out/prefill2_full -f \
    python tools/run_prefill.py --quick > gpurun_out/ncu_prefill2.log 2>&1
tail -3 gpurun_out/ncu_prefill2.log
echo "== full gpu pytest"
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --tb=short > gpurun_out/pytest_gpu_full.log 2>&1; tail -5 gpurun_out/pytest_gpu_full.log | cut -c1-300
