#!/bin/bash
# GPU visit 3: pair prefill kernel after the remote-arrive fix (+ ncu), decode phase timelines for variants 52 / 6 / 62.
set -u
mkdir -p gpurun_out
echo "== pytest prefill pair kernel"
timeout 600 python -m pytest tests -m gpu -q --timeout=240 --tb=short -x -k "prefill_pair" > gpurun_out/pytest_pair.log 2>&1; tail -3 gpurun_out/pytest_pair.log | cut -c1-400
echo "== prefill sweep (pair kernel)"
timeout 300 python tools/run_prefill.py 2>&1 | tail -6 | tee gpurun_out/prefill_sweep_v2.log
echo "== decode timelines"
for v in 52 6 62; do
  GLB_LIB_PATH=$PWD/gemlite_b200/libgemlite_b200_timeline.so GLB_GEMV_VARIANT=$v timeout 200 python tools/decode_timeline.py > gpurun_out/timeline_v$v.txt 2>&1
  echo "-- variant $v"; tail -13 gpurun_out/timeline_v$v.txt
done
echo "== ncu full (pair prefill kernel)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_a16w4_tc2 -c 1 -o gpurun_out/prefill2_full -f \
    python tools/run_prefill.py --quick > gpurun_out/ncu_prefill2.log 2>&1
tail -2 gpurun_out/ncu_prefill2.log
