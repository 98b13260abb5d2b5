#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== new tests first"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=600 --tb=short -k "batched_decode or chained_decode or column_shards or nvfp or mx_activation or mx_dynamic or fixtures" > gpurun_out/pytest_v4_new.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v4_new.log | cut -c1-330 | head -30
run() { tag=$1; shift; extra=""; if [ "${1:-}" = "--link" ]; then extra="--link $2"; shift 2; fi; env "$@" timeout 120 python tools/decode_chain.py --tag $tag $extra 2>&1 | tail -1 | cut -c1-200; }
echo "== decode chain"
run v52              GLB_GEMV_VARIANT=52
run v52_l1           --link 1 GLB_GEMV_VARIANT=52
run v52_l2           --link 2 GLB_GEMV_VARIANT=52
run v52_l3           --link 3 GLB_GEMV_VARIANT=52
run v7_t2            GLB_GEMV_TRIGGER=2
run v7_t2_l1         --link 1 GLB_GEMV_TRIGGER=2
run v7_t2_l2         --link 2 GLB_GEMV_TRIGGER=2
run v7_t2_l3         --link 3 GLB_GEMV_TRIGGER=2
run v7_t0            GLB_GEMV_TRIGGER=0
run v7_t0_l2         --link 2 GLB_GEMV_TRIGGER=0
run v7_t2_l2_d1      --link 2 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=1
run v7_t2_l2_d2      --link 2 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=2
run v7_t2_d3         GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=3
echo "== timelines"
TL_LINK=2 timeout 120 python tools/decode_timeline.py > gpurun_out/timeline_v7_l2.txt 2>&1; tail -12 gpurun_out/timeline_v7_l2.txt
timeout 120 python tools/decode_timeline.py > gpurun_out/timeline_v7_l0.txt 2>&1; tail -8 gpurun_out/timeline_v7_l0.txt
echo "== mid-M timing (ours)"
timeout 300 python tools/bench_vs_reference.py --only c2_M --skip-reference --out gpurun_out/v4_mid_timing.json 2>&1 | grep -E "graph|flush_median" | cut -c1-160
echo "== full suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short > gpurun_out/pytest_v4.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v4.log | cut -c1-330 | head -30
