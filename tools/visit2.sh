#!/bin/bash
# round 2, visit 2: full GPU suite with the switches gone, decode v7 experiments
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short > gpurun_out/pytest_v2.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_v2.log | cut -c1-330 | head -40
echo "== decode chain"
run() { tag=$1; shift; env "$@" timeout 120 python tools/decode_chain.py --tag $tag 2>&1 | tail -1 | cut -c1-250; }
run v52            GLB_GEMV_VARIANT=52
run v7_noflag_t2   GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=2
run v7_noflag_t0   GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=0
run v7_flag_t2     GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2
run v7_flag_t0     GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=0
run v7_flag_t2_d1  GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=1
run v7_flag_t2_d2  GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=2
run v7_flag_t2_d3  GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=3
run v7_flag_t0_d1  GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=0 GLB_GEMV_DEBUG=1
run v7_flag_t0_d3  GLB_GEMV_VARIANT=7 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=0 GLB_GEMV_DEBUG=3
echo "== A8 timing (ours only)"
timeout 300 python tools/bench_vs_reference.py --only c4_int8_M1,c4_int8_M64,c2_M --skip-reference --out gpurun_out/v2_a8_timing.json 2>&1 | grep -E "^\[ours\]|graph" | cut -c1-220
