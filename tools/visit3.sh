#!/bin/bash
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; extra=""; if [ "${1:-}" = "--link" ]; then extra="--link $2"; shift 2; fi; env "$@" timeout 120 python tools/decode_chain.py --tag $tag $extra 2>&1 | tail -1 | cut -c1-230; }
echo "== decode chain"
run v52              GLB_GEMV_VARIANT=52
run v7_nf_t2         GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=2
run v7_nf_t2_pre     GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=2 GLB_GEMV_PREWAIT=1
run v7_nf_t0         GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=0
run v7_f_t2          GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2
run v7_f_t0          GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=0
run v7_nf_t2_l1      --link 1 GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=2
run v7_nf_t2_l2      --link 2 GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=2
run v7_nf_t2_l3      --link 3 GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=2
run v7_nf_t2_l2_pre  --link 2 GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=2 GLB_GEMV_PREWAIT=1
run v7_f_t2_l2       --link 2 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2
run v7_f_t2_l2_pre   --link 2 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2 GLB_GEMV_PREWAIT=1
run v7_f_t0_l2       --link 2 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=0
run v7_nf_t0_l2      --link 2 GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=0
run v7_f_t2_l2_d1    --link 2 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=1
run v7_f_t2_l2_d2    --link 2 GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=2
run v7_f_t2_d3       GLB_GEMV_FLAGS=1 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=3
run v7_nf_t2_d3      GLB_GEMV_FLAGS=0 GLB_GEMV_TRIGGER=2 GLB_GEMV_DEBUG=3
echo "== timelines"
GLB_GEMV_FLAGS=0 timeout 120 python tools/decode_timeline.py > gpurun_out/timeline_v7_nf.txt 2>&1; tail -14 gpurun_out/timeline_v7_nf.txt
GLB_GEMV_FLAGS=0 TL_LINK=2 timeout 120 python tools/decode_timeline.py > gpurun_out/timeline_v7_nf_l2.txt 2>&1; tail -14 gpurun_out/timeline_v7_nf_l2.txt
GLB_GEMV_FLAGS=1 TL_LINK=2 timeout 120 python tools/decode_timeline.py > gpurun_out/timeline_v7_f_l2.txt 2>&1; tail -14 gpurun_out/timeline_v7_f_l2.txt
echo "== ncu full of the v7 kernel (isolated launches under replay)"
GLB_GEMV_FLAGS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_w4_m1 -s 150 -c 2 -o gpurun_out/decode_v7_full -f python tools/decode_chain.py --layers 64 --iters 2 > gpurun_out/ncu_v7.log 2>&1; tail -3 gpurun_out/ncu_v7.log
ls -la gpurun_out/*.ncu-rep | tail -3
