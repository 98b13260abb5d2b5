#!/bin/bash
mkdir -p gpurun_out
BENCH_SKIP_EXTRAS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_w4 -s 140 -c 2 -o gpurun_out/decode_full -f \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
