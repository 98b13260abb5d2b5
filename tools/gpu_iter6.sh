#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py 2>&1 | tee gpurun_out/configs.txt | cut -c1-250
