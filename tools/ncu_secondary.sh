#!/bin/bash
# ncu --set full captures of the secondary kernels (second launch of each in tools/probe_kernels.py, i.e. warmed code and
# descriptors).  ncu's -k regex matches the base kernel name (no template arguments), so launches are picked by -s.
# Summaries: python tools/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>_r1_ncu.txt
set -u
mkdir -p gpurun_out
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $4 -c 1 -o gpurun_out/$1 -f python tools/probe_kernels.py $3 > gpurun_out/ncu_$1.log 2>&1; tail -1 gpurun_out/ncu_$1.log | cut -c1-120; }
cap a8_int8_tc2 gemm_a8w8_tc2_kernel a8 1      # int8 layer, second call
cap a8_fp8_tc2 gemm_a8w8_tc2_kernel a8 3       # fp8 layer, second call
cap act_quant_vec act_quant_vec a8 1
cap gemv_mxfp4 gemv_mxfp4 mx 1
cap mxfp4_tc2 gemm_a16w4_tc2_kernel mx 1
cap a16w4_splitk gemm_a16w4_tc2_kernel splitk 1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
