#!/bin/bash
# One GPU-box visit that validates everything written without GPU time at the end of round 1:
#   1. the default suite (must stay green),
#   2. GEMLITE_B200_EXPERIMENTAL=1: MX / NV activation quantisers + dynamic MX layers (gated tests),
#   3. GLB_A8_GEMV=1: A8W8 decode kernel (gated tests + the A8W8 tests re-routed through it) and its C4 timings,
#   3a. GLB_A8_SPLITK=1: split-K instantiation of the single-CTA A8W8 tcgen05 kernel (16 < M <= 128),
#   3b. GLB_MXV_VARIANT=1: MXFP4 decode kernel on mma.sync,
#   4. decode flavours (leader-lane polling, pre-wait-all): parity of the decode tests + chained / 4-stream timings.
# Flavour libraries must be built BEFORE the visit (nvcc cross-compiles here; the box only loads them):
#   python -m gemlite_b200.build --flavour leaderwait -DGLB_GEMV_LEADER_WAIT=1
#   python -m gemlite_b200.build --flavour prewait    -DGLB_GEMV_PREWAIT_ALL=1
#   python -m gemlite_b200.build --flavour both       -DGLB_GEMV_LEADER_WAIT=1 -DGLB_GEMV_PREWAIT_ALL=1
# Output: gpurun_out/optin_*.log (merged back by gpurun).
set -u
mkdir -p gpurun_out
sum() { grep -E "^(FAILED|ERROR)|passed|failed" "$1" | cut -c1-240 | tail -12; }
echo "== 1. default suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 --tb=short > gpurun_out/optin_default.log 2>&1; sum gpurun_out/optin_default.log
echo "== 2. MX / NV activation quantisers + dynamic MX layers"
GEMLITE_B200_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 --tb=short \
    -k "mx_activation or mx_dynamic or fixtures" > gpurun_out/optin_mx.log 2>&1; sum gpurun_out/optin_mx.log
echo "== 3. A8W8 decode kernel"
GLB_A8_GEMV=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 --tb=short \
    -k "a8w8 or int8 or fp8 or fixtures" > gpurun_out/optin_a8.log 2>&1; sum gpurun_out/optin_a8.log
timeout 300 python tools/bench_configs.py --only=C4 --small > gpurun_out/optin_c4_default.txt 2>&1
GLB_A8_GEMV=1 timeout 300 python tools/bench_configs.py --only=C4 --small > gpurun_out/optin_c4_gemv.txt 2>&1
grep -E "M= +(1|2|4|8|16) " gpurun_out/optin_c4_default.txt | cut -c1-150
grep -E "M= +(1|2|4|8|16) " gpurun_out/optin_c4_gemv.txt | cut -c1-150
echo "== 3a. A8W8 split-K for 16 < M <= 128 (GLB_A8_SPLITK=1)"
GLB_A8_SPLITK=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 --tb=short \
    -k "a8w8 or int8 or fp8 or fixtures" > gpurun_out/optin_a8_splitk.log 2>&1; sum gpurun_out/optin_a8_splitk.log
GLB_A8_SPLITK=1 timeout 300 python tools/bench_configs.py --only=C4 --mid > gpurun_out/optin_c4_splitk.txt 2>&1
timeout 300 python tools/bench_configs.py --only=C4 --mid > gpurun_out/optin_c4_nosplit.txt 2>&1
grep "M= " gpurun_out/optin_c4_nosplit.txt | cut -c1-150; grep "M= " gpurun_out/optin_c4_splitk.txt | cut -c1-150
echo "== 3b. MXFP4 decode kernel on mma.sync (GLB_MXV_VARIANT=1)"
GLB_MXV_VARIANT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 --tb=short -k "mxfp" > gpurun_out/optin_mxv.log 2>&1; sum gpurun_out/optin_mxv.log
timeout 300 python tools/bench_configs.py --only=C5 --quick 2>&1 | grep "M= " | cut -c1-150
GLB_MXV_VARIANT=1 timeout 300 python tools/bench_configs.py --only=C5 --quick 2>&1 | grep "M= " | cut -c1-150
echo "== 4. decode flavours"
for f in "" leaderwait prewait both; do
  lib="gemlite_b200/libgemlite_b200${f:+_$f}.so"
  [ -f "$lib" ] || { echo "($lib not built)"; continue; }
  tag=${f:-default}
  GLB_LIB_PATH=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 --tb=short \
      -k "decode or fixtures" > gpurun_out/optin_decode_$tag.log 2>&1; echo "$tag: $(sum gpurun_out/optin_decode_$tag.log | tail -1)"
  GLB_LIB_PATH=$PWD/$lib timeout 300 python tools/decode_streams.py --streams 1,4 --steps 40 2>&1 | sed "s/^/$tag /" | tee -a gpurun_out/optin_decode_streams.txt
done
ls -la gpurun_out | tail -15
