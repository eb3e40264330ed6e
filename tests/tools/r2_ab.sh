#!/bin/bash
# A/B timing of kernel builds: usage  EXTRA_LIBS="ab_x ab_y" bash tests/tools/r2_ab.sh
mkdir -p gpurun_out
{
for lib in librt_b200 $EXTRA_LIBS; do
for cfg in "1 1" "1 0" "0 0"; do set -- $cfg
  echo "== $lib STAGE_OUT=$1 COUNT_DONE=$2"
  RT_B200_LIB=$PWD/ray-tracer-s8_b200/lib/$lib.so RT_B200_STAGE_OUT=$1 RT_B200_COUNT_DONE=$2 timeout 300 python tests/tools/kbench.py C2 C3 | grep -E "ndiff [1-9]|isect="
done; done
} > gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log
