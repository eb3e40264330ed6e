#!/bin/bash
# A/B timing of kernel builds: default settings and with the output stage / counters off
mkdir -p gpurun_out
{
for cfg in "1 1" "0 0"; do set -- $cfg
  echo "== STAGE_OUT=$1 COUNT_DONE=$2"
  RT_B200_STAGE_OUT=$1 RT_B200_COUNT_DONE=$2 timeout 300 python tests/tools/kbench.py C2 C3
done
for lib in $EXTRA_LIBS; do
  echo "== $lib STAGE_OUT=0 COUNT_DONE=0"
  RT_B200_LIB=$PWD/ray-tracer-s8_b200/lib/$lib.so RT_B200_STAGE_OUT=0 RT_B200_COUNT_DONE=0 timeout 300 python tests/tools/kbench.py C2 C3 | grep -v "^p[0-9]"
done
} > gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log
