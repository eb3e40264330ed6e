#!/bin/bash
mkdir -p gpurun_out
{
for lib in librt_b200 $EXTRA_LIBS; do
for so in 0 1; do
  echo "== $lib STAGE_OUT=$so"
  RT_B200_LIB=$PWD/ray-tracer-s8_b200/lib/$lib.so RT_B200_STAGE_OUT=$so timeout 300 python tests/tools/kbench.py C2 C3 | grep -E "ndiff [1-9]|isect=2"
done; done
} > gpurun_out/r2_ab2.log 2>&1
cat gpurun_out/r2_ab2.log
