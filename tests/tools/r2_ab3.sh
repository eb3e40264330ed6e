#!/bin/bash
# A/B timing of kernel builds at default settings, twice round-robin: usage  LIBS="ab_base librt_b200 ab_x" bash tests/tools/r2_ab3.sh
mkdir -p gpurun_out
{
for rep in 1 2; do
for lib in $LIBS; do
  echo "== $lib (pass $rep)"
  RT_B200_LIB=$PWD/ray-tracer-s8_b200/lib/$lib.so timeout 300 python tests/tools/kbench.py C2 C3 C4 | grep -E "ndiff [1-9]|isect="
done; done
} > gpurun_out/r2_ab3.log 2>&1
cat gpurun_out/r2_ab3.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
