#!/bin/bash
mkdir -p gpurun_out
{
echo "== round-1 tree (c7e65d3) on this box"
(cd _r1ref && timeout 300 python tests/tools/kbench.py C2 C3 | grep -v "^p[0-9]")
for cfg in "def" "world" ; do
  echo "== round-2 PID_ORDER=$cfg COUNT_DONE=0 STAGE=0"
  RT_B200_PID_ORDER=$cfg RT_B200_COUNT_DONE=0 RT_B200_STAGE_OUT=0 timeout 300 python tests/tools/kbench.py C2 C3 | grep -v "^p[0-9]"
done
echo "== round-2 default"
timeout 300 python tests/tools/kbench.py C2 C3 | grep -v "^p[0-9]"
for lib in ab_nounsure ab_ldg; do
  echo "== round-2 $lib (COUNT_DONE=0 STAGE=0)"
  RT_B200_LIB=$PWD/ray-tracer-s8_b200/lib/$lib.so RT_B200_COUNT_DONE=0 RT_B200_STAGE_OUT=0 timeout 300 python tests/tools/kbench.py C2 C3 | grep -v "^p[0-9]"
done
} > gpurun_out/r2_third_kbench.log 2>&1
cat gpurun_out/r2_third_kbench.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_third_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_third_pytest.log
tail -30 gpurun_out/r2_third_pytest.log | cut -c1-220
