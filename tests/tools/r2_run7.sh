#!/bin/bash
mkdir -p gpurun_out
export RT_B200_LIB=exp
{
for alt in 0 2; do
  echo "== RT_B200_TB_ALT=$alt (0 = the product's binary traversal, 2 = 4-ary collapse)"
  RT_B200_TB_ALT=$alt timeout 600 python tests/tools/trace_bench.py C3 240000000
done
} > gpurun_out/r2_trace_bvh4.log 2>&1
cat gpurun_out/r2_trace_bvh4.log
unset RT_B200_LIB
timeout 900 python -m pytest tests -m gpu -x -q -k "controller or variants" 2>&1 | tail -4
timeout 300 python tests/tools/kbench.py C2 2>&1 | grep -E "ndiff [1-9]|isect=|Error"
