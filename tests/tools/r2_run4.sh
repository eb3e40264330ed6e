#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_run4_pytest.log
tail -4 gpurun_out/r2_run4_pytest.log | cut -c1-250
{
timeout 300 python tests/tools/kbench.py C2 C3 C4 | grep -E "ndiff [1-9]|isect="
for tail in 0 30 120; do
  RT_B200_TAIL_PERMILLE=$tail timeout 300 python tests/tools/rank_share.py C3 1 2 8
done
RT_B200_TAIL_PERMILLE=30 timeout 300 python tests/tools/rank_share.py C4 1 8
} > gpurun_out/r2_run4_kbench.log 2>&1
cat gpurun_out/r2_run4_kbench.log
RT_B200_TIMING=1 timeout 300 python tests/tools/c5_sweep.py > gpurun_out/r2_run4_c5.log 2>&1; tail -30 gpurun_out/r2_run4_c5.log
