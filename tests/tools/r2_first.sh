#!/bin/bash
# Round 2, first GPU call: the refactored library (bulk-copy staging, output stage, completion counters, tail tickets,
# async reference tree) against the oracle, then the A/B switches of the new kernel features.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_first_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_first_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_first_pytest.log
tail -5 gpurun_out/r2_first_pytest.log
{
for stage in 1 0; do for tail in 30 0; do
  echo "== STAGE_OUT=$stage TAIL_PERMILLE=$tail"
  RT_B200_STAGE_OUT=$stage RT_B200_TAIL_PERMILLE=$tail timeout 300 python tests/tools/kbench.py C2 C3
done; done
for tail in 0 30 60 120; do
  RT_B200_TAIL_PERMILLE=$tail timeout 300 python tests/tools/rank_share.py C3 1 8
done
RT_B200_STAGE_OUT=0 RT_B200_TAIL_PERMILLE=30 timeout 300 python tests/tools/rank_share.py C3 1 8
} > gpurun_out/r2_first_kbench.log 2>&1
tail -40 gpurun_out/r2_first_kbench.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_first_bench.json 2> gpurun_out/r2_first_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r2_first_bench.json
