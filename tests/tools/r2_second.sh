#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_second_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_second_pytest.log
tail -5 gpurun_out/r2_second_pytest.log
{
for cfg in "1 30 1" "1 0 1" "0 30 1" "1 30 0" "0 30 0" "0 0 0"; do set -- $cfg
  echo "== STAGE_OUT=$1 TAIL_PERMILLE=$2 COUNT_DONE=$3"
  RT_B200_STAGE_OUT=$1 RT_B200_TAIL_PERMILLE=$2 RT_B200_COUNT_DONE=$3 timeout 300 python tests/tools/kbench.py C2 C3 | grep -v "^p[0-9]"
done
for tail in 0 30 120; do
  RT_B200_TAIL_PERMILLE=$tail timeout 300 python tests/tools/rank_share.py C3 1 8
done
RT_B200_STAGE_OUT=0 RT_B200_COUNT_DONE=0 RT_B200_TAIL_PERMILLE=30 timeout 300 python tests/tools/rank_share.py C3 1 8
} > gpurun_out/r2_second_kbench.log 2>&1
cat gpurun_out/r2_second_kbench.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_second_bench.json 2> gpurun_out/r2_second_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_second_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['kernel_ms_avg'])
PY
