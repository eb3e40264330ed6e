#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
{
for so in 1 0; do echo "== STAGE_OUT=$so"; RT_B200_STAGE_OUT=$so timeout 300 python tests/tools/kbench.py C2 C3 C4 | grep -E "ndiff [1-9]|isect="; done
timeout 300 python tests/tools/rank_share.py C3 1 8
timeout 300 python tests/tools/rank_share.py C4 1 8
} > gpurun_out/r2_ab_completer_stage.log 2>&1
cat gpurun_out/r2_ab_completer_stage.log
for n in 1 2; do
  timeout 900 python bench.py --gpus $n --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_run12_bench$n.json 2> gpurun_out/r2_run12_bench$n.err; echo "bench$n exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_run12_bench$n.json'))
print('N=$n value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'pageable',d['e2e']['pageable_destination_ms'],'redo',d['e2e']['redo_pixels_max'],'sha',d['frame_sha256'][:16],'share',round(d['roofline']['kernel_share_of_step'],4), d['e2e']['rank0_phases_ms_median'])
PY
done
