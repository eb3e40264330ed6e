#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_run11_bench1.json 2> gpurun_out/r2_run11_bench1.err; echo "bench1 exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_run11_bench1.json'))
print('N=1 value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'pageable',d['e2e']['pageable_destination_ms'],'redo',d['e2e']['redo_pixels_max'],'sha',d['frame_sha256'][:16],'share',round(d['roofline']['kernel_share_of_step'],4), d['e2e']['rank0_phases_ms_median'])
print(d['parity']); print(json.dumps(d['extra']))
PY
