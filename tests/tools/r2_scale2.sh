#!/bin/bash
mkdir -p gpurun_out
for n in 8 1; do
  timeout 300 python bench.py --gpus $n --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2_scale_c3_${n}gpu.json 2> gpurun_out/r2_scale_c3_${n}gpu.err; echo "N=$n exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_scale_c3_${n}gpu.json'))
print('N=$n value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'redo',d['e2e']['redo_pixels_max'],'sha',d['frame_sha256'][:16],'share',round(d['roofline']['kernel_share_of_step'],4),'kernel',round(d['roofline']['kernel_ms_avg'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
