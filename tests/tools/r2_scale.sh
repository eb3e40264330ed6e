#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for n in 8 4 2 1; do
  timeout 600 python bench.py --gpus $n --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2_scale_c3_${n}gpu.json 2> gpurun_out/r2_scale_c3_${n}gpu.err; echo "N=$n exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_scale_c3_${n}gpu.json'))
print('N=$n value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'redo',d['e2e']['redo_pixels_max'],'sha',d['frame_sha256'][:16],'share',round(d['roofline']['kernel_share_of_step'],4),'kernel',round(d['roofline']['kernel_ms_avg'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
timeout 600 python bench.py --gpus 8 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --mode nccl > gpurun_out/r2_scale_c3_8gpu_nccl.json 2> /dev/null
python - <<PY
import json
d=json.load(open('gpurun_out/r2_scale_c3_8gpu_nccl.json'))
print('N=8 nccl value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'sha',d['frame_sha256'][:16])
PY
timeout 600 python bench.py --gpus 8 --steps 20 --warmup 5 --no-extra --no-cpu-baseline --workload C4 > gpurun_out/r2_scale_c4_8gpu.json 2> /dev/null
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extra --no-cpu-baseline --workload C4 > gpurun_out/r2_scale_c4_1gpu.json 2> /dev/null
python - <<PY
import json
for n in (8,1):
    d=json.load(open('gpurun_out/r2_scale_c4_%dgpu.json'%n))
    print('C4 N=%d value'%n,round(d['value']),'ms',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3),'sha',d['frame_sha256'][:16])
PY
