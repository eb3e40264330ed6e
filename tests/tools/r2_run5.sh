#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_run5_bench1.json 2> gpurun_out/r2_run5_bench1.err; echo "bench1 exit $?"
tail -3 gpurun_out/r2_run5_bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run5_bench1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e']['pageable_destination_ms'],'redo',d['e2e']['redo_pixels_max'])
print('parity',d.get('parity')); print('sha',d['frame_sha256'][:16]); print('roofline frac',d['roofline']['frac'],'share',d['roofline']['kernel_share_of_step'], 'flush', d['run'])
print('extra',json.dumps(d.get('extra'),indent=1))
print('cpu',d.get('cpu_baseline',{}).get('value'))
PY
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
for mode in p2p nccl; do
timeout 900 python bench.py --gpus 2 --steps 10 --warmup 3 --mode $mode > gpurun_out/r2_run5_bench2_$mode.json 2> gpurun_out/r2_run5_bench2_$mode.err; echo "bench2 $mode exit $?"
tail -3 gpurun_out/r2_run5_bench2_$mode.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_run5_bench2_$mode.json'))
print('$mode value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'sha',d['frame_sha256'][:16],'share',d['roofline']['kernel_share_of_step'],'kernel',d['roofline']['kernel_ms_avg'])
PY
done
timeout 600 python -m pytest tests -m gpu -x -q -k "two_devices or multi_context" 2>&1 | tail -5
fi
