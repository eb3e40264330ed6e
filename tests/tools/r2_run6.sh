#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_run6_pytest.log
tail -4 gpurun_out/r2_run6_pytest.log | cut -c1-250
timeout 300 python tests/tools/kbench.py C2 C3 2>&1 | grep -E "ndiff [1-9]|isect=|Error"
RT_B200_STAGE_OUT=1 timeout 300 python tests/tools/kbench.py C3 2>&1 | grep -E "ndiff [1-9]|isect=|Error"
timeout 300 python tests/tools/rank_share.py C3 1 8
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2_run6_bench1.json 2> gpurun_out/r2_run6_bench1.err; echo "bench1 exit $?"
tail -3 gpurun_out/r2_run6_bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_run6_bench1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'pageable',d['e2e']['pageable_destination_ms'],'redo',d['e2e']['redo_pixels_max'], 'parity', d['parity']['n_diff'], d['frame_sha256'][:16])
PY
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
for mode in p2p; do
timeout 900 python bench.py --gpus 2 --steps 10 --warmup 3 --mode $mode > gpurun_out/r2_run6_bench2_$mode.json 2> gpurun_out/r2_run6_bench2_$mode.err; echo "bench2 $mode exit $?"
tail -3 gpurun_out/r2_run6_bench2_$mode.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_run6_bench2_$mode.json'))
print('$mode value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'redo',d['e2e']['redo_pixels_max'],'sha',d['frame_sha256'][:16],'share',d['roofline']['kernel_share_of_step'],'kernel',d['roofline']['kernel_ms_avg'])
PY
done
fi
