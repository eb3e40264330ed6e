#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run8_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_run8_pytest.log
tail -4 gpurun_out/r2_run8_pytest.log | cut -c1-250
{
for lib in librt_b200 ab_postpone; do
  echo "== $lib"
  RT_B200_LIB=$PWD/ray-tracer-s8_b200/lib/$lib.so timeout 300 python tests/tools/kbench.py C2 C3 | grep -E "ndiff [1-9]|isect="
done
} > gpurun_out/r2_ab_postpone.log 2>&1
cat gpurun_out/r2_ab_postpone.log
timeout 300 python tests/tools/rank_share.py C3 1 8
for n in 1 2; do
  timeout 900 python bench.py --gpus $n --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_run8_bench$n.json 2> gpurun_out/r2_run8_bench$n.err; echo "bench$n exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_run8_bench$n.json'))
print('N=$n value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'redo',d['e2e']['redo_pixels_max'],'sha',d['frame_sha256'][:16],'share',round(d['roofline']['kernel_share_of_step'],4), d['e2e']['rank0_phases_ms_median'])
PY
done
