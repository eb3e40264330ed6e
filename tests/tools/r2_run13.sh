#!/bin/bash
# 1024-thread CTAs + stream-ordered waits by value: tests, kernel times, end to end with and without the value waits
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
{
timeout 300 python tests/tools/kbench.py C2 C3 C4 | grep -E "ndiff [1-9]|isect="
timeout 300 python tests/tools/rank_share.py C3 1 8
} > gpurun_out/r2_cta1024.log 2>&1
cat gpurun_out/r2_cta1024.log
for w in value kernel; do
for n in 1 2; do
  [ $w = kernel ] && export RT_B200_NO_STREAM_WAIT=1
  tag=${w}_$n
  if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py"; fi
  timeout 900 $cmd --gpus $n --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_run13_bench_$tag.json 2> gpurun_out/r2_run13_bench_$tag.err; echo "bench $tag exit $?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_run13_bench_$tag.json').read().strip().splitlines()[-1])
print('$tag value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['ms_per_step'],3),'pageable',d['e2e'].get('pageable_destination_ms'),'redo',d['e2e'].get('redo_pixels_max'),'sha',d['frame_sha256'][:16],'share',round(d['roofline']['kernel_share_of_step'],4), d['e2e'].get('rank0_phases_ms_median'))
PY
done
done
