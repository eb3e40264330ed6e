#!/bin/bash
# what the driver runs at round end, on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','unit','ms_per_step','n_gpus','steps','warmup','gpu_launches','frame_sha256')})
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'pageable',d['e2e'].get('pageable_destination_ms'))
print('roofline',{k:d['roofline'][k] for k in ('bound','achieved','peak','frac','traffic','kernel_share_of_step')})
print('cpu_baseline',d['cpu_baseline'])
print('parity',d['parity'])
print('clocks',d['clocks'])
print('extra',json.dumps(d.get('extra_workloads'))[:900])
r=json.loads(open('gpurun_out/r2_final_bench_reference.json').read().strip().splitlines()[-1])
print('reference',{k:r.get(k) for k in ('impl','value','unit','ms_per_step','cpu_baseline','e2e')})
PY
