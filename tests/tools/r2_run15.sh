#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "device_built_tree" 2>&1 | tail -30
CUDA_LAUNCH_BLOCKING=1 timeout 200 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_blocking_bench.json 2> gpurun_out/r2_blocking_bench.err; echo "blocking bench exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "ncu bench exit $?"; tail -3 gpurun_out/r2_launches.csv | cut -c1-200
