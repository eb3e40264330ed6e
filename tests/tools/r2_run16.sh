#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tests/tools/kbench.py C2 C3 C4 | grep -E "ndiff [1-9]|isect=" > gpurun_out/r2_ffma2_kbench.log; cat gpurun_out/r2_ffma2_kbench.log
timeout 300 python tests/tools/rank_share.py C3 1 8 | tee -a gpurun_out/r2_ffma2_kbench.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench exit $?"
NCU="ncu --set full --clock-control none --import-source on -k regex:render_kernel_lanes --launch-skip 2 -c 1 -f"
timeout 300 $NCU -o gpurun_out/r2_ncu_c3 python tests/tools/prof_one.py C3 2 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "ncu bench exit $?"
