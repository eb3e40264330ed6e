#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:render_kernel_lanes --launch-skip 2 -c 1 -f"
$NCU -o gpurun_out/r2_ncu_c3 python tests/tools/prof_one.py C3 2 3 > /dev/null 2>&1
RT_B200_STAGE_OUT=1 $NCU -o gpurun_out/r2_ncu_c3_stage python tests/tools/prof_one.py C3 2 3 > /dev/null 2>&1
$NCU -o gpurun_out/r2_ncu_c2_k1 python tests/tools/prof_one.py C2 1 3 > /dev/null 2>&1
ls -la gpurun_out/r2_ncu_c*.ncu-rep
# launch list of the bench command (no-extra, short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
tail -3 gpurun_out/r2_launches.csv
