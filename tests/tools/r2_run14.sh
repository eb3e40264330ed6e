#!/bin/bash
mkdir -p gpurun_out
LIBS="ab_base librt_b200" bash tests/tools/r2_ab3.sh
mv gpurun_out/r2_ab3.log gpurun_out/r2_ab_node80.log
# launch list of the bench command under ncu (value waits fall back to wait kernels there)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "ncu bench exit $?"; tail -3 gpurun_out/r2_launches.csv | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on -k regex:render_kernel_lanes --launch-skip 2 -c 1 -f"
timeout 300 $NCU -o gpurun_out/r2_ncu_c3 python tests/tools/prof_one.py C3 2 3 > /dev/null 2>&1
