#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:render_kernel_lanes --launch-skip 2 -c 1 -f"
RT_B200_STAGE_OUT=0 RT_B200_COUNT_DONE=0 $NCU -o gpurun_out/r2_ncu_nostage python tests/tools/prof_one.py C3 2 3 > /dev/null 2>&1
$NCU -o gpurun_out/r2_ncu_default python tests/tools/prof_one.py C3 2 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_ncu_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_ncu_pytest.log
tail -30 gpurun_out/r2_ncu_pytest.log | cut -c1-250
