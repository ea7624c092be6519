#!/bin/bash
# round 2, call B: GPU suite after the model / regression changes, regression timings and its ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_tests.txt
timeout 300 python scripts/time_regress.py > gpurun_out/r2b_reg.txt 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lmpc_regress -s 1 -c 1 -o gpurun_out/reg_r2_pair -f python scripts/prof_reg.py > gpurun_out/reg_r2_pair.log 2>&1
tail -4 gpurun_out/r2b_tests.txt; cat gpurun_out/r2b_reg.txt; tail -2 gpurun_out/reg_r2_pair.log
