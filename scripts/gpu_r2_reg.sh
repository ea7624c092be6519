#!/bin/bash
# round 2: regression kernel variants (exact-size tile scan; 2 or 3 resident blocks per SM)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_regress.py -m gpu -x -q > gpurun_out/r2reg_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2reg_tests.txt
LMPC_REG_MINB=2 timeout 300 python scripts/time_regress.py > gpurun_out/r2reg_b2.txt 2>&1
LMPC_REG_MINB=3 timeout 300 python scripts/time_regress.py > gpurun_out/r2reg_b3.txt 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lmpc_regress -s 1 -c 1 -o gpurun_out/reg_r2_exact -f python scripts/prof_reg.py > gpurun_out/reg_r2_exact.log 2>&1
tail -3 gpurun_out/r2reg_tests.txt; cat gpurun_out/r2reg_b2.txt gpurun_out/r2reg_b3.txt
