#!/bin/bash
# round 2 profiling call (one GPU): ncu --set full of the QP kernel for BASELINE configs 2, 3, 5 + the launch list of bench.py
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:lmpc_qp_kernel -s 2 -c 1"
timeout 600 $NCU -o gpurun_out/qp_r2_c2 -f python scripts/prof_qp.py 1024 3 barc_lmpc > gpurun_out/qp_r2_c2.log 2>&1
timeout 600 $NCU -o gpurun_out/qp_r2_c3 -f python scripts/prof_qp.py 4096 3 iac_tracking > gpurun_out/qp_r2_c3.log 2>&1
timeout 600 $NCU -o gpurun_out/qp_r2_c5 -f python scripts/prof_qp.py 8192 3 barc_lmpc > gpurun_out/qp_r2_c5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-configs > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/qp_r2_c2.log gpurun_out/qp_r2_c3.log gpurun_out/qp_r2_c5.log
