#!/bin/bash
# round 2, GPU call A: test suite, bench line, row timings, ncu launch list (single GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_tests.txt
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?" >> gpurun_out/r2a_bench.err
timeout 600 python scripts/time_rows.py > gpurun_out/r2a_rows.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 3 --warmup 3 --no-configs > gpurun_out/r2a_ncu_bench.log 2>&1
tail -3 gpurun_out/r2a_tests.txt; tail -c 600 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
