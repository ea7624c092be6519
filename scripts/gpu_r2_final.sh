#!/bin/bash
# round 2 closing call (one GPU): row timings, bench line, launch list, compute-sanitizer on the kernels changed last
mkdir -p gpurun_out
timeout 600 python scripts/time_rows.py > gpurun_out/r2e_rows.txt 2>&1; cat gpurun_out/r2e_rows.txt | cut -c1-230
timeout 600 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 200 gpurun_out/r2e_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 3 --warmup 3 --no-configs > gpurun_out/r2e_ncu_bench.log 2>&1
{
echo "compute-sanitizer, second session of round 2 (B200): the kernels changed after sanitizer_r2.txt's first part"; echo
echo '$ compute-sanitizer --tool memcheck python -m pytest tests/test_regress.py -m gpu -q   (regression scan: cp.async double-buffered tiles, paired regressions; staged linearisation kernel inside the tick test)'
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_regress.py -m gpu -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
echo; echo '$ compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_regress.py -m gpu -q -k "regression_matches or tick_with"'
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_regress.py -m gpu -q -k "regression_matches or tick_with" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -8
echo; echo '$ compute-sanitizer --tool memcheck python scripts/prof_qp.py 64 1   (linearise with the staged stores, safe-set query, QP)'
timeout 600 compute-sanitizer --tool memcheck python scripts/prof_qp.py 64 1 2>&1 | grep -E "iters hist|ERROR SUMMARY|Invalid|Error" | head -8
} > gpurun_out/sanitizer_r2b.txt 2>&1
cat gpurun_out/sanitizer_r2b.txt
