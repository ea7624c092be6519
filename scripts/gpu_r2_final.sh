#!/bin/bash
# round 2 closing call (one GPU): row timings, bench line, compute-sanitizer on the new device paths
mkdir -p gpurun_out
timeout 600 python scripts/time_rows.py > gpurun_out/r2e_rows.txt 2>&1; cat gpurun_out/r2e_rows.txt | cut -c1-230
timeout 600 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 200 gpurun_out/r2e_bench.err
{
echo "compute-sanitizer evidence (B200, round 2)"; echo
echo '$ compute-sanitizer --tool memcheck python scripts/prof_qp.py 64 1'
timeout 900 compute-sanitizer --tool memcheck python scripts/prof_qp.py 64 1 2>&1 | grep -E "iters hist|ERROR SUMMARY|Invalid|Error" | head -8
echo; echo '$ compute-sanitizer --tool memcheck python -m pytest tests/test_agents.py -m gpu -q   (per-agent safe sets: recorder, add_lap, query kernels; 560 + 300 closed-loop ticks)'
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_agents.py -m gpu -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
echo; echo '$ compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_agents.py -m gpu -q -k recorder'
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_agents.py -m gpu -q -k recorder 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -8
echo; echo '$ compute-sanitizer --tool memcheck python -m pytest tests/test_cpp_adapter.py tests/test_regress.py -m gpu -q   (control-map kernel, windowed regression scan)'
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_cpp_adapter.py tests/test_regress.py -m gpu -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -8
} > gpurun_out/sanitizer_r2.txt 2>&1
cat gpurun_out/sanitizer_r2.txt
