#!/bin/bash
# round 2 closing call (one GPU): GPU suite with its parity figures, bench line, launch list, row timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2f_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_tests.txt
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-configs > gpurun_out/r2f_ncu_bench.log 2>&1
timeout 600 python scripts/time_rows.py > gpurun_out/r2f_rows.txt 2>&1
tail -3 gpurun_out/r2f_tests.txt; cut -c1-200 gpurun_out/r2f_rows.txt; tail -c 300 gpurun_out/r2f_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['other_kernels_ms'], d['roofline']['fp64']['frac'], d['cpu_baseline']['value'])
for k,v in (d.get('configs') or {}).items(): print(k, v['value'], v['ms_per_step'], v['solved_fraction'], v['kernel_ms'])
"
