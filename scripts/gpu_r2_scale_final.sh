#!/bin/bash
# end of round 2: the multi-GPU test and the default bench line on N GPUs (run under gpurun --gpus N).  usage: gpu_r2_scale_final.sh N
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/r2f_multigpu_test.txt 2>&1; tail -3 gpurun_out/r2f_multigpu_test.txt; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/r2f_scale${N}_single.json 2> gpurun_out/r2f_scale${N}_single.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/r2f_scale${N}_peer_pool_slack.json 2> gpurun_out/r2f_scale${N}_peer_pool_slack.err
echo "bench exit $?"; tail -c 300 gpurun_out/r2f_scale${N}_peer_pool_slack.err
python - <<PY
import json
for f in ("gpurun_out/r2f_scale${N}_single.json", "gpurun_out/r2f_scale${N}_peer_pool_slack.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "value %.4e" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.4e" % d["e2e"]["value"], "kernel min/max", d["roofline"].get("kernel_ms_min_max"))
        for k, v in (d.get("configs") or {}).items(): print("   ", k, "%.4e" % v["value"], v["ms_per_step"], v["solved_fraction"]) if "value" in v else print("   ", k, v)
    except Exception as e: print(f, "unreadable", e)
PY
