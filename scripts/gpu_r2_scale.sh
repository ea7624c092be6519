#!/bin/bash
# round 2: N-GPU A/B of the trajectory exchange (run under gpurun --gpus N).  usage: gpu_r2_scale.sh N [tag]
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
run() { # name, extra args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 "$@" \
    > gpurun_out/${TAG}_scale${N}_${name}.json 2> gpurun_out/${TAG}_scale${N}_${name}.err
  echo "$name exit $?"; tail -c 300 gpurun_out/${TAG}_scale${N}_${name}.err
}
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/${TAG}_multigpu_test.txt 2>&1; tail -3 gpurun_out/${TAG}_multigpu_test.txt; fi
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs > gpurun_out/${TAG}_scale${N}_single.json 2> gpurun_out/${TAG}_scale${N}_single.err
run peer_pool_slack
run peer_pool_barrier --sync barrier --no-configs
run peer_fixed --inputs fixed --no-configs
run nccl_pool --gather nccl --no-configs
run peer_sameseed --same-seed --no-configs
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_scale${N}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f.split("/")[-1], "value %.4g e2e %.4g ms/step %.4f qp_ms/rank %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], ["%.3f"%x for x in r["kernel_ms_per_rank"]]))
    except Exception as e:
        print(f, "ERR", e)
PY
