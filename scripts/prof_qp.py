"""Small driver for ncu: a few solves of BASELINE config 2 (BARC LMPC, N=20, K=96) at a given batch."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np, torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
which = sys.argv[3] if len(sys.argv) > 3 else "barc_lmpc"      # barc_lmpc (configs 2 / 5) or iac_tracking (config 3)
if which == "iac_tracking":
    veh = P.configs.IAC_VEHICLE; cfg = P.configs.iac_tracking_config(40)
else:
    veh = P.configs.BARC_VEHICLE; cfg = P.configs.barc_lmpc_config(20)
if "LMPC_TOL" in os.environ: cfg["tol"] = float(os.environ["LMPC_TOL"])
if "LMPC_MAXIT" in os.environ: cfg["max_iter"] = int(os.environ["LMPC_MAXIT"])
laps = P.workload.load_laps(); tr = P.workload.load_track("putnam_optm" if which == "iac_tracking" else "barc_center")
mpc = BatchedRacingMPC(veh, cfg, max_batch=Bn)
if cfg["learning"]:
    for l in laps: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
# the seeds of bench.py's lines: config 2 -> 0xB200 + 2 (pool batch 0), configs 3 / 5 -> 0xB200 + 40 (extra_config_line, rank 0)
seed = 0xB200 + 2 if (which == "barc_lmpc" and Bn == 1024) else 0xB200 + 40
bb = P.workload.make_batch(veh, cfg, Bn, seed, tr, laps, mode="track" if which == "iac_tracking" else "barc")
dev = {k: torch.from_numpy(v).cuda() for k, v in bb.items()}
out = mpc.alloc_device_outputs(Bn)
for _ in range(reps): mpc.solve(dev, out)
torch.cuda.synchronize()
it = out["iters"].cpu().numpy(); st = out["status"].cpu().numpy()
print("iters hist", np.bincount(it), "status", np.bincount(st, minlength=5))
