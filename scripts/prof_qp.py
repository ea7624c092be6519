"""Small driver for ncu: a few solves of BASELINE config 2 (BARC LMPC, N=20, K=96) at a given batch."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np, torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
veh = P.configs.BARC_VEHICLE; cfg = P.configs.barc_lmpc_config(20)
if "LMPC_TOL" in os.environ: cfg["tol"] = float(os.environ["LMPC_TOL"])
if "LMPC_MAXIT" in os.environ: cfg["max_iter"] = int(os.environ["LMPC_MAXIT"])
laps = P.workload.load_laps(); tr = P.workload.load_track("barc_center")
mpc = BatchedRacingMPC(veh, cfg, max_batch=Bn)
for l in laps: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
bb = P.workload.make_batch(veh, cfg, Bn, 0xB200 + 2, tr, laps)
dev = {k: torch.from_numpy(v).cuda() for k, v in bb.items()}
out = mpc.alloc_device_outputs(Bn)
for _ in range(reps): mpc.solve(dev, out)
torch.cuda.synchronize()
it = out["iters"].cpu().numpy(); st = out["status"].cpu().numpy()
print("iters hist", np.bincount(it), "status", np.bincount(st, minlength=5))
