"""BASELINE config 4 exactly as bench.py's `configs` line builds it (seed, cap 60), a few solves -- for compute-sanitizer."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np, torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
from racing_lmpc_ros2_b200.binding import make_reg_spec
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reg = (sys.argv[3] != "noreg") if len(sys.argv) > 3 else True
laps = P.workload.load_laps(); tr = P.workload.load_track("barc_center"); veh = P.configs.BARC_VEHICLE
cfg = dict(P.configs.barc_lmpc_config(20), num_ss_pts_per_lap=2, max_lap_stored=50, max_iter=60)
mpc = BatchedRacingMPC(veh, cfg, max_batch=Bn)
for l in P.workload.synthesise_laps(laps, 50): mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
if reg: mpc.set_error_dynamics(make_reg_spec([3, 4, 5], [[3, 4, 5]] * 3, [[0], [1], [1]], 0.6))
bb = P.workload.make_batch(veh, cfg, Bn, 0xB200 + 40, tr, laps)
dev = {k: torch.from_numpy(v).cuda() for k, v in bb.items()}
out = mpc.alloc_device_outputs(Bn)
for _ in range(reps): mpc.solve(dev, out)
torch.cuda.synchronize()
print("status", np.bincount(out["status"].cpu().numpy(), minlength=7), "iters max", int(out["iters"].max()))
