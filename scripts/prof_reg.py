"""Small driver for ncu: a few ticks of BASELINE config 4 (50 stored laps, 2 nearest per lap) with the error-dynamics
regression on every stage, 2048 instances."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np, torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
from racing_lmpc_ros2_b200.binding import make_reg_spec
Bn = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
laps = P.workload.load_laps(); tr = P.workload.load_track("barc_center")
veh = P.configs.BARC_VEHICLE
cfg4 = dict(P.configs.barc_lmpc_config(20), num_ss_pts_per_lap=2, max_lap_stored=50)
mpc = BatchedRacingMPC(veh, cfg4, max_batch=Bn)
for l in P.workload.synthesise_laps(laps, 50): mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
mpc.set_error_dynamics(make_reg_spec([3, 4, 5], [[3, 4, 5]] * 3, [[0], [1], [1]], 0.6))
bb = P.workload.make_batch(veh, cfg4, Bn, 0xB200 + 3, tr, laps)
dev = {k: torch.from_numpy(v).cuda() for k, v in bb.items()}
out = mpc.alloc_device_outputs(Bn)
for _ in range(reps): mpc.solve(dev, out)
torch.cuda.synchronize()
print("status", np.bincount(out["status"].cpu().numpy(), minlength=6))
