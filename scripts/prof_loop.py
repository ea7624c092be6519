"""Small driver for ncu: a closed loop of 8192 agents, a few ticks (launch list of the tick's kernels)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np, torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 6
agents = len(sys.argv) > 3 and sys.argv[3] == "agents"
laps = P.workload.load_laps(); tr = P.workload.load_track("barc_center"); veh = P.configs.BARC_VEHICLE
tb = np.ascontiguousarray(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tracks.npz"))["barc_center_table"], dtype=np.float64)
cfg = P.configs.barc_lmpc_config(20)
mpc = BatchedRacingMPC(veh, cfg, max_batch=nb)
for l in laps: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
mpc.set_track(tb)
bb = P.workload.make_batch(veh, cfg, nb, 0xB200 + 5, tr, laps)
opt = mpc.loop_options(0.025)
if agents: mpc.agents_create(nb, 1024)
f = mpc.closed_loop_agents if agents else mpc.closed_loop
o = f(opt, ticks, bb["x_ic"].copy(), bb["u_ic"].copy(), bb["X_ref"].copy(), bb["U_ref"].copy(), log=False)
torch.cuda.synchronize()
print("failed agents", float((o["fail_count"] > 0).mean()))
