import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); warnings.filterwarnings("ignore")
import numpy as np
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
from oracle import Oracle
lib = sys.argv[1] if len(sys.argv) > 1 else None
laps = P.workload.load_laps(); tr = P.workload.load_track("barc_center")
for name, cfg in (("tracking", P.configs.barc_tracking_config(20)), ("lmpc", P.configs.barc_lmpc_config(20))):
    veh = P.configs.BARC_VEHICLE
    mpc = BatchedRacingMPC(veh, cfg, max_batch=64, lib_path=lib); orc = Oracle(veh, cfg)
    if cfg["learning"]:
        for l in laps: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"]); orc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
    batch = P.workload.make_batch(veh, cfg, 64, 1, tr, laps, mode="barc")
    out = mpc.solve(batch); ref = orc.step_batch(batch, impl="port", nthreads=8)
    print(name, "gpu status", np.bincount(out["status"], minlength=5), "iters", out["iters"].mean(), "| port iters", ref["iters"].mean(), "maxdiff X", np.abs(out["X_optm"] - ref["X"]).max())
