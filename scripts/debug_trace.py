"""Builds a trace variant of the library (per-iteration printf from instance 0) and runs one small batch."""
import os, sys, subprocess, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
lib = os.path.join(ROOT, "build_dbg", "liblmpc_trace.so")
veh = P.configs.BARC_VEHICLE; cfg = P.configs.barc_lmpc_config(20)
laps = P.workload.load_laps(); tr = P.workload.load_track("barc_center")
mpc = BatchedRacingMPC(veh, cfg, max_batch=8, lib_path=lib)
for l in laps: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
batch = P.workload.make_batch(veh, cfg, 8, 1, tr, laps, mode="barc")
out = mpc.solve(batch)
print("status", out["status"], "iters", out["iters"])
