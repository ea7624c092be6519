"""Timings of the error-dynamics regression only (the f3 and configs[3] parts of time_rows.py); LMPC_REG_SHARED=0 selects the
one-scan-per-regression kernel for comparison."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np, torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
from racing_lmpc_ros2_b200.binding import make_reg_spec

laps = P.workload.load_laps()
tr = P.workload.load_track("barc_center")
veh = P.configs.BARC_VEHICLE
cfg = P.configs.barc_lmpc_config(20)
spec = make_reg_spec([3, 4, 5], [[3, 4, 5]] * 3, [[0], [1], [1]], 0.6)


def wall(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def dev_solve_time(mpc, batch, reps=5):
    dev = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
    out = mpc.alloc_device_outputs(batch["x_ic"].shape[0])
    t = wall(lambda: mpc.solve(dev, out), reps)
    st = out["status"].cpu().numpy()
    return t, float((st == 0).mean()), out


print("LMPC_REG_SHARED =", os.environ.get("LMPC_REG_SHARED", "(default: on)"))
mpc = BatchedRacingMPC(veh, cfg, max_batch=1024)
for l in laps: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
n = 1024 * 19
rng = np.random.default_rng(1)
idx = rng.integers(0, laps[-1]["x"].shape[0] - 1, n)
xq = laps[-1]["x"][idx]; uq = laps[-1]["u"][idx]
A0 = np.tile(np.eye(6), (n, 1, 1)); B0 = np.zeros((n, 6, 2)); C0 = np.zeros((n, 6))
t = wall(lambda: mpc.regress(spec, xq, uq, A0, B0, C0), 3)
r = mpc.regress(spec, xq, uq, A0, B0, C0)
print(f"regression, {n} queries x 3 outputs (host buffers): {t*1e3:.2f} ms; checksum A {np.abs(r[0]).sum():.12e} B {np.abs(r[1]).sum():.12e}", flush=True)
b1 = P.workload.make_batch(veh, cfg, 1024, 0xB200 + 2, tr, laps)
t0, ok0, _ = dev_solve_time(mpc, b1)
mpc.set_error_dynamics(spec)
t1, ok1, o1 = dev_solve_time(mpc, b1)
print(f"tick, 1024 instances: {t0*1e3:.3f} ms; with the regression: {t1*1e3:.3f} ms (solved {ok1:.4f}); cost checksum {float(o1['cost'].double().abs().sum()):.10e}", flush=True)
mpc.close()

cfg4 = dict(cfg, num_ss_pts_per_lap=2, max_lap_stored=50)
mpc = BatchedRacingMPC(veh, cfg4, max_batch=2048)
many = P.workload.synthesise_laps(laps, 50)
for l in many: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
b4 = P.workload.make_batch(veh, cfg4, 2048, 0xB200 + 3, tr, laps)
t0, ok0, _ = dev_solve_time(mpc, b4, 3)
mpc.set_error_dynamics(spec)
t1, ok1, _ = dev_solve_time(mpc, b4, 3)
print(f"50-lap safe set, 2048 instances: {t0*1e3:.2f} ms (solved {ok0:.4f}); with the regression over all 50 laps: {t1*1e3:.2f} ms (solved {ok1:.4f})", flush=True)
mpc.close()
