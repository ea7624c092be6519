"""Timings of the SURVEY 8f rows and of the parity-case configurations (not the bench line): closed loop, SQP to
convergence, error-dynamics regression, track functions, IAC N=40, 50-lap safe set.  Wall clock around the library
calls after a warm-up call; host-buffer calls include their H2D / D2H copies, device-buffer calls end with a synchronise."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np, torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
from racing_lmpc_ros2_b200.binding import make_reg_spec

laps = P.workload.load_laps()
tb = np.ascontiguousarray(np.load(os.path.join(ROOT, "tests", "golden", "tracks.npz"))["barc_center_table"], dtype=np.float64)
tr = P.workload.load_track("barc_center")
veh = P.configs.BARC_VEHICLE


def wall(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def dev_solve_time(mpc, batch, reps=5):
    Bn = batch["x_ic"].shape[0]
    dev = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
    out = mpc.alloc_device_outputs(Bn)
    t = wall(lambda: mpc.solve(dev, out), reps)
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    return t, float((st == 0).mean()), float(it.mean()), int(it.max())


# ---- f2: closed loop of 8192 agents (configs[4]: one GPU's share of the 65536-agent Monte-Carlo run)
cfg = P.configs.barc_lmpc_config(20)
nb, ticks = 8192, 40
mpc = BatchedRacingMPC(veh, cfg, max_batch=nb)
for l in laps: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
mpc.set_track(tb)
bb = P.workload.make_batch(veh, cfg, nb, 0xB200 + 5, tr, laps)
x, u_prev, X_last, U_last = bb["x_ic"].copy(), bb["u_ic"].copy(), bb["X_ref"].copy(), bb["U_ref"].copy()
opt = mpc.loop_options(0.025)
res = {}
def run_loop(): res["o"] = mpc.closed_loop(opt, ticks, x, u_prev, X_last, U_last, log=False)
t = wall(run_loop, 2)
print(f"closed loop (prepare + solve + plant on the device): {nb} agents x {ticks} ticks in {t*1e3:.1f} ms -> {nb*ticks/t:.3e} agent-ticks/s; "
      f"agents with a failed tick {float((res['o']['fail_count'] > 0).mean()):.4f}", flush=True)

# the same loop with per-agent safe sets and device-side lap recording (every agent learns from its own laps)
t_ag = []
for rep in range(3):      # the agents' recorders and safe sets persist across calls: a fresh set per repetition
    mpc.agents_create(nb, 1024)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res["a"] = mpc.closed_loop_agents(opt, ticks, x, u_prev, X_last, U_last, log=False)
    torch.cuda.synchronize(); t_ag.append(time.perf_counter() - t0)
t = min(t_ag[1:])
print(f"closed loop, per-agent safe sets + recording on the device: {nb} agents x {ticks} ticks in {t*1e3:.1f} ms -> {nb*ticks/t:.3e} agent-ticks/s; "
      f"agents with a failed tick {float((res['a']['fail_count'] > 0).mean()):.4f}", flush=True)
mpc.agents_destroy()

# ---- f3: error-dynamics regression, stand-alone and inside the tick
spec = make_reg_spec([3, 4, 5], [[3, 4, 5]] * 3, [[0], [1], [1]], 0.6)
n = 1024 * 19
rng = np.random.default_rng(1)
idx = rng.integers(0, laps[-1]["x"].shape[0] - 1, n)
xq = laps[-1]["x"][idx]; uq = laps[-1]["u"][idx]
A0 = np.tile(np.eye(6), (n, 1, 1)); B0 = np.zeros((n, 6, 2)); C0 = np.zeros((n, 6))
t = wall(lambda: mpc.regress(spec, xq, uq, A0, B0, C0), 3)
print(f"regression, {n} queries x 3 outputs over {sum(l['x'].shape[0]-1 for l in laps)} samples (host buffers): {t*1e3:.2f} ms", flush=True)
b1 = P.workload.make_batch(veh, cfg, 1024, 0xB200 + 2, tr, laps)
t0, *_ = dev_solve_time(mpc, b1)
mpc.set_error_dynamics(spec)
t1, ok, im, ix = dev_solve_time(mpc, b1)
mpc.set_error_dynamics(None)
print(f"tick, 1024 instances, device buffers: {t0*1e3:.3f} ms; with the regression on every stage: {t1*1e3:.3f} ms (solved {ok:.4f}, iters mean {im:.2f})", flush=True)
mpc.close()

# ---- configs[3]-like: 50 stored laps, 2 nearest per lap (48 laps searched), 2048 instances, with and without the regression
cfg4 = dict(cfg, num_ss_pts_per_lap=2, max_lap_stored=50)
mpc = BatchedRacingMPC(veh, cfg4, max_batch=2048)
many = P.workload.synthesise_laps(laps, 50)
for l in many: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
b4 = P.workload.make_batch(veh, cfg4, 2048, 0xB200 + 3, tr, laps)
t0, ok0, im0, ix0 = dev_solve_time(mpc, b4, 3)
mpc.set_error_dynamics(spec)
t1, ok1, im1, ix1 = dev_solve_time(mpc, b4, 3)
print(f"50-lap safe set ({sum(m_['x'].shape[0] for m_ in many)} points), 2 per lap, 2048 instances: {t0*1e3:.2f} ms (solved {ok0:.4f}, iters {im0:.2f}/{ix0}); "
      f"with the regression over all 50 laps: {t1*1e3:.2f} ms (solved {ok1:.4f})", flush=True)
mpc.close()

# ---- a12 / f4: SQP to convergence, 1024 instances of the BARC tracking problem
cfgt = P.configs.barc_tracking_config(20)
mpc = BatchedRacingMPC(veh, cfgt, max_batch=1024)
bt = P.workload.make_batch(veh, cfgt, 1024, 0xB200 + 1, tr, laps)
r = {}
def run_sqp(): r["o"] = mpc.solve_sqp(bt, max_sqp_iter=100, tol=1e-9)
t = wall(run_sqp, 2)
so = r["o"]; okq = so["status"] == 0
print(f"SQP to convergence (full dynamics), 1024 instances (host buffers): {t*1e3:.1f} ms, QP solves per instance mean {so['sqp_iters'].mean():.2f} max {so['sqp_iters'].max()} "
      f"(cap 100, step tolerance 1e-9); converged {okq.mean():.4f}; dynamics defect of those: median {np.median(so['defect'][okq]):.2e}, "
      f"99th percentile {np.percentile(so['defect'][okq], 99):.2e}, max {so['defect'][okq].max():.2e}; status histogram {np.bincount(so['status'], minlength=7)}", flush=True)
# ---- f1: track functions
mpc.set_track(tb)
s = np.random.default_rng(2).uniform(0, tr["length"], 1_000_000)
t = wall(lambda: mpc.track_eval(s), 2)
print(f"track interpolants, 1e6 abscissae (host buffers): {t*1e3:.1f} ms", flush=True)
f = np.column_stack([s, np.random.default_rng(3).uniform(-0.3, 0.3, s.size), np.zeros(s.size)])
g = mpc.frenet_to_global(f)
t = wall(lambda: mpc.global_to_frenet(g), 2)
print(f"global -> Frenet projection, 1e6 poses (host buffers): {t*1e3:.1f} ms", flush=True)
mpc.close()

# ---- configs[2]: IAC tracking, N = 40, 4096 instances
vi = P.configs.IAC_VEHICLE; ci = P.configs.iac_tracking_config(40)
tp = P.workload.load_track("putnam_optm")
mpc = BatchedRacingMPC(vi, ci, max_batch=4096)
bi = P.workload.make_batch(vi, ci, 4096, 0xB200 + 2, tp, laps, mode="track")
t, ok, im, ix = dev_solve_time(mpc, bi, 3)
print(f"IAC tracking N=40, 4096 instances, device buffers: {t*1e3:.2f} ms -> {4096/t:.3e} steps/s (solved {ok:.4f}, iters {im:.2f}/{ix})", flush=True)
mpc.close()
