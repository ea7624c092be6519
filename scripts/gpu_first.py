"""First-contact GPU script: parity of the three kernels vs the CPU oracle + a rough timing."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
warnings.filterwarnings("ignore")
import numpy as np
import torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
from oracle import Oracle

def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float((np.abs(a - b).reshape(-1, a.shape[-1]).max(axis=0) / np.maximum(1, np.abs(b).reshape(-1, b.shape[-1]).max(axis=0))).max())

print(torch.cuda.get_device_name(0), flush=True)
veh = P.configs.BARC_VEHICLE; cfg = P.configs.barc_lmpc_config(20); cfg["tol"] = 1e-13
laps = P.workload.load_laps(); tr = P.workload.load_track("barc_center")
mpc = BatchedRacingMPC(veh, cfg, max_batch=8192)
orc = Oracle(veh, cfg)
for l in laps:
    mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"]); orc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
# --- model
rng = np.random.default_rng(0)
x = np.column_stack([rng.uniform(0, 17, 256), rng.uniform(-.3, .3, 256), rng.uniform(-.3, .3, 256), rng.uniform(.2, 3, 256), rng.uniform(-.5, .5, 256), rng.uniform(-2, 2, 256)])
u = np.column_stack([rng.uniform(-.01, .01, 256), rng.uniform(-.3, .3, 256)]); kap = rng.uniform(-1, 1, 256)
A, Bm, g, xn = mpc.linearise(x, u, kap, 0.025)
e = 0
for i in range(256):
    A2, B2, g2, xn2 = orc.linearise(x[i], u[i], kap[i], 0.025)
    e = max(e, np.abs(A[i] - A2).max(), np.abs(Bm[i] - B2).max(), np.abs(g[i] - g2).max(), np.abs(xn[i] - xn2).max())
print("linearise max abs diff vs oracle", e, flush=True)
# --- safe set
q = np.column_stack([rng.uniform(-5, 25, 64), rng.uniform(-.4, .4, 64)])
sx, sj = mpc.ss_query(q)
bad = 0
for i in range(64):
    ox, oj = orc.ss_query(q[i, 0], q[i, 1])
    if not (np.array_equal(ox, sx[i]) and np.array_equal(oj, sj[i])): bad += 1
print("ss_query mismatching queries", bad, "of 64; count", sx.shape[1], flush=True)
# --- solve parity
b = P.workload.make_batch(veh, cfg, 64, 0xB202, tr, laps)
out = mpc.solve(b)
ref = orc.step_batch(b, impl="port", nthreads=8)
print("status", np.bincount(out["status"], minlength=5), "iters mean", out["iters"].mean(), "oracle iters", ref["iters"].mean())
print("parity vs port: X %.2e U %.2e dU %.2e cost %.2e" % (relerr(out["X_optm"], ref["X"]), relerr(out["U_optm"], ref["U"]), relerr(out["dU_optm"], ref["dU"]), np.abs(out["cost"] - ref["cost"]).max()), flush=True)
d = orc.step(P.workload.instance(b, 0))
print("vs dense[0]: X %.2e U %.2e" % (relerr(out["X_optm"][0], d["X"]), relerr(out["U_optm"][0], d["U"])), flush=True)
# --- timing, device path
for Bn in (1024, 4096, 8192):
    bb = P.workload.make_batch(veh, cfg, Bn, 0xB200 + 2, tr, laps)
    dev = {k: torch.from_numpy(v).cuda() for k, v in bb.items()}
    outd = mpc.alloc_device_outputs(Bn)
    s = torch.cuda.Stream(); mpc.set_stream(s)
    with torch.cuda.stream(s):
        for _ in range(3): mpc.solve(dev, outd)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): mpc.solve(dev, outd)
        e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1) / 10
    st = outd["status"].cpu().numpy()
    print(f"B={Bn}: {ms:.3f} ms/batch -> {Bn / ms * 1e3:.3e} steps/s; status {np.bincount(st, minlength=5)} iters {outd['iters'].float().mean().item():.2f}", flush=True)
    mpc.set_stream(None)
t0 = time.time(); ref = orc.step_batch(P.workload.make_batch(veh, cfg, 256, 1, tr, laps), impl="port", nthreads=os.cpu_count()); t1 = time.time()
print(f"cpu port: {256 / (t1 - t0):.1f} steps/s on {os.cpu_count()} threads")
