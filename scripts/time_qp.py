"""Quick device-path timing of the solve for a few batch sizes (run with LMPC_WARPS_PER_INSTANCE=1|2|4)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np, torch
import racing_lmpc_ros2_b200 as P
from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
veh = P.configs.BARC_VEHICLE; cfg = P.configs.barc_lmpc_config(20)
laps = P.workload.load_laps(); tr = P.workload.load_track("barc_center")
for Bn in [int(a) for a in sys.argv[1:]] or [1024]:
    mpc = BatchedRacingMPC(veh, cfg, max_batch=Bn)
    for l in laps: mpc.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
    bb = P.workload.make_batch(veh, cfg, Bn, 0xB200 + 2, tr, laps)
    dev = {k: torch.from_numpy(v).cuda() for k, v in bb.items()}
    out = mpc.alloc_device_outputs(Bn)
    s = torch.cuda.Stream(); mpc.set_stream(s)
    with torch.cuda.stream(s):
        for _ in range(3): mpc.solve(dev, out)
    s.synchronize()
    mpc.set_timing(True)
    with torch.cuda.stream(s):
        for _ in range(10): mpc.solve(dev, out)
    (a, b, c), n = mpc.kernel_ms()
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    import hashlib
    chk = hashlib.sha1(out["X_optm"].cpu().numpy().tobytes() + out["U_optm"].cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"NW={os.environ.get('LMPC_WARPS_PER_INSTANCE','1')} B={Bn}: lin {a/n*1e3:.0f} us, ss {b/n*1e3:.0f} us, qp {c/n*1e3:.0f} us -> {Bn/((a+b+c)/n)*1e3:.3e} steps/s; solved {np.mean(st==0):.4f} iters mean {it.mean():.2f} max {it.max()} sha {chk}", flush=True)
    mpc.set_stream(None); mpc.close()
