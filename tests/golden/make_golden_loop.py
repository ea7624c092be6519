"""Generates tests/golden/golden_closed_loop_barc.npz and golden_sqp_<case>.npz with the CPU oracle:

  closed loop : 6 agents x 10 ticks of prepare -> QP (dense oracle: IPM + polish + KKT certificate) -> actuation ->
                plant on the BARC track with the recorded safe set (oracle/oracle_loop.py, oracle/oracle_track.py)
  SQP         : the full-dynamics variant (racing_mpc.cpp:67-84) converged with the dense QP inside

Golden vectors of the ORACLE (parity unpinned, see oracle/lmpc_oracle.h); they pin oracle, port and CUDA path against
drift.  Run in the build container:  python tests/golden/make_golden_loop.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import make_oracle  # noqa: E402
import racing_lmpc_ros2_b200 as P  # noqa: E402
import oracle_loop as OL  # noqa: E402
from oracle_track import OracleTrack  # noqa: E402

tb = np.load(os.path.join(HERE, "tracks.npz"))["barc_center_table"]
trk = OracleTrack(tb)
laps = P.workload.load_laps()
o, veh, cfg, track, mode = make_oracle(P, "barc_lmpc", tol=1e-11)
NB, N, DT, TICKS = 6, cfg["N"], 0.025, 10
rng = np.random.default_rng(0x100B)
lap = laps[-1]
j0 = rng.integers(0, lap["x"].shape[0] - N - 1, NB)
x0 = lap["x"][j0] + rng.standard_normal((NB, 6)) * np.array([0.02, 0.01, 0.01, 0.02, 0.01, 0.02])
u0 = lap["u"][j0].copy()
U0 = np.stack([lap["u"][j:j + N - 1] for j in j0])
X0 = np.zeros((NB, N, 6)); X0[:, 0] = x0
for b in range(NB):
    for i in range(N - 1):
        X0[b, i + 1] = OL.step_on_track(o, trk, X0[b, i], U0[b, i], DT)
opt = dict(step_mode="step", delay_step=0, plant_substeps=2, dt=DT, plant_dt=DT / 2, speed_limit=1e9, speed_scale=1.0, max_vel_ref_diff=1.0)
res = [OL.closed_loop(o, trk, opt, TICKS, x0[b], u0[b], X0[b], U0[b], impl="dense") for b in range(NB)]
assert all(r["fail_count"] == 0 for r in res)
np.savez_compressed(os.path.join(HERE, "golden_closed_loop_barc.npz"), x0=x0, u0=u0, X0=X0, U0=U0, dt=DT, plant_dt=DT / 2,
                    plant_substeps=2, ticks=TICKS, log_x=np.stack([r["log_x"] for r in res], axis=1),
                    log_u=np.stack([r["log_u"] for r in res], axis=1), X_last=np.stack([r["X_last"] for r in res]),
                    lap_count=np.array([r["lap_count"] for r in res]))
print("closed loop: final s", np.round([r["x"][0] for r in res], 4))

for ci, name in enumerate(("barc_tracking", "iac_tracking")):
    o, veh, cfg, track, mode = make_oracle(P, name, tol=1e-11)
    batch = P.workload.make_batch(veh, cfg, 4, 0x5A90 + ci, track, laps, mode=mode)
    outs = dict(X=[], U=[], dU=[], sqp_iters=[], defect=[])
    for b in range(4):
        r = o.step_sqp(P.workload.instance(batch, b), max_sqp_iter=80, tol=1e-10, impl="dense")
        assert r["status"] == 0 and r["sqp_iters"] < 80 and r["defect"] < 1e-8, (name, b, r["status"], r["sqp_iters"], r["defect"])
        for k in outs:
            outs[k].append(r[k])
    np.savez_compressed(os.path.join(HERE, f"golden_sqp_{name}.npz"), **{f"in_{k}": v for k, v in batch.items()},
                        **{f"out_{k}": np.array(v) for k, v in outs.items()})
    print(name, "sqp iters", outs["sqp_iters"], "defect", np.array(outs["defect"]))
