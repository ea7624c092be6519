"""Generates tests/golden/golden_<case>.npz : seeded inputs + the certified optimum of the reference
QP for each case, computed by the dense CPU oracle (IPM + active-set polish, KKT <= 1e-9).

The reference stack (CasADi/OSQP) cannot run in this environment, so these are golden vectors of the
ORACLE (parity unpinned, see oracle/lmpc_oracle.h); they pin the oracle, the port and the CUDA path
against silent drift.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import CASES, make_oracle  # noqa: E402
import racing_lmpc_ros2_b200 as P  # noqa: E402

NB = 8
for ci, name in enumerate(CASES):
    o, veh, cfg, track, mode = make_oracle(P, name, tol=1e-11)
    batch = P.workload.make_batch(veh, cfg, NB, 0x601D + ci, track, P.workload.load_laps(), mode=mode)
    outs = dict(X=[], U=[], dU=[], cost=[], sslam=[], kkt=[], sigma_b=[])
    for b in range(NB):
        r = o.step(P.workload.instance(batch, b), impl="dense")
        assert r["status"] == 0 and r["polished"] == 1 and r["kkt"] < 1e-9, (name, b, r["status"], r["polished"], r["kkt"])
        outs["X"].append(r["X"]); outs["U"].append(r["U"]); outs["dU"].append(r["dU"]); outs["cost"].append(r["cost"])
        outs["sslam"].append(r["ss_x"].T @ r["lam"] if cfg["learning"] else np.zeros(6))
        outs["kkt"].append(r["kkt"]); outs["sigma_b"].append(r["sigma_b"])
    np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **{f"in_{k}": v for k, v in batch.items()},
                        **{f"out_{k}": np.array(v) for k, v in outs.items()})
    print(name, "kkt max", max(outs["kkt"]), "cost", np.round(outs["cost"], 6))

# the two shipped parameter sets outside BASELINE's configs (conftest.make_extra_case): same file layout; the IAC LMPC
# laps are synthesised deterministically from the tracked Putnam table and are not stored
from conftest import make_extra_case  # noqa: E402
from oracle import Oracle  # noqa: E402

for ci, (name, nb) in enumerate((("hawaii_kart_tracking", 8), ("iac_lmpc", 4))):
    veh, cfg, track, dt, laps = make_extra_case(P, name)
    o = Oracle(veh, dict(cfg, tol=1e-11))
    for l in laps or []:
        o.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    batch = P.workload.make_batch(veh, cfg, nb, 0x601D + 16 + ci, track, laps, dt=dt, mode="track")
    outs = dict(X=[], U=[], dU=[], cost=[], sslam=[], kkt=[], sigma_b=[])
    for b in range(nb):
        r = o.step(P.workload.instance(batch, b), impl="dense")
        assert r["status"] == 0 and r["polished"] == 1 and r["kkt"] < 1e-9, (name, b, r["status"], r["polished"], r["kkt"])
        outs["X"].append(r["X"]); outs["U"].append(r["U"]); outs["dU"].append(r["dU"]); outs["cost"].append(r["cost"])
        outs["sslam"].append(r["ss_x"].T @ r["lam"] if cfg["learning"] else np.zeros(6))
        outs["kkt"].append(r["kkt"]); outs["sigma_b"].append(r["sigma_b"])
    np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **{f"in_{k}": v for k, v in batch.items()},
                        **{f"out_{k}": np.array(v) for k, v in outs.items()})
    print(name, "kkt max", max(outs["kkt"]), "cost", np.round(outs["cost"], 6))
