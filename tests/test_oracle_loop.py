"""Pins the oracle's restatement of the steps either side of the solve (oracle/oracle_loop.py: tick preparation,
actuation, plant, closed loop) and of the full-dynamics variant against the committed golden
vectors (generated with the DENSE certified QP; here the structure-exploiting port runs inside), plus the invariants
the reference's code implies."""
import math
import os

import numpy as np
import pytest

from conftest import ROOT, make_oracle, relerr

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def barc(pkg):
    from oracle_track import OracleTrack
    tb = np.load(os.path.join(GOLD, "tracks.npz"))["barc_center_table"]
    o, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc", tol=1e-10)
    return o, OracleTrack(tb), cfg


def test_closed_loop_port_reproduces_dense_golden(barc):
    import oracle_loop as OL
    o, trk, cfg = barc
    z = np.load(os.path.join(GOLD, "golden_closed_loop_barc.npz"))
    opt = dict(step_mode="step", delay_step=0, plant_substeps=int(z["plant_substeps"]), dt=float(z["dt"]), plant_dt=float(z["plant_dt"]),
               speed_limit=1e9, speed_scale=1.0, max_vel_ref_diff=1.0)
    for b in range(z["x0"].shape[0]):
        r = OL.closed_loop(o, trk, opt, int(z["ticks"]), z["x0"][b], z["u0"][b], z["X0"][b], z["U0"][b], impl="port")
        assert r["fail_count"] == 0 and r["lap_count"] == z["lap_count"][b]
        assert max(relerr(r["log_x"], z["log_x"][:, b]), relerr(r["log_u"], z["log_u"][:, b]), relerr(r["X_last"], z["X_last"][b])) < 1e-7


def test_prepare_semantics(barc):
    """racing_mpc_node.cpp:236-292 by its own invariants."""
    import oracle_loop as OL
    o, trk, cfg = barc
    z = np.load(os.path.join(GOLD, "golden_closed_loop_barc.npz"))
    x, u, X, U = z["x0"][0], z["u0"][0], z["X0"][0], z["U0"][0]
    base = dict(delay_step=0, plant_substeps=1, dt=0.025, plant_dt=0.025, speed_limit=1.2, speed_scale=1.0, max_vel_ref_diff=0.25)
    st = OL.prepare(o, trk, dict(base, step_mode="step"), x, u, X, U)
    ct = OL.prepare(o, trk, dict(base, step_mode="continuous"), x, u, X, U)
    assert np.array_equal(st["x_ic"], x) and np.array_equal(st["u_ic"], u)
    assert np.array_equal(ct["x_ic"], OL.step_on_track(o, trk, x, U[0], 0.025))            # one step ahead with last_u[0]
    assert np.array_equal(st["X_ref"][:-1], X[1:]) and np.array_equal(st["U_ref"][:-1], U[1:]) and np.array_equal(st["U_ref"][-1], U[-1])
    assert np.array_equal(st["X_ref"][-1], OL.step_on_track(o, trk, X[-1], U[-1], 0.025))  # extension by the dynamics
    vx = st["X_ref"][:, 3]
    assert (np.abs(st["vel_ref"] - vx) <= 0.25 + 1e-12).all()                               # within +-max_vel_ref_diff of the speed
    assert (st["vel_ref"] <= np.maximum(1.2, vx - 0.25) + 1e-12).all()                      # and capped by the speed limit
    e = trk.eval(st["X_ref"][:, 0])
    assert np.array_equal(st["bound_left"], e["left"]) and (st["bound_left"] > 0).all() and (st["bound_right"] < 0).all()


def test_actuation_and_plant(barc):
    import oracle_loop as OL
    o, trk, cfg = barc
    for ul in (0.01, -0.01, 0.0, 3.0, -7.0):
        ua = OL.actuation(np.array([ul, 0.1]))
        assert ua[1] == 0.1 and abs(ua[0] - ul / (1.0 + math.exp(-abs(ul)))) < 1e-15       # u sigma(|u|): the larger of (Fd, Fb)
    L = trk.L
    x = np.array([L - 0.01, 0.0, 0.0, 1.5, 0.0, 0.0])
    xn, laps = OL.plant_step(o, trk, dict(plant_substeps=2, plant_dt=0.0125), x, np.array([0.0, 0.0]), 0)
    assert laps == 1 and 0.0 <= xn[0] < 0.1                                                 # wrapped across the start line
    xs, _ = OL.plant_step(o, trk, dict(plant_substeps=1, plant_dt=0.01), np.array([1.0, 0, 0, 0.0, 0, 0]), np.array([0.0, 0.0]), 0)
    assert np.isfinite(xs).all()                                                            # v_x = 0 is floored to 1e-6 first


@pytest.mark.parametrize("name", ["barc_tracking", "iac_tracking"])
def test_sqp_port_reproduces_dense_golden(pkg, name):
    o, veh, cfg, track, mode = make_oracle(pkg, name, tol=1e-10)
    z = np.load(os.path.join(GOLD, f"golden_sqp_{name}.npz"))
    batch = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    for b in range(batch["x_ic"].shape[0]):
        r = o.step_sqp(pkg.workload.instance(batch, b), max_sqp_iter=80, tol=1e-10)
        assert r["status"] == 0 and r["sqp_iters"] < 80
        assert max(relerr(r["X"], z["out_X"][b]), relerr(r["U"], z["out_U"][b]), relerr(r["dU"], z["out_dU"][b])) < 1e-7
