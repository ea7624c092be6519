"""Pins the oracle's QP: two independent methods + KKT certificate, golden vectors, invariants,
and the structure-exploiting port against the dense statement."""
import os

import numpy as np
import pytest

from conftest import CASES, make_oracle, relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(name):
    z = np.load(os.path.join(GOLD, f"golden_{name}.npz"))
    batch = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    outs = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return batch, outs


@pytest.mark.parametrize("name", list(CASES))
def test_dense_oracle_reproduces_golden_and_certifies(pkg, name):
    o, veh, cfg, track, mode = make_oracle(pkg, name, tol=1e-11)
    batch, gold = _golden(name)
    for b in range(4):
        r = o.step(pkg.workload.instance(batch, b), impl="dense")
        assert r["status"] == 0 and r["polished"] == 1
        assert r["kkt"] < 1e-9                      # stationarity / feasibility / complementarity of the dense QP
        assert relerr(r["X"], gold["X"][b]) < 1e-9
        assert relerr(r["U"], gold["U"][b]) < 1e-9
        assert relerr(r["dU"], gold["dU"][b]) < 1e-9
        assert abs(r["cost"] - gold["cost"][b]) < 1e-9 * max(1, abs(gold["cost"][b]))


@pytest.mark.parametrize("name", list(CASES))
def test_invariants_of_the_optimum(pkg, name):
    o, veh, cfg, track, mode = make_oracle(pkg, name, tol=1e-11)
    batch, gold = _golden(name)
    N = cfg["N"]
    for b in range(3):
        inp = pkg.workload.instance(batch, b)
        r = o.step(inp, impl="dense")
        X, U, dU = r["X"], r["U"], r["dU"]
        assert np.allclose(X[0], inp["x_ic"], atol=1e-12)                         # x_0 = x_ic
        for i in range(N - 1):                                                     # dynamics rows hold
            xr = inp["X_ref"][i].copy()
            xr[0] = o.align_abscissa(xr[0], inp["x_ic"][0], inp["total_length"])
            A, B, g, _ = o.linearise(xr, inp["U_ref"][i], inp["curvatures"][i], inp["T_ref"][i])
            assert np.allclose(X[i + 1], A @ X[i] + B @ U[i] + g, atol=1e-9 * max(1, np.abs(X).max()))
            up = inp["u_ic"] if i == 0 else U[i - 1]
            assert np.allclose(U[i], up + inp["T_ref"][i] * dU[i], atol=1e-11)    # rate rows hold
        assert (U <= np.array(cfg["u_max"]) + 1e-9).all() and (U >= np.array(cfg["u_min"]) - 1e-9).all()
        if cfg["learning"]:
            lam = r["lam"]
            assert abs(lam.sum() - 1) < 1e-10 and lam.min() > -1e-10
            assert np.allclose(X[N - 1] - r["ss_x"].T @ lam, r["sigma_h"], atol=1e-9)
            assert (np.abs(lam) > 1e-8).sum() <= 7   # vertex / low-dimensional face of the hull
        cost, inf = o.check_candidate(inp, X, U, dU, r["lam"])
        assert inf < 1e-9 and abs(cost - r["cost"]) < 1e-9 * max(1, abs(cost))


@pytest.mark.parametrize("name", list(CASES))
def test_port_matches_dense(pkg, name):
    """The Riccati/IPM port (the timed CPU baseline and the CUDA kernel's CPU counterpart) against the
    certified dense optimum."""
    o, veh, cfg, track, mode = make_oracle(pkg, name, tol=1e-13)
    od, *_ = make_oracle(pkg, name, tol=1e-11)
    batch = pkg.workload.make_batch(veh, cfg, 12, 0x5EED, track, pkg.workload.load_laps(), mode=mode)
    worst = 0.0
    for b in range(12):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        if not (d["status"] == 0 and d["polished"] == 1 and d["kkt"] < 1e-9):
            continue
        p = o.step(inp, impl="port")
        assert p["status"] == 0
        e = max(relerr(p["X"], d["X"]), relerr(p["U"], d["U"]), relerr(p["dU"], d["dU"]))
        worst = max(worst, e)
        assert abs(p["cost"] - d["cost"]) < 1e-8 * max(1, abs(d["cost"]))
        if cfg["learning"]:   # lambda is not unique; SS*lambda is
            assert np.allclose(p["ss_x"].T @ p["lam"], d["ss_x"].T @ d["lam"], atol=1e-7)
    assert worst < 1e-6, worst   # north-star tolerance; typical is 1e-10


def test_infeasible_initial_state_and_missing_safe_set(pkg):
    o, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc", tol=1e-11)
    batch = pkg.workload.make_batch(veh, cfg, 1, 7, track, pkg.workload.load_laps(), mode=mode)
    inp = pkg.workload.instance(batch, 0)
    bad = dict(inp); bad["x_ic"] = inp["x_ic"].copy(); bad["x_ic"][3] = 0.05   # v_x below x_min (racing_mpc.cpp:147)
    assert o.step(bad, impl="dense")["status"] == 2
    assert o.step(bad, impl="port")["status"] == 2
    o2, *_ = make_oracle(pkg, "barc_lmpc", with_laps=False)
    assert o2.step(inp, impl="dense")["status"] == 3
    assert o2.step(inp, impl="port")["status"] == 3


def test_boundary_slack_start_keeps_long_horizons_short(pkg):
    """Start-point rule shared by port and kernel: a rollout that leaves the track starts with sigma_b covering the
    violation.  IAC tracking N = 40 (BASELINE configs[2]), 256 instances of the full-size test's seed: every instance
    solves, no instance needs more than 10 iterations (16 before the rule), mean below 6."""
    o, veh, cfg, track, mode = make_oracle(pkg, "iac_tracking")
    batch = pkg.workload.make_batch(veh, cfg, 256, 0xB200 + 3, track, pkg.workload.load_laps(), mode=mode)
    r = o.step_batch(batch, impl="port", nthreads=4)
    assert (r["status"] == 0).all()
    assert r["iters"].max() <= 10 and r["iters"].mean() < 6.0, (r["iters"].max(), r["iters"].mean())
