"""Oracle restatement of the full-dynamics variant (RacingMPC(..., full_dynamics=true), racing_mpc.cpp:67-84,162-166):
SQP to convergence.  Checked here by what defines the answer: at a converged point the trajectory satisfies the
NONLINEAR dynamics, and it is a fixed point of "linearise -> solve the QP" (= the KKT conditions of the nonlinear
problem with the reference's cost and rows)."""
import numpy as np
import pytest

from conftest import make_oracle, relerr


@pytest.mark.parametrize("name,n,need", [("barc_tracking", 6, 6), ("barc_lmpc", 8, 6)])
def test_sqp_converges_to_a_kkt_point_of_the_nonlinear_problem(pkg, name, n, need):
    o, veh, cfg, track, mode = make_oracle(pkg, name, tol=1e-10)
    batch = pkg.workload.make_batch(veh, cfg, n, 0x5A9, track, pkg.workload.load_laps(), mode=mode)
    conv = 0
    for b in range(n):
        inp = pkg.workload.instance(batch, b)
        r = o.step_sqp(inp, max_sqp_iter=60, tol=1e-9)
        assert r["status"] == 0
        if r["sqp_iters"] >= 60:
            continue
        conv += 1
        assert r["defect"] < 1e-7, r["defect"]          # x_{i+1} = f_d(x_i, u_i, k_i, T_i)
        assert np.array_equal(r["X"][0], inp["x_ic"])
        # fixed point: the QP linearised at the answer returns the answer
        X0 = inp["X_ref"].copy()
        X0[:, 0] = [o.align_abscissa(s, inp["x_ic"][0], inp["total_length"]) for s in X0[:, 0]]
        again = o.step(dict(inp, X_ref=r["X"], U_ref=r["U"], ss_query_point=X0[-1, :2]), impl="port")
        assert again["status"] == 0
        assert max(relerr(again["X"], r["X"]), relerr(again["U"], r["U"])) < 1e-7
        # and it differs from the single linearised tick (otherwise the test would be vacuous)
        one = o.step(inp, impl="port")
        assert np.abs(one["X"] - r["X"]).max() > 1e-6
    assert conv >= need, conv
