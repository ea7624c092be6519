"""The full-dynamics solve (racing_mpc.cpp:67-84,162-166) against an INDEPENDENT nonlinear-programming solve.

The oracle's SQP (oracle/oracle_port.c: orc_step_sqp, the iteration the device runs as lmpc_solve_sqp_batch) iterates the
tick's QP; scipy's SLSQP knows nothing of that QP: it gets the problem as the reference states it to IPOPT -- the tracking
cost, the NONLINEAR discrete dynamics as equality constraints, the boxes, the soft track boundary with its slack -- in the
variables (X, U, sigma_b), with finite-difference derivatives.  Small horizon so that SLSQP converges tightly."""
import numpy as np
import pytest
from scipy.optimize import minimize

from conftest import make_case


def _nlp_solve(o, veh, cfg, inp):
    N = cfg["N"]
    nx, nu = 6 * N, 2 * (N - 1)
    T, kap, bl, br, vref = inp["T_ref"], inp["curvatures"], inp["bound_left"], inp["bound_right"], inp["vel_ref"]
    m = cfg["margin"] + veh["chassis_b"] / 2
    R = np.array(cfg["R"]).reshape(2, 2); Rd = np.array(cfg["R_d"]).reshape(2, 2)
    w = np.array([0.0, cfg["q_contour"], cfg["q_heading"], cfg["q_vel"], cfg["q_vy"], cfg["q_vyaw"]])
    ulo = np.maximum(cfg["u_min"], [veh["Fb_max"] / 1000.0, -veh["max_steer"]]); uhi = np.minimum(cfg["u_max"], [veh["Fd_max"] / 1000.0, veh["max_steer"]])
    dlo = np.array([veh["Fb_max"] / 1000.0 / veh["Tb"], -veh["max_steer_rate"]]); dhi = np.array([veh["Fd_max"] / 1000.0 / veh["Td"], veh["max_steer_rate"]])

    def split(z):
        return z[:nx].reshape(N, 6), z[nx:nx + nu].reshape(N - 1, 2), z[-1]

    def rates(U):
        up = np.vstack([inp["u_ic"][None, :], U[:-1]])
        return (U - up) / T[:, None]

    def cost(z):                                  # racing_mpc.cpp:442-477 (tracking) + :539 (boundary slack)
        X, U, sb = split(z)
        dU = rates(U)
        c = cfg["q_boundary"] * sb * sb + np.einsum("ia,ab,ib->", U, R, U) + np.einsum("ia,ab,ib->", dU, Rd, dU)
        e = X.copy(); e[:, 3] -= vref
        c += (w[None, 1:] * e[:-1, 1:] ** 2).sum()
        c += 10.0 * (w[1] * e[-1, 1] ** 2 + w[2] * e[-1, 2] ** 2 + w[3] * e[-1, 3] ** 2)
        return c

    def eq(z):                                    # x_0 = x_ic; x_{i+1} = f_d(x_i, u_i)  (:162-166, :199-201)
        X, U, _ = split(z)
        r = [X[0] - inp["x_ic"]]
        for i in range(N - 1):
            r.append(X[i + 1] - o.discrete_dynamics(X[i], U[i], kap[i], T[i]))
        return np.concatenate(r)

    xmax = np.array(cfg["x_max"]); xmin = np.array(cfg["x_min"])

    def ineq(z):                                  # >= 0
        X, U, sb = split(z)
        dU = rates(U)
        r = [np.array([sb])]
        for k in range(6):
            if np.isfinite(xmax[k]): r.append(xmax[k] - X[1:N - 1, k])
            if np.isfinite(xmin[k]): r.append(X[1:N - 1, k] - xmin[k])
        r += [(uhi - U).ravel(), (U - ulo).ravel(), (dhi - dU).ravel(), (dU - dlo).ravel()]
        r += [bl - m + sb - X[:, 1], X[:, 1] - (br + m - sb)]
        return np.concatenate(r)

    Xr = inp["X_ref"].copy()
    for j in range(N):
        Xr[j, 0] = o.align_abscissa(Xr[j, 0], inp["x_ic"][0], inp["total_length"])
    z0 = np.concatenate([Xr.ravel(), inp["U_ref"].ravel(), [0.01]])
    res = minimize(cost, z0, method="SLSQP", constraints=[dict(type="eq", fun=eq), dict(type="ineq", fun=ineq)],
                   options=dict(ftol=1e-15, maxiter=400))
    X, U, sb = split(res.x)
    return res, X, U, cost(res.x), np.abs(eq(res.x)).max(), min(ineq(res.x).min(), 0.0)


@pytest.mark.parametrize("seed", [3, 11])
def test_oracle_sqp_matches_an_independent_nlp_solve(pkg, seed):
    from oracle import Oracle
    N = 6
    veh, cfg, track, mode = make_case(pkg, "barc_tracking", None, N)
    o = Oracle(veh, dict(cfg, tol=1e-11))
    batch = pkg.workload.make_batch(veh, cfg, 4, seed, track, pkg.workload.load_laps(), mode=mode)
    n = 0
    for b in range(4):
        inp = pkg.workload.instance(batch, b)
        r = o.step_sqp(inp, max_sqp_iter=100, tol=1e-11)
        assert r["status"] == 0 and r["defect"] < 1e-10, (b, r["status"], r["defect"], r["sqp_iters"])
        res, X, U, c, eqv, inv = _nlp_solve(o, veh, cfg, inp)
        # SLSQP's own exit flag is not used: with finite-difference derivatives and ftol 1e-15 it ends on "positive
        # directional derivative for linesearch" at the solution; what counts is that its point is feasible
        assert eqv < 1e-8 and inv > -1e-8, (b, res.message, eqv, inv)
        scale_x = np.maximum(1.0, np.abs(r["X"]).max(axis=0)); scale_u = np.maximum(1.0, np.abs(r["U"]).max(axis=0))
        ex = (np.abs(X - r["X"]).max(axis=0) / scale_x).max(); eu = (np.abs(U - r["U"]).max(axis=0) / scale_u).max()
        # both are local solutions of the same problem from the same start: same point (SLSQP with finite-difference
        # derivatives lands within 1e-6 of it), same value
        assert abs(c - r["cost"]) < 1e-8 * max(1.0, abs(c)), (b, c, r["cost"])
        assert ex < 1e-5 and eu < 1e-5, (b, ex, eu)
        n += 1
    assert n == 4
