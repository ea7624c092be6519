"""CPU ORACLE for the steps either side of the solve -- TEST INFRASTRUCTURE ONLY (see oracle/lmpc_oracle.h).

Plain-Python restatement, one agent at a time, of
  prepare      RacingMPCNode::on_step_timer's input preparation   mpc/racing_mpc/src/racing_mpc_node.cpp:236-292
  actuation    to_base_control + the actuation message            racing_mpc_node.cpp:386-401, single_track_planar_model.cpp:395-407
  plant_step   RacingSimulator::step + the node's lap counter     simulation/racing_simulator/src/racing_simulator.cpp:97-113,
                                                                  racing_simulator_node.cpp:241-286
  closed_loop  the three around Oracle.step (the tick's QP)
"""
import math

import numpy as np


def step_on_track(orc, trk, x, u, dt):
    """the node's / simulator's `discrete_dynamics_`: curvature from the track at the state's abscissa
    (racing_mpc_node.cpp:69-76, racing_simulator.cpp:46-57)"""
    k = float(trk.eval(np.array([x[0]]))["curvature"][0])
    return orc.discrete_dynamics(x, u, k, dt)


def prepare(orc, trk, opt, x, u_prev, X_last, U_last):
    N = X_last.shape[0]
    x_ic = step_on_track(orc, trk, x, U_last[0], opt["dt"]) if opt["step_mode"] == "continuous" else np.array(x, dtype=float)  # :238-244
    Xr = np.vstack([X_last[1:], np.zeros((1, 6))])                                  # :245
    Ur = np.vstack([U_last[1:], U_last[-1:]])                                       # :246
    Xr[-1] = step_on_track(orc, trk, Xr[-2], Ur[-1], opt["dt"])                     # :248-249
    e = trk.eval(Xr[:, 0])                                                          # :261-265
    vel = np.zeros(N)
    for i in range(N):                                                              # :269-287
        cur = Xr[i, 3]
        ref = e["vel"][i] * opt["speed_scale"]
        lo, hi = cur - opt["max_vel_ref_diff"], cur + opt["max_vel_ref_diff"]
        lim = min(max(opt["speed_limit"], lo), hi)
        vel[i] = min(min(max(ref, lo), hi), lim) if ref > 0.0 else lim
    return dict(x_ic=x_ic, u_ic=np.array(u_prev, dtype=float), X_ref=Xr, U_ref=Ur, T_ref=np.full(N - 1, opt["dt"]),
                bound_left=e["left"], bound_right=e["right"], curvatures=e["curvature"], vel_ref=vel, total_length=trk.L)


def actuation(u):
    fd = u[0] * 1.0 / (1.0 + math.exp(-u[0]))                                       # single_track_planar_model.cpp:395-400
    fb = u[0] * 1.0 / (1.0 + math.exp(u[0]))
    return np.array([fd if abs(fd) > abs(fb) else fb, u[1]])                        # racing_mpc_node.cpp:397-401


def plant_step(orc, trk, opt, x, ua, laps):
    x = np.array(x, dtype=float)
    for _ in range(opt["plant_substeps"]):
        if abs(x[3]) < 1e-6:                                                        # racing_simulator.cpp:99-103
            x[3] = math.copysign(1e-6, x[3])
        xn = step_on_track(orc, trk, x, ua, opt["plant_dt"])
        xn[0] = float(trk.wrap(xn[0]))                                              # :58-62
        if x[0] - xn[0] > 0.5 * trk.L:                                              # racing_simulator_node.cpp:283-286
            laps += 1
        x = xn
    return x, laps


def closed_loop(orc, trk, opt, ticks, x, u_prev, X_last, U_last, impl="port", recorder=None, t0=0.0):
    """recorder: an oracle_regress.Recorder -- the agent then learns from its own laps: every tick feeds the recorder with
    (x_ic, u_ic, curvatures(0), t_ic) before the safe-set query, and a completed lap joins the agent's safe set (`orc`
    must then be the agent's own Oracle), as RacingMPC::solve does (racing_mpc.cpp:245-255)."""
    x = np.array(x, dtype=float); u_prev = np.array(u_prev, dtype=float)
    X_last = np.array(X_last, dtype=float); U_last = np.array(U_last, dtype=float)
    laps, fails = 0, 0
    log_x, log_u = [], []
    for tick in range(ticks):
        inp = prepare(orc, trk, opt, x, u_prev, X_last, U_last)
        if recorder is not None:
            if recorder.step(inp["x_ic"], inp["u_ic"], float(inp["curvatures"][0]), t0 + opt["dt"] * tick, trk.L):
                l = recorder.laps[-1]
                orc.add_lap(l["x"], l["u"], l["k"], l["t"], trk.L)
        r = orc.step(inp, impl=impl)
        if r["status"] == 0:                                                        # racing_mpc_node.cpp:322-331
            X_last, U_last = r["X"], r["U"]
        else:
            X_last, U_last = inp["X_ref"], inp["U_ref"]; fails += 1
        ua = actuation(U_last[opt["delay_step"]])
        x, laps = plant_step(orc, trk, opt, x, ua, laps)
        u_prev = ua
        log_x.append(x.copy()); log_u.append(ua.copy())
    return dict(x=x, u_prev=u_prev, X_last=X_last, U_last=U_last, lap_count=laps, fail_count=fails,
                log_x=np.array(log_x), log_u=np.array(log_u))
