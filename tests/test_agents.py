"""Per-agent safe sets and device-side lap recording (lmpc_agents_*, csrc/lmpc_agents.cuh): every agent of a closed
loop owns a SafeSetRecorder and a SafeSetManager as every RacingMPC does in the reference (racing_mpc.hpp:99-100,
racing_mpc.cpp:245-255, safe_set.cpp:116-180,278-322)."""
import os

import numpy as np
import pytest

from conftest import ROOT, make_oracle, relerr

GOLD = os.path.join(ROOT, "tests", "golden")


def _track_table(name):
    return np.ascontiguousarray(np.load(os.path.join(GOLD, "tracks.npz"))[f"{name}_table"], dtype=np.float64)


def _start_near_the_line(pkg, veh, otrk, o, laps, nb, N, dt, seed, back=(30, 130)):
    """agents start `back` samples before the end of the newest recorded lap: the first wrap (which arms the recorder)
    comes within a few seconds, the second one a lap later completes the first recorded lap"""
    import oracle_loop as OL
    rng = np.random.default_rng(seed)
    lap = laps[-1]
    n = lap["x"].shape[0]
    j0 = n - rng.integers(back[0], back[1], nb)
    x = lap["x"][j0] + rng.standard_normal((nb, 6)) * np.array([0.02, 0.01, 0.01, 0.02, 0.01, 0.02])
    u_prev = lap["u"][j0].copy()
    X_last = np.zeros((nb, N, 6)); U_last = np.zeros((nb, N - 1, 2))
    for b in range(nb):
        idx = (j0[b] + np.arange(N - 1)) % n
        U_last[b] = lap["u"][idx]
        X_last[b, 0] = x[b]
        for i in range(N - 1):
            X_last[b, i + 1] = OL.step_on_track(o, otrk, X_last[b, i], U_last[b, i], dt)
    return x, u_prev, X_last, U_last


@pytest.mark.gpu
def test_device_recorder_and_per_agent_safe_sets_equal_the_host_recorder_bit_for_bit(pkg, laps):
    """560 ticks of 6 agents on the device, each recording its own laps.  What every agent's recorder was fed is logged
    and replayed (a) through the host recorder of the C ABI (lmpc_recorder_step -> lmpc_safe_set_add_lap, one fresh handle
    per agent) and (b) through the Python restatement of SafeSetRecorder::step: lap counters, the recorded laps and the
    answers of the safe-set query on the agent's own laps must be IDENTICAL (tripling, cost-to-go, circular buffer of
    max_lap_stored laps, newest-first order, duplicate keys)."""
    from oracle_regress import Recorder
    from oracle_track import OracleTrack
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    tb = _track_table("barc_center")
    o, veh, cfg, track, _ = make_oracle(pkg, "barc_lmpc", tol=1e-10)
    nb, N, dt, ticks = 6, cfg["N"], 0.025, 560
    m = BatchedRacingMPC(veh, cfg, max_batch=nb)
    for l in laps:
        m.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    m.set_track(tb)
    otrk = OracleTrack(tb)
    m.agents_create(nb, 1024)
    x, u_prev, X_last, U_last = _start_near_the_line(pkg, veh, otrk, o, laps, nb, N, dt, 8)
    opt = m.loop_options(dt)
    out = m.closed_loop_agents(opt, ticks, x, u_prev, X_last, U_last, t0=100.0)
    st = m.agents_status()
    assert (st["flags"] & 8 == 0).all()
    assert (out["fail_count"] <= 2).all(), out["fail_count"]
    assert (st["lap_count"] >= 3 + 2).all(), st["lap_count"]            # three loaded laps + two wraps
    rng = np.random.default_rng(1)
    queries = np.column_stack([rng.uniform(-5, 25, 40), rng.uniform(-.5, .5, 40)])
    for b in range(nb):
        rec = Recorder(); rec.lap_count = 3
        h = BatchedRacingMPC(veh, cfg, max_batch=40)
        for l in laps:
            h.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
        added = 0
        for t in range(ticks):
            s = out["log_rec"][t, b]
            a1 = rec.step(s[:6], s[6:8], s[8], s[9], track["length"])
            a2 = h.recorder_step(s[:6], s[6:8], s[8], s[9], track["length"])
            assert a1 == a2
            added += int(a2)
        assert added >= 1 and st["lap_count"][b] == rec.lap_count == h.recorder_lap_count() + 3
        assert st["stored"][b] == min(cfg["max_lap_stored"], 3 + added) == h.num_laps()
        for which in range(min(added, cfg["max_lap_stored"])):
            got = m.agents_get_lap(b, which)
            assert np.array_equal(got, rec.laps[-1 - which]["x"]), (b, which)
        # the query on the agent's own laps == the shared-set query of a handle holding exactly those laps
        for q in queries:
            ag = m.agents_query(np.tile(q, (nb, 1)))[b]
            sx, sj = h.ss_query(q[None, :])
            assert np.array_equal(ag[0], sx[0]) and np.array_equal(ag[1], sj[0]), (b, q)
        h.close()
    print(f"device recorder: {nb} agents x {ticks} ticks, laps recorded per agent {st['lap_count'] - 3}, stored {st['stored']}: laps and queries identical to the host recorder")


@pytest.mark.gpu
def test_closed_loop_with_per_agent_learning_matches_oracle(pkg, laps):
    """The learning loop against the oracle's plain-Python loop in which every agent has its own Oracle (safe set) and
    Recorder.  A closed loop at the limit of the car amplifies differences (1e-10 per tick becomes 1e-2 after 300 ticks,
    measured), so the comparison is made where it is sharp:
      (a) the first 80 ticks (the first wrap, which arms the recorders, included) tick by tick, tight;
      (b) after 300 ticks -- two wraps, every agent's safe set now holds a lap of its own -- ONE further tick from the
          device's state on both sides, the oracle's safe sets rebuilt from the laps the device recorded: tight again.
    Laps, lap counters and queries on the agents' own sets are compared bit for bit in the test above."""
    import oracle_loop as OL
    from oracle import Oracle
    from oracle_regress import Recorder
    from oracle_track import OracleTrack
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    tb = _track_table("barc_center")
    o, veh, cfg, track, _ = make_oracle(pkg, "barc_lmpc", tol=1e-10)
    nb, N, dt, ticks = 3, cfg["N"], 0.025, 300
    m = BatchedRacingMPC(veh, dict(cfg, tol=1e-9), max_batch=nb)
    for l in laps:
        m.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    m.set_track(tb)
    otrk = OracleTrack(tb)
    m.agents_create(nb, 1024)
    x, u_prev, X_last, U_last = _start_near_the_line(pkg, veh, otrk, o, laps, nb, N, dt, 9, back=(20, 60))
    out = m.closed_loop_agents(m.loop_options(dt), ticks, x, u_prev, X_last, U_last, t0=0.0)
    st = m.agents_status()
    assert (st["lap_count"] == 3 + 2).all() and (st["stored"] == 3).all(), st       # two wraps: one own lap stored
    nxt = m.closed_loop_agents(m.loop_options(dt), 1, out["x"], out["u_prev"], out["X_last"], out["U_last"], t0=dt * ticks)
    od = dict(step_mode="step", delay_step=0, plant_substeps=1, dt=dt, plant_dt=dt, speed_limit=1e9, speed_scale=1.0, max_vel_ref_diff=1.0)
    worst_a, worst_b = 0.0, 0.0
    for b in range(nb):
        # (a)
        ob = Oracle(veh, dict(cfg, tol=1e-10))
        for l in laps:
            ob.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
        rec = Recorder(); rec.lap_count = 3
        r = OL.closed_loop(ob, otrk, od, 80, x[b], u_prev[b], X_last[b], U_last[b], impl="port", recorder=rec, t0=0.0)
        assert r["fail_count"] == 0 and rec.initialized
        worst_a = max(worst_a, relerr(out["log_x"][:80, b], r["log_x"]), relerr(out["log_u"][:80, b], r["log_u"]))
        # (b) the oracle's state of tick 300: the device's recorder log replayed, its laps added to a fresh safe set
        ob = Oracle(veh, dict(cfg, tol=1e-10))
        for l in laps:
            ob.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
        rec = Recorder(); rec.lap_count = 3
        for t in range(ticks):
            s_ = out["log_rec"][t, b]
            if rec.step(s_[:6], s_[6:8], s_[8], s_[9], track["length"]):
                l = rec.laps[-1]
                ob.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
        assert len(rec.laps) == 1 and np.array_equal(m.agents_get_lap(b, 0), rec.laps[-1]["x"])
        r1 = OL.closed_loop(ob, otrk, od, 1, out["x"][b], out["u_prev"][b], out["X_last"][b], out["U_last"][b], impl="port", recorder=rec, t0=dt * ticks)
        assert r1["fail_count"] == 0 == nxt["fail_count"][b]
        worst_b = max(worst_b, relerr(nxt["log_x"][:, b], r1["log_x"]), relerr(nxt["log_u"][:, b], r1["log_u"]), relerr(nxt["X_last"][b], r1["X_last"]))
        # the tick really drew on the agent's own lap: the query at its terminal reference returns columns of that lap first
        q = np.array([[out["X_last"][b][-1][0], out["X_last"][b][-1][1]]])
        sx_own = m.agents_query(np.tile(q, (nb, 1)))[b][0]
        assert np.abs(sx_own[:32, None, 1:] - rec.laps[-1]["x"][None, :, 1:]).max(axis=2).min(axis=1).max() == 0.0
    assert worst_a < 1e-7 and worst_b < 1e-8, (worst_a, worst_b)
    print(f"closed loop with per-agent learning, {nb} agents: first 80 ticks vs oracle {worst_a:.2e}; the tick after 300 (own laps in the safe sets) {worst_b:.2e}")
