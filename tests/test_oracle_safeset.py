"""Pins the oracle's safe-set restatement against a plain numpy brute force and the reference's
documented semantics (newest lap first, 32 per lap, nearest first, tripled laps, J bookkeeping)."""
import numpy as np

from conftest import make_oracle


def _brute(laps, L, qs, qe, max_total, per_lap):
    xs, js = [], []
    total = 0
    for lap in reversed(laps):
        if total >= max_total:
            break
        x = lap["x"]; n = x.shape[0]
        off = np.zeros_like(x); off[:, 0] = L
        xr = np.vstack([x - off, x, x + off])
        J = np.linspace(n - 1, 0, n)
        Jr = np.concatenate([J + n - 1, J, J - n + 1])
        d2 = (qs - xr[:, 0]) ** 2 + (qe - xr[:, 1]) ** 2
        idx = np.lexsort((np.arange(3 * n), d2))[:per_lap]
        xs.append(xr[idx]); js.append(Jr[idx]); total += len(idx)
    return np.vstack(xs)[:max_total], np.concatenate(js)[:max_total]


def test_query_matches_numpy_bruteforce(pkg, laps, barc_track):
    o, *_ = make_oracle(pkg, "barc_lmpc")
    L = barc_track["length"]
    rng = np.random.default_rng(11)
    for _ in range(40):
        qs, qe = rng.uniform(-3, 20), rng.uniform(-0.4, 0.4)
        sx, sj = o.ss_query(qs, qe)
        bx, bj = _brute(laps, L, qs, qe, 96, 32)
        assert sx.shape == (96, 6)
        assert np.array_equal(sx, bx) and np.array_equal(sj, bj)
        # per lap block: ascending distance
        for blk in range(3):
            d2 = (qs - sx[32 * blk:32 * blk + 32, 0]) ** 2 + (qe - sx[32 * blk:32 * blk + 32, 1]) ** 2
            assert (np.diff(d2) >= 0).all()


def test_truncation_padding_and_cost_shift(pkg, laps, barc_track):
    from oracle import Oracle
    veh = pkg.configs.BARC_VEHICLE
    cfg = pkg.configs.barc_lmpc_config(20)
    L = barc_track["length"]
    # one lap only: 32 found, padded to 96 by repeating the last column (racing_mpc.cpp:263-272)
    o = Oracle(veh, cfg)
    o.add_lap(laps[0]["x"], laps[0]["u"], laps[0]["k"], laps[0]["t"], L)
    sx, sc, cnt = o.ss_query_padded(3.0, 0.0)
    assert cnt == 32
    assert np.array_equal(sx[32:], np.repeat(sx[31:32], 64, axis=0)) and np.all(sc[32:] == sc[31])
    assert sc[0] == 0.0                                       # J - J[0] (racing_mpc.cpp:280)
    # circular buffer keeps the newest max_lap_stored laps (safe_set.cpp:139-151)
    o = Oracle(veh, dict(cfg, max_lap_stored=2))
    for l in laps:
        o.add_lap(l["x"], l["u"], l["k"], l["t"], L)
    assert o.num_laps() == 2
    sx, sj = o.ss_query(3.0, 0.0, max_total=96, per_lap=32)
    assert sx.shape[0] == 64
    # truncation: per-lap 40 -> 40 + 40 + 40 = 120 > 96 -> first 96 (safe_set.cpp:175-178)
    o = Oracle(veh, cfg)
    for l in laps:
        o.add_lap(l["x"], l["u"], l["k"], l["t"], L)
    sx, sj = o.ss_query(3.0, 0.0, max_total=96, per_lap=40)
    assert sx.shape[0] == 96


def test_cost_to_go_bookkeeping(pkg, laps, barc_track):
    """J = steps to the finish line; the s-L copy is one lap behind (J + n - 1), s+L one ahead."""
    o, *_ = make_oracle(pkg, "barc_lmpc")
    L = barc_track["length"]
    n3 = laps[2]["x"].shape[0]
    x0 = laps[2]["x"][10]
    sx, sj = o.ss_query(x0[0], x0[1], max_total=1, per_lap=1)
    assert np.array_equal(sx[0], x0) and sj[0] == n3 - 1 - 10
    sx, sj = o.ss_query(x0[0] - L, x0[1], max_total=1, per_lap=1)
    assert sj[0] == (n3 - 1 - 10) + (n3 - 1)
    sx, sj = o.ss_query(x0[0] + L, x0[1], max_total=1, per_lap=1)
    assert sj[0] == (n3 - 1 - 10) - (n3 - 1)


def test_duplicate_keys_resolve_to_first_index(pkg):
    from oracle import Oracle
    veh = pkg.configs.BARC_VEHICLE
    cfg = pkg.configs.barc_lmpc_config(20)
    x = np.zeros((5, 6)); x[:, 0] = [0.0, 1.0, 1.0, 2.0, 3.0]; x[:, 3] = [10, 11, 12, 13, 14]
    o = Oracle(veh, cfg)
    o.add_lap(x, np.zeros((5, 2)), np.zeros(5), np.arange(5.0), 100.0)
    sx, sj = o.ss_query(1.0, 0.0, max_total=2, per_lap=2)
    # both nearest points share the key (1, 0): the coordinate hash returns the first inserted index twice
    assert np.all(sx[:, 3] == 11) and np.all(sj == 3)


def test_fifty_lap_safe_set_config4(pkg, laps, barc_track):
    """BASELINE configs[3]: 50 laps stored; 32 per lap searches the newest 3 only (safe_set.cpp:164), 2 per lap
    draws from 48 laps."""
    from oracle import Oracle
    veh = pkg.configs.BARC_VEHICLE
    L = barc_track["length"]
    many = pkg.workload.synthesise_laps(laps, 50)
    rng = np.random.default_rng(3)
    for per_lap in (32, 2):
        cfg = dict(pkg.configs.barc_lmpc_config(20), max_lap_stored=50, num_ss_pts_per_lap=per_lap)
        o = Oracle(veh, cfg)
        for l in many:
            o.add_lap(l["x"], l["u"], l["k"], l["t"], L)
        for _ in range(10):
            qs, qe = rng.uniform(-3, 20), rng.uniform(-0.4, 0.4)
            sx, sj = o.ss_query(qs, qe)
            bx, bj = _brute(many, L, qs, qe, 96, per_lap)
            assert np.array_equal(sx, bx) and np.array_equal(sj, bj)
