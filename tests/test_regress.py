"""Error-dynamics regression (safe_set.cpp:56-114,182-245) and the lap recorder (row a8;
safe_set.cpp:278-322): the CUDA path (lane-loop emulated on CPU, through the C ABI on the GPU) against the numpy
restatement in oracle/oracle_regress.py."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import make_oracle, relerr
from test_emulator import emu, _p  # noqa: F401  (fixture)

# the LMPC paper's choice: v_x, v_y, omega regressed on (v_x, v_y, omega) and one input each
OUT = [3, 4, 5]
IN_X = [[3, 4, 5]] * 3
IN_U = [[0], [1], [1]]
H = 0.6


def _points(pkg):
    import oracle_regress as R
    o, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc")
    laps = pkg.workload.load_laps()
    return o, laps, [R.lap_points(o, l) for l in laps]


def _queries(laps, n, seed):
    rng = np.random.default_rng(seed)
    l = laps[-1]
    idx = rng.integers(0, l["x"].shape[0] - 1, n)
    xq = l["x"][idx] + rng.normal(0, [0.05, 0.02, 0.02, 0.05, 0.02, 0.05], (n, 6))
    uq = l["u"][idx] + rng.normal(0, [0.002, 0.02], (n, 2))
    return xq, uq, l["k"][idx]


def _emu_regress(emu, spec, Z, E, xq, uq, A, B, Cv, reverse=0):
    emu.emu_set_reverse(reverse)
    Ac = np.ascontiguousarray(A.T).copy(); Bc = np.ascontiguousarray(B.T).copy(); Cc = Cv.copy()   # column-major
    zq = np.concatenate([xq, uq]); npts = np.zeros(spec.n_out, dtype=np.int32)
    rc = emu.emu_regress(C.byref(spec), Z.shape[0], _p(Z), _p(E), _p(zq), _p(Ac), _p(Bc), _p(Cc), npts.ctypes.data_as(C.c_void_p))
    emu.emu_set_reverse(0)
    assert rc == 0
    return Ac.T.copy(), Bc.T.copy(), Cc, npts


def test_emulated_regression_matches_oracle(emu, pkg):
    import oracle_regress as R
    from racing_lmpc_ros2_b200.binding import make_reg_spec
    o, laps, pts = _points(pkg)
    Z = np.ascontiguousarray(np.vstack([p[0] for p in pts])); E = np.ascontiguousarray(np.vstack([p[1] for p in pts]))
    xq, uq, kq = _queries(laps, 24, 5)
    worst, total = 0.0, 0
    for sign in (1.0, -1.0):
        spec = make_reg_spec(OUT, IN_X, IN_U, H, sign=sign)
        for i in range(len(xq)):
            A0, B0, g0, _ = o.linearise(xq[i], uq[i], kq[i], 0.025)
            A2, B2, C2, n2 = R.regress(pts, OUT, IN_X, IN_U, H, xq[i], uq[i], A0, B0, g0, sign=sign)
            A1, B1, C1, n1 = _emu_regress(emu, spec, Z, E, xq[i], uq[i], A0, B0, g0)
            assert (n1 == n2).all()
            total += int(n1.sum())
            worst = max(worst, np.abs(A1 - A2).max(), np.abs(B1 - B2).max(), np.abs(C1 - C2).max())
            # only the regressed entries move
            mask = np.zeros((6, 6), bool)
            for r, oi in enumerate(OUT):
                mask[oi, IN_X[r]] = True
            assert np.array_equal(A1[~mask], np.asarray(A0)[~mask])
            # bit-identical with the lanes run in reverse order
            A3, B3, C3, _ = _emu_regress(emu, spec, Z, E, xq[i], uq[i], A0, B0, g0, reverse=1)
            assert np.array_equal(A1, A3) and np.array_equal(B1, B3) and np.array_equal(C1, C3)
    assert total > 24 * 3 * 5, total          # the bandwidth actually catches samples
    assert worst < 1e-9, worst


def test_emulated_regression_general_size_class(emu, pkg):
    """Regressions with more than four inputs take the general (9-wide) instantiation: five states + both controls."""
    import oracle_regress as R
    from racing_lmpc_ros2_b200.binding import make_reg_spec
    o, laps, pts = _points(pkg)
    Z = np.ascontiguousarray(np.vstack([p[0] for p in pts])); E = np.ascontiguousarray(np.vstack([p[1] for p in pts]))
    out, in_x, in_u, h = [3, 5], [[1, 2, 3, 4, 5], [3, 4, 5]], [[0, 1], [1]], 0.8
    spec = make_reg_spec(out, in_x, in_u, h)
    xq, uq, kq = _queries(laps, 8, 21)
    worst = 0.0
    for i in range(len(xq)):
        A0, B0, g0, _ = o.linearise(xq[i], uq[i], kq[i], 0.025)
        A2, B2, C2, n2 = R.regress(pts, out, in_x, in_u, h, xq[i], uq[i], A0, B0, g0)
        A1, B1, C1, n1 = _emu_regress(emu, spec, Z, E, xq[i], uq[i], A0, B0, g0)
        assert (n1 == n2).all() and n1.sum() > 0
        worst = max(worst, np.abs(A1 - A2).max(), np.abs(B1 - B2).max(), np.abs(C1 - C2).max())
    assert worst < 1e-8, worst


def test_emulated_regression_sign_and_empty(emu, pkg):
    from racing_lmpc_ros2_b200.binding import make_reg_spec
    o, laps, pts = _points(pkg)
    Z = np.ascontiguousarray(np.vstack([p[0] for p in pts])); E = np.ascontiguousarray(np.vstack([p[1] for p in pts]))
    xq, uq, kq = _queries(laps, 1, 9)
    A0, B0, g0, _ = o.linearise(xq[0], uq[0], kq[0], 0.025)
    Ap, Bp, Cp, _ = _emu_regress(emu, make_reg_spec(OUT, IN_X, IN_U, H, sign=1.0), Z, E, xq[0], uq[0], A0, B0, g0)
    Am, Bm, Cm, _ = _emu_regress(emu, make_reg_spec(OUT, IN_X, IN_U, H, sign=-1.0), Z, E, xq[0], uq[0], A0, B0, g0)
    assert np.allclose(Ap - A0, -(Am - A0), atol=1e-13) and np.allclose(Cp - g0, -(Cm - g0), atol=1e-13)
    assert np.abs(Cp - g0).max() > 0
    # a query far from every sample: nothing within dist_max, the nominal model is returned untouched (:203-205)
    far = xq[0] + np.array([0, 0, 0, 50.0, 0, 0])
    Af, Bf, Cf, nf = _emu_regress(emu, make_reg_spec(OUT, IN_X, IN_U, H), Z, E, far, uq[0], A0, B0, g0)
    assert (nf == 0).all() and np.array_equal(Af, A0) and np.array_equal(Bf, B0) and np.array_equal(Cf, g0)


def test_oracle_recorder_segments_laps(pkg):
    """The recorder restatement on a synthetic 3.5-lap stream: the partial first lap is dropped, complete laps start at
    the wrap sample (safe_set.cpp:290-315)."""
    import oracle_regress as R
    L = 10.0
    s = (0.37 * np.arange(100) + 4.0)
    rec = R.Recorder(); added = []
    for j, sj in enumerate(s):
        x = np.array([sj % L, 0, 0, 1, 0, 0.0])
        added.append(rec.step(x, np.array([0.1, 0.0]), 0.0, 0.1 * j, L))
    wraps = [j for j in range(1, 100) if (s[j - 1] % L) - (s[j] % L) > 0.5 * L]
    assert len(rec.laps) == len(wraps) - 1 and sum(added) == len(wraps) - 1
    for q, lap in enumerate(rec.laps):
        assert lap["x"].shape[0] == wraps[q + 1] - wraps[q]
        assert np.isclose(lap["t"][0], 0.1 * wraps[q])
        assert lap["x"].shape[0] == lap["u"].shape[0] == lap["k"].shape[0] == lap["t"].shape[0]


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_regression_matches_oracle(pkg):
    import oracle_regress as R
    from racing_lmpc_ros2_b200.binding import make_reg_spec
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    o, laps, pts = _points(pkg)
    veh = pkg.configs.BARC_VEHICLE; cfg = pkg.configs.barc_lmpc_config(20)
    track = pkg.workload.load_track("barc_center")
    m = BatchedRacingMPC(veh, cfg, max_batch=8)
    for l in laps:
        m.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    n = 96
    xq, uq, kq = _queries(laps, n, 11)
    A0 = np.zeros((n, 6, 6)); B0 = np.zeros((n, 6, 2)); C0 = np.zeros((n, 6))
    for i in range(n):
        A0[i], B0[i], C0[i], _ = o.linearise(xq[i], uq[i], kq[i], 0.025)
    spec = make_reg_spec(OUT, IN_X, IN_U, H)
    n0 = m.launch_count
    A1, B1, C1, n1 = m.regress(spec, xq, uq, A0, B0, C0)
    assert m.launch_count - n0 == 2     # the one-off preparation of the points + the regression kernel
    worst = 0.0
    for i in range(n):
        A2, B2, C2, n2 = R.regress(pts, OUT, IN_X, IN_U, H, xq[i], uq[i], A0[i], B0[i], C0[i])
        assert (n1[i] == n2).all()
        worst = max(worst, np.abs(A1[i] - A2).max(), np.abs(B1[i] - B2).max(), np.abs(C1[i] - C2).max())
    assert n1.sum() > n * 3 * 5
    assert worst < 1e-8, worst
    n0 = m.launch_count
    m.regress(spec, xq, uq, A0, B0, C0)
    assert m.launch_count - n0 == 1     # points are prepared once per safe-set update
    # other shapes of the plan: a pair and singles of different sizes (the exact-size scans for 1, 3 and 4 regressors),
    # and the general size class (six states and both controls: the index-list scan)
    for out_s, inx_s, inu_s in (([3, 4, 5, 2], [[3, 4], [3, 4], [5], [5, 3, 4]], [[0], [0], [], [1]]),
                                ([5, 3], [[0, 1, 2, 3, 4, 5], [3, 4, 5]], [[0, 1], [1]])):
        sp = make_reg_spec(out_s, inx_s, inu_s, H)
        A3, B3, C3, n3 = m.regress(sp, xq[:32], uq[:32], A0[:32], B0[:32], C0[:32])
        w2 = 0.0
        for i in range(32):
            A2, B2, C2, n2 = R.regress(pts, out_s, inx_s, inu_s, H, xq[i], uq[i], A0[i], B0[i], C0[i])
            assert (n3[i] == n2).all(), (out_s, i, n3[i], n2)
            w2 = max(w2, np.abs(A3[i] - A2).max(), np.abs(B3[i] - B2).max(), np.abs(C3[i] - C2).max())
        assert n3.sum() > 0 and w2 < 1e-8, (out_s, w2)


@pytest.mark.gpu
def test_gpu_tick_with_error_dynamics(pkg):
    """lmpc_set_error_dynamics: the tick's QP is posed on the corrected (A, B, g) -- the returned trajectory satisfies
    x_{i+1} = A'_i x_i + B'_i u_i + g'_i with the oracle's corrected matrices, and differs from the nominal tick."""
    import oracle_regress as R
    from racing_lmpc_ros2_b200.binding import make_reg_spec
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    o, laps, pts = _points(pkg)
    veh = pkg.configs.BARC_VEHICLE; cfg = pkg.configs.barc_lmpc_config(20)
    track = pkg.workload.load_track("barc_center")
    m = BatchedRacingMPC(veh, cfg, max_batch=16)
    for l in laps:
        m.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    batch = pkg.workload.make_batch(veh, cfg, 16, 0x8F3, track, laps, mode="barc")
    nominal = m.solve(batch)
    m.set_error_dynamics(make_reg_spec(OUT, IN_X, IN_U, H))
    n0 = m.launch_count
    out = m.solve(batch)
    assert m.launch_count - n0 == 5      # prepare (once), linearise, regression, safe-set query, QP
    assert (out["status"] == 0).all()
    N = cfg["N"]; worst = 0.0
    for b in range(16):
        inp = pkg.workload.instance(batch, b)
        for i in range(N - 1):
            xr = inp["X_ref"][i].copy(); xr[0] = o.align_abscissa(xr[0], inp["x_ic"][0], inp["total_length"])
            A0, B0, g0, _ = o.linearise(xr, inp["U_ref"][i], inp["curvatures"][i], inp["T_ref"][i])
            A2, B2, C2, _ = R.regress(pts, OUT, IN_X, IN_U, H, xr, inp["U_ref"][i], A0, B0, g0)
            pred = A2 @ out["X_optm"][b, i] + B2 @ out["U_optm"][b, i] + C2
            worst = max(worst, np.abs(pred - out["X_optm"][b, i + 1]).max())
    assert worst < 1e-8, worst
    assert np.abs(out["X_optm"] - nominal["X_optm"]).max() > 1e-6
    m.set_error_dynamics(None)
    again = m.solve(batch)
    assert np.array_equal(again["X_optm"], nominal["X_optm"])


@pytest.mark.gpu
def test_gpu_recorder_matches_oracle(pkg, tmp_path):
    import oracle_regress as R
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    veh = pkg.configs.BARC_VEHICLE; cfg = pkg.configs.barc_lmpc_config(20)
    m = BatchedRacingMPC(veh, cfg, max_batch=4)
    L = 10.0
    prefix = str(tmp_path) + os.sep
    m.recorder_config(True, prefix)
    rng = np.random.default_rng(3)
    rec = R.Recorder()
    s = 4.0
    for j in range(160):
        s += rng.uniform(0.2, 0.5)
        x = np.array([s % L, *rng.normal(0, 0.1, 5)]); u = rng.normal(0, 0.1, 2); k = rng.normal(); t = 0.1 * j
        a1 = m.recorder_step(x, u, k, t, L); a2 = rec.step(x, u, k, t, L)
        assert a1 == a2
    assert m.num_laps() == min(len(rec.laps), cfg["max_lap_stored"]) and len(rec.laps) >= 3
    assert m.recorder_lap_count() == rec.lap_count
    # files: "%.16e" text, read back bit-exactly by lmpc_safe_set_load (SafeSetRecorder::load)
    for q, lap in enumerate(rec.laps):
        fx = np.loadtxt(prefix + f"lap_{q + 1}_x.txt").reshape(-1, 6)
        assert np.array_equal(fx, lap["x"])
        assert np.array_equal(np.loadtxt(prefix + f"lap_{q + 1}_u.txt").reshape(-1, 2), lap["u"])
        assert np.array_equal(np.loadtxt(prefix + f"lap_{q + 1}_t.txt").ravel(), lap["t"])
        assert np.array_equal(np.loadtxt(prefix + f"lap_{q + 1}_k.txt").ravel(), lap["k"])
    m2 = BatchedRacingMPC(veh, cfg, max_batch=4)
    m2.load_lap(prefix + "lap_1", L)
    sx1, sj1 = m2.ss_query(np.array([[3.0, 0.0]]), 8, 8)
    o, *_ = make_oracle(pkg, "barc_lmpc", with_laps=False)
    o.add_lap(rec.laps[0]["x"], rec.laps[0]["u"], rec.laps[0]["k"], rec.laps[0]["t"], L)
    sx2, sj2 = o.ss_query(3.0, 0.0, 8, 8)
    assert np.array_equal(sx1[0], sx2) and np.array_equal(sj1[0], sj2)


@pytest.mark.gpu
def test_gpu_config4_batch_survives_instances_that_blow_up(pkg):
    """BASELINE config 4 as bench.py's `configs` line builds it (50 synthetic laps, the regression on every stage, cap 60,
    the line's seed).  A few of these QPs are infeasible in all but name and one of them drives its
    iterate to NaN inside the iterations; the arg-max over its NaN safe-set weights then has no winner, and its "none" index
    once became a shared-memory/scratch address (illegal memory access, found by the bench run at the end of round 2).
    The batch must come back: the failures flagged, everything else solved."""
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    from racing_lmpc_ros2_b200.binding import make_reg_spec
    laps = pkg.workload.load_laps(); tr = pkg.workload.load_track("barc_center"); veh = pkg.configs.BARC_VEHICLE
    cfg = dict(pkg.configs.barc_lmpc_config(20), num_ss_pts_per_lap=2, max_lap_stored=50, max_iter=60)
    m = BatchedRacingMPC(veh, cfg, max_batch=2048)
    for l in pkg.workload.synthesise_laps(laps, 50):
        m.add_lap(l["x"], l["u"], l["k"], l["t"], tr["length"])
    m.set_error_dynamics(make_reg_spec([3, 4, 5], [[3, 4, 5]] * 3, [[0], [1], [1]], 0.6))
    out = m.solve(pkg.workload.make_batch(veh, cfg, 2048, 0xB200 + 40, tr, laps))
    st = out["status"]
    assert set(np.unique(st)) <= {0, 1, 4, 5}, np.bincount(st)
    assert (st == 0).mean() > 0.9 and (st == 4).sum() >= 1, np.bincount(st)
    assert np.isfinite(out["X_optm"][st == 0]).all()
    m.close()


def test_regression_plan_pairs_identical_input_lists(emu, pkg):
    """Regressions with the same input lists share regressors and weights, hence M'KM: the plan marks them as a pair and
    the tiled kernel scans once for both (lmpc_reg_core.cuh, LmpcRegRow::lead / follower).  Pairs only: a third
    regression with the same lists leads its own scan; different lists never pair."""
    import ctypes as C
    from racing_lmpc_ros2_b200.binding import make_reg_spec

    def pairs(out, in_x, in_u):
        spec = make_reg_spec(out, in_x, in_u, 0.6)
        lead = (C.c_int * 6)(); fol = (C.c_int * 6)()
        n = emu.emu_reg_plan_pairs(C.byref(spec), lead, fol)
        assert n == len(out)
        return list(lead[:n]), list(fol[:n])

    assert pairs([3, 4, 5], [[3, 4, 5]] * 3, [[0], [1], [1]]) == ([0, 1, 1], [-1, 2, -1])      # the LMPC choice: v_y and omega pair
    assert pairs([3, 4, 5], [[3, 4, 5]] * 3, [[1], [1], [1]]) == ([0, 0, 2], [1, -1, -1])      # three alike: one pair, one alone
    assert pairs([3, 4], [[3, 4, 5], [3, 5, 4]], [[1], [1]]) == ([0, 1], [-1, -1])             # same set, other order: not a pair
    assert pairs([3, 4, 5, 2], [[3, 4], [3, 4], [5], [5]], [[0], [0], [], []]) == ([0, 0, 2, 2], [1, -1, 3, -1])
