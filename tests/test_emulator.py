"""Lane-loop emulation of the CUDA kernels' source (tests/emu, -DLMPC_EMULATE build of the same .cuh
files) against the CPU oracle.  This is how the warp kernels' logic is checked on a box without a GPU;
the emulator is a test artefact and is never loaded by the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import CASES, ROOT, make_oracle, relerr

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "liblmpc_emu.so")
DP = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(DP)


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU_DIR, "emu_core.cpp")] + [
        os.path.join(ROOT, "racing-lmpc-ros2_b200", "csrc", f)
        for f in ("lmpc_qp_core.cuh", "lmpc_ss_core.cuh", "lmpc_model.cuh", "lmpc_warp.cuh", "lmpc_host_params.h")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", EMU_SO, srcs[0]])
    return C.CDLL(EMU_SO)


def _emu_solve(emu, pkg, od, veh, cfg, inp, reverse=0, nw=1):
    from racing_lmpc_ros2_b200 import binding as Bd
    emu.emu_set_reverse(reverse)
    vs = Bd.fill_struct(Bd.VehicleParams(), veh)
    cs = Bd.fill_struct(Bd.MpcConfig(), cfg)
    N, K = cfg["N"], cfg["num_ss_pts"]
    Xr = inp["X_ref"].copy()
    for j in range(N):
        Xr[j, 0] = od.align_abscissa(Xr[j, 0], inp["x_ic"][0], inp["total_length"])
    ABg = np.zeros((N - 1, 54))
    A1 = np.zeros(36); B1 = np.zeros(12); g1 = np.zeros(6); xn = np.zeros(6)
    for j in range(N - 1):   # the product's own (analytic) linearisation, emulated
        u = np.ascontiguousarray(inp["U_ref"][j])
        emu.emu_linearise(C.byref(vs), _p(np.ascontiguousarray(Xr[j])), _p(u), C.c_double(inp["curvatures"][j]),
                          C.c_double(inp["T_ref"][j]), _p(A1), _p(B1), _p(g1), _p(xn))
        ABg[j, :36] = A1; ABg[j, 36:48] = B1; ABg[j, 48:] = g1
    if cfg["learning"]:
        sx, sj = od.ss_query(Xr[N - 1, 0], Xr[N - 1, 1])
        cnt = len(sj)
        ssx = np.zeros((K, 6)); ssj = np.zeros(K)
        ssx[:cnt] = sx; ssj[:cnt] = sj; ssx[cnt:] = sx[-1]; ssj[cnt:] = sj[-1]
    else:
        ssx = np.zeros((1, 6)); ssj = np.zeros(1); cnt = 0
    cen = np.ascontiguousarray(Xr[N - 1])
    X = np.zeros((N, 6)); U = np.zeros((N - 1, 2)); dU = np.zeros((N - 1, 2)); lam = np.zeros(max(K, 1))
    cost = C.c_double(); st = C.c_int(); it = C.c_int(); smd = C.c_int()
    rc = emu.emu_qp_solve(C.byref(cs), C.byref(vs), _p(inp["x_ic"]), _p(inp["u_ic"]), _p(np.ascontiguousarray(inp["U_ref"])),
                          _p(inp["T_ref"]), _p(inp["bound_left"]), _p(inp["bound_right"]), _p(inp["vel_ref"]), _p(ABg),
                          _p(ssx), _p(ssj), _p(cen), cnt, _p(X), _p(U), _p(dU), _p(lam), C.byref(cost), C.byref(st),
                          C.byref(it), C.byref(smd), int(nw))
    assert rc == 0
    return dict(X=X, U=U, dU=dU, lam=lam, cost=cost.value, status=st.value, iters=it.value, smem=smd.value * 8, ss_x=ssx)


@pytest.mark.parametrize("nw", [1, 2, 4])
@pytest.mark.parametrize("name", list(CASES))
def test_emulated_qp_kernel_matches_dense_oracle(emu, pkg, name, nw):
    od, veh, cfg, track, mode = make_oracle(pkg, name, tol=1e-11)
    cfgk = dict(cfg, tol=1e-13)
    batch = pkg.workload.make_batch(veh, cfg, 10, 0xE31 + len(name), track, pkg.workload.load_laps(), mode=mode)
    worst = 0.0
    for b in range(10):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        if not (d["status"] == 0 and d["polished"] == 1 and d["kkt"] < 1e-9):
            continue
        k = _emu_solve(emu, pkg, od, veh, cfgk, inp, nw=nw)
        assert k["status"] == 0
        worst = max(worst, relerr(k["X"], d["X"]), relerr(k["U"], d["U"]), relerr(k["dU"], d["dU"]))
        assert abs(k["cost"] - d["cost"]) < 1e-8 * max(1, abs(d["cost"]))
        if cfg["learning"]:
            assert np.allclose(k["ss_x"].T @ k["lam"], d["ss_x"].T @ d["lam"], atol=1e-7)
            assert abs(k["lam"].sum() - 1) < 1e-9 and k["lam"].min() >= 0
    assert worst < 1e-6, worst


@pytest.mark.parametrize("nw", [1, 2, 4])
def test_emulated_qp_kernel_is_lane_order_independent(emu, pkg, nw):
    """Running the lanes of every phase in reverse order must give bit-identical results: a phase that
    read shared memory written by another lane in the same phase would differ."""
    od, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc", tol=1e-11)
    cfgk = dict(cfg, tol=1e-13)
    batch = pkg.workload.make_batch(veh, cfg, 3, 0xAB, track, pkg.workload.load_laps(), mode=mode)
    for b in range(3):
        inp = pkg.workload.instance(batch, b)
        f = _emu_solve(emu, pkg, od, veh, cfgk, inp, reverse=0, nw=nw)
        r = _emu_solve(emu, pkg, od, veh, cfgk, inp, reverse=1, nw=nw)
        for key in ("X", "U", "dU", "lam"):
            assert np.array_equal(f[key], r[key]), key
        assert f["iters"] == r["iters"] and f["cost"] == r["cost"]


def test_qp_shared_memory_budget(emu, pkg):
    """BASELINE config 2 (N=20, K=96) must fit 7 instances per SM: <= (228 KB - 7 KB reserved) / 7 with one or two
    warps per instance (one wave of 1024 instances on 148 SMs); the 4-warp variant (cross-warp reduction scratch) fits 6."""
    od, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc")
    batch = pkg.workload.make_batch(veh, cfg, 1, 1, track, pkg.workload.load_laps(), mode=mode)
    for nw in (1, 2, 4):
        k = _emu_solve(emu, pkg, od, veh, cfg, pkg.workload.instance(batch, 0), nw=nw)
        per_sm = 7 if nw <= 2 else 6
        assert k["smem"] <= (228 * 1024 - per_sm * 1024) // per_sm, (nw, k["smem"])


def test_emulated_ss_query_matches_oracle(emu, pkg, laps, barc_track):
    od, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc")
    L = barc_track["length"]
    rng = np.random.default_rng(2)
    # build the slab exactly as the C-ABI's add_lap does (tripled points, J, canon)
    slabs = []
    for lap in reversed(laps):
        x = lap["x"]; n = x.shape[0]
        off = np.zeros_like(x); off[:, 0] = L
        xr = np.ascontiguousarray(np.vstack([x - off, x, x + off]))
        J = np.linspace(n - 1, 0, n); Jr = np.ascontiguousarray(np.concatenate([J + n - 1, J, J - n + 1]))
        canon = np.arange(3 * n, dtype=np.int32)
        slabs.append((np.ascontiguousarray(xr[:, 0]), np.ascontiguousarray(xr[:, 1]), xr, Jr, canon))
    for rev in (0, 1):
        emu.emu_set_reverse(rev)
        for _ in range(15):
            qs, qe = rng.uniform(-3, 20), rng.uniform(-.4, .4)
            sx = np.zeros((96, 6)); sj = np.zeros(96)
            for j, (ps, pe, xr, Jr, canon) in enumerate(slabs):
                emu.emu_ss_query_lap(len(ps), _p(ps), _p(pe), _p(xr), _p(Jr), canon.ctypes.data_as(C.POINTER(C.c_int)), 32,
                                     32 * j, C.c_double(qs), C.c_double(qe), 96, _p(sx), _p(sj), int(j == 2), 96, 96)
            ox, oj = od.ss_query(qs, qe)
            assert np.array_equal(sx, ox) and np.array_equal(sj, oj)


def test_emulated_ss_query_exact_fallback(emu, pkg):
    """Clustered data: most of the nearest points fall on few lanes, forcing the exact continuation."""
    from oracle import Oracle
    veh = pkg.configs.BARC_VEHICLE
    cfg = pkg.configs.barc_lmpc_config(20)
    n = 400
    x = np.zeros((n, 6)); x[:, 0] = np.linspace(0, 40, n)
    # every 32nd point (same lane) is pulled next to the query
    x[::32, 0] = 5.0 + 1e-3 * np.arange(len(x[::32])); x[::32, 1] = 0.01
    L = 1000.0
    o = Oracle(veh, cfg)
    o.add_lap(x, np.zeros((n, 2)), np.zeros(n), np.arange(float(n)), L)
    off = np.zeros_like(x); off[:, 0] = L
    xr = np.ascontiguousarray(np.vstack([x - off, x, x + off]))
    J = np.linspace(n - 1, 0, n); Jr = np.ascontiguousarray(np.concatenate([J + n - 1, J, J - n + 1]))
    canon = np.arange(3 * n, dtype=np.int32)
    ps = np.ascontiguousarray(xr[:, 0]); pe = np.ascontiguousarray(xr[:, 1])
    sx = np.zeros((32, 6)); sj = np.zeros(32)
    emu.emu_set_reverse(0)
    emu.emu_ss_query_lap(3 * n, _p(ps), _p(pe), _p(xr), _p(Jr), canon.ctypes.data_as(C.POINTER(C.c_int)), 32, 0,
                         C.c_double(5.0), C.c_double(0.0), 32, _p(sx), _p(sj), 1, 32, 32)
    ox, oj = o.ss_query(5.0, 0.0, max_total=32, per_lap=32)
    assert np.array_equal(sx, ox) and np.array_equal(sj, oj)


def test_emulated_qp_kernel_fifty_lap_variant(emu, pkg, laps, barc_track):
    """configs[3] with 2 points per lap: the 96 columns come from 48 different laps (spread-out hull)."""
    from oracle import Oracle
    veh = pkg.configs.BARC_VEHICLE
    many = pkg.workload.synthesise_laps(laps, 50)
    cfg = dict(pkg.configs.barc_lmpc_config(20), max_lap_stored=50, num_ss_pts_per_lap=2)
    od = Oracle(veh, dict(cfg, tol=1e-11))
    for l in many:
        od.add_lap(l["x"], l["u"], l["k"], l["t"], barc_track["length"])
    batch = pkg.workload.make_batch(veh, cfg, 6, 0xB200 + 4, barc_track, many, mode="barc")
    worst = 0.0; n = 0
    for b in range(6):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        k = _emu_solve(emu, pkg, od, veh, cfg, inp)
        if d["status"] == 0 and d["kkt"] < 1e-10 and k["status"] == 0:
            worst = max(worst, relerr(k["X"], d["X"]), relerr(k["U"], d["U"]), relerr(k["dU"], d["dU"])); n += 1
    assert n >= 4 and worst < 1e-6, (n, worst)


@pytest.mark.parametrize("name", ["barc_lmpc", "barc_tracking"])
def test_emulated_qp_kernel_all_state_boxes(emu, pkg, name):
    """Every state boxed (12 + 10 = 22 rows per stage): the lane-per-row passes take their one-half, 32-slot mapping and
    the run-time shared-memory layout; loose e_y / e_psi / s boxes so that the optimum has both active and idle boxes."""
    from oracle import Oracle
    od, veh, cfg, track, mode = make_oracle(pkg, name, tol=1e-11)
    cfg = dict(cfg, x_max=[1e3, 0.5, 0.6, cfg["x_max"][3], 1.0, 3.0], x_min=[-1e3, -0.5, -0.6, 0.1, -1.0, -3.0], tol=1e-11)
    od = Oracle(veh, cfg)
    for l in pkg.workload.load_laps():
        od.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    batch = pkg.workload.make_batch(veh, cfg, 6, 0xA11, track, pkg.workload.load_laps(), mode=mode)
    worst, n = 0.0, 0
    for b in range(6):
        inp = pkg.workload.instance(batch, b)
        dref = od.step(inp, impl="dense")
        if not (dref["status"] == 0 and dref["polished"] == 1 and dref["kkt"] < 1e-9):
            continue
        k = _emu_solve(emu, pkg, od, veh, dict(cfg, tol=1e-13), inp)
        assert k["status"] == 0
        worst = max(worst, relerr(k["X"], dref["X"]), relerr(k["U"], dref["U"]), relerr(k["dU"], dref["dU"]))
        n += 1
    assert n >= 3 and worst < 1e-6, (n, worst)


@pytest.mark.parametrize("hs", [[0.0] * 6, [20.0, 20.0, 0.0, 20.0, 0.0, 2.0]], ids=["hard_hull", "partly_free_slack"])
def test_emulated_qp_kernel_hull_slack_variants(emu, pkg, hs):
    """`convex_hull_slack` all zero -> the hull is a hard equality x_N = SS lambda (racing_mpc.cpp:493,502-503); a zero
    weight on some components only -> those slack components are free and their hull rows vacuous.  Infeasible hard-hull
    instances must be reported by kernel and oracle alike."""
    from oracle import Oracle
    od, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc", tol=1e-11)
    cfg = dict(cfg, convex_hull_slack=hs, tol=1e-11)
    od = Oracle(veh, cfg)
    for l in pkg.workload.load_laps():
        od.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    batch = pkg.workload.make_batch(veh, cfg, 8, 0xC0, track, pkg.workload.load_laps(), mode=mode)
    worst, n, nfail = 0.0, 0, 0
    for b in range(8):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        k = _emu_solve(emu, pkg, od, veh, dict(cfg, tol=1e-13), inp)
        if d["status"] != 0:
            assert k["status"] != 0
            nfail += 1
            continue
        if not (d["polished"] == 1 and d["kkt"] < 1e-9):
            continue
        assert k["status"] == 0
        worst = max(worst, relerr(k["X"], d["X"]), relerr(k["U"], d["U"]), relerr(k["dU"], d["dU"]))
        assert abs(k["cost"] - d["cost"]) < 1e-8 * max(1, abs(d["cost"]))
        if not any(hs):
            assert np.abs(k["X"][-1] - k["ss_x"].T @ k["lam"]).max() < 1e-9      # terminal state inside the hull
        n += 1
    assert n >= 6 and worst < 1e-6, (n, worst)
    if not any(hs):
        assert nfail >= 1        # seed 0xC0 holds one initial state that cannot reach the hull in N steps


@pytest.mark.parametrize("name", ["barc_lmpc", "barc_tracking"])
def test_emulated_qp_kernel_euler_integrator(emu, pkg, name):
    """`integrator_type: euler` (single_track_planar_model.cpp:362-366; no shipped parameter file selects it).  Explicit
    Euler on the BARC's lateral dynamics is unstable at dt = 0.025: the linear rollout from x_ic grows to 1e5 over 19
    stages, which used to cost 30 iterations and set the channel scales of the stopping test (errors up to 5e-4 with
    status 0).  With the guarded start the kernel needs the usual 10-14 iterations; what is left is the conditioning of
    the problem itself (rounding x growth), which the polish cannot always certify: the LMPC case lands below 1e-6
    throughout, the tracking case on at least 6 of 8 instances and never worse than 1e-3."""
    from conftest import make_case
    from oracle import Oracle
    veh, cfg, track, mode = make_case(pkg, name, None, None)      # the kernel runs at its default tolerance
    veh = dict(veh, integrator=1)
    od = Oracle(veh, dict(cfg, tol=1e-11))
    if cfg["learning"]:
        for l in pkg.workload.load_laps():
            od.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    batch = pkg.workload.make_batch(veh, cfg, 8, 0xEE, track, pkg.workload.load_laps(), mode=mode)
    errs, its, sts = [], [], []
    for b in range(8):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        if not (d["status"] == 0 and d["polished"] == 1 and d["kkt"] < 1e-9):
            continue
        k = _emu_solve(emu, pkg, od, veh, cfg, inp)
        assert k["status"] in (0, 5)      # 5 = SOLVED_INACCURATE: interior-point answer, polish not certified
        errs.append(max(relerr(k["X"], d["X"]), relerr(k["U"], d["U"]), relerr(k["dU"], d["dU"])))
        its.append(k["iters"]); sts.append(k["status"])
    errs = np.array(errs); sts = np.array(sts)
    assert len(errs) >= 7 and max(its) <= 20, (len(errs), its)
    assert errs.max() < 1e-3 and (errs < 1e-6).sum() >= (len(errs) if name == "barc_lmpc" else 6), errs
    # status SOLVED means the certified optimum: 1e-6 on the LMPC case; on the tracking case, whose unstable Euler rollout
    # multiplies every last-place difference of the linearisation by 1e5, one instance sits at 2e-6
    assert (errs[sts == 0] < (1e-6 if name == "barc_lmpc" else 1e-5)).all(), (errs, sts)


@pytest.mark.parametrize("name,nb", [("hawaii_kart_tracking", 8), ("iac_lmpc", 4)])
def test_emulated_qp_kernel_other_shipped_parameter_sets(emu, pkg, name, nb):
    """The go-kart tracking set (rear-only braking, control weights 1e-12, N = 10, margin 0) and the IAC LMPC set (N = 60,
    K = 96, its own hull-slack weights) against the dense oracle; see conftest.make_extra_case."""
    from conftest import make_extra_case
    from oracle import Oracle
    veh, cfg, track, dt, laps = make_extra_case(pkg, name)
    od = Oracle(veh, dict(cfg, tol=1e-11))
    for l in laps or []:
        od.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    batch = pkg.workload.make_batch(veh, cfg, nb, 0xB200 + 7, track, laps, dt=dt, mode="track")
    worst, n = 0.0, 0
    for b in range(nb):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        if not (d["status"] == 0 and d["polished"] == 1 and d["kkt"] < 1e-9):
            continue
        k = _emu_solve(emu, pkg, od, veh, cfg, inp)
        assert k["status"] == 0
        worst = max(worst, relerr(k["X"], d["X"]), relerr(k["U"], d["U"]), relerr(k["dU"], d["dU"]))
        assert abs(k["cost"] - d["cost"]) < 1e-7 * max(1, abs(d["cost"]))
        n += 1
    assert n >= nb - 1 and worst < 1e-6, (n, worst)


@pytest.mark.parametrize("name", ["hawaii_kart_tracking", "iac_lmpc"])
def test_extra_parameter_sets_reproduce_golden_vectors(emu, pkg, name):
    """Committed golden vectors of the two extra parameter sets (tests/golden/make_golden.py): the dense oracle reproduces
    them, the port and the emulated kernel land on them within the bar.  iac_lmpc instance 3 (reference end point far
    outside the hull, rollout 69 m off the track, cost 3.5e3) is the case the boundary-slack start was made for: 87
    structured iterations before it, 37 with it (the dense oracle needs 44)."""
    from conftest import make_extra_case
    from oracle import Oracle
    z = np.load(os.path.join(ROOT, "tests", "golden", f"golden_{name}.npz"))
    batch = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    veh, cfg, track, dt, laps = make_extra_case(pkg, name)
    od = Oracle(veh, dict(cfg, tol=1e-11)); op = Oracle(veh, cfg)
    for l in laps or []:
        od.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"]); op.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    for b in range(batch["x_ic"].shape[0]):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        assert d["status"] == 0 and d["kkt"] < 1e-9
        assert relerr(d["X"], z["out_X"][b]) < 1e-9 and relerr(d["U"], z["out_U"][b]) < 1e-9
        p = op.step(inp, impl="port")
        k = _emu_solve(emu, pkg, od, veh, cfg, inp)
        for r, X, U, dU in ((p, p["X"], p["U"], p["dU"]), (k, k["X"], k["U"], k["dU"])):
            assert r["status"] == 0
            assert max(relerr(X, z["out_X"][b]), relerr(U, z["out_U"][b]), relerr(dU, z["out_dU"][b])) < 1e-6
            assert abs(r["cost"] - z["out_cost"][b]) < 1e-7 * max(1, abs(z["out_cost"][b]))


def test_emulated_kernel_and_port_share_the_start_rules(emu, pkg):
    """The kernel and the port carry the same start-point rules (guarded rollout, boundary-slack start): on IAC tracking
    N = 40, where the rollouts leave the track, their iteration counts stay within two of each other and below 11."""
    od, veh, cfg, track, mode = make_oracle(pkg, "iac_tracking")
    batch = pkg.workload.make_batch(veh, cfg, 12, 0xB200 + 3, track, pkg.workload.load_laps(), mode=mode)
    for b in range(12):
        inp = pkg.workload.instance(batch, b)
        p = od.step(inp, impl="port")
        k = _emu_solve(emu, pkg, od, veh, cfg, inp)
        assert p["status"] == 0 and k["status"] == 0
        assert abs(k["iters"] - p["iters"]) <= 2 and k["iters"] <= 11, (b, k["iters"], p["iters"])


@pytest.mark.parametrize("name,N,over", [("iac_tracking", 80, {}), ("barc_lmpc", 20, {"q_boundary": 0.0}),
                                          ("barc_tracking", 20, {"q_boundary": 0.0}), ("iac_tracking", 40, {"q_boundary": 0.0})],
                         ids=["iac_tracking_shipped_n80", "barc_lmpc_hard_boundary", "barc_tracking_hard_boundary", "iac_tracking_hard_boundary"])
def test_emulated_qp_kernel_shipped_horizon_and_hard_boundary(emu, pkg, name, N, over):
    """iac_car_tracking_mpc.param.yaml ships n: 80 (:7) -- the kernel's run-time layout must take it (LMPC_MAX_N = 128);
    q_boundary = 0 selects the HARD track boundary without sigma_b (racing_mpc.cpp:540-542).  Every instance must be
    certified by the dense oracle and matched: nothing is skipped."""
    from conftest import make_case
    from oracle import Oracle
    veh, cfg, track, mode = make_case(pkg, name, None, N)
    cfg = dict(cfg, **over)
    od = Oracle(veh, dict(cfg, tol=1e-11))
    laps = pkg.workload.load_laps()
    if cfg["learning"]:
        for l in laps:
            od.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    batch = pkg.workload.make_batch(veh, cfg, 5, 0x80, track, laps, mode=mode)
    for b in range(5):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        assert d["status"] == 0 and d["polished"] == 1 and d["kkt"] < 1e-9, (b, d["status"], d["kkt"])
        k = _emu_solve(emu, pkg, od, veh, dict(cfg, tol=1e-9), inp)
        assert k["status"] == 0, (b, k["status"])
        e = max(relerr(k["X"], d["X"]), relerr(k["U"], d["U"]), relerr(k["dU"], d["dU"]))
        assert e < 1e-9, (b, e)
        if over.get("q_boundary") == 0.0:   # hard rows hold exactly (to rounding) on every stage 1..N-1
            m = cfg["margin"] + veh["chassis_b"] / 2
            ey = k["X"][1:, 1]
            assert (ey <= inp["bound_left"][1:] - m + 1e-9).all() and (ey >= inp["bound_right"][1:] + m - 1e-9).all()


def test_emulated_qp_kernel_abandons_a_diverging_polish(emu, pkg):
    """One instance per ~1000 used to set the time of a one-wave batch: its first polish came apart (2 -> 4 -> 75 rows
    changing side) and ran all six rounds before the interior point resumed (19 trips).  A polish whose active set is
    diverging is abandoned at once (kernel and port alike): 16 trips, same certified optimum.  Instance 39 of the batch
    rank 3 solves in bench.py (seed 0xB200 + 2 + 7919 * 3)."""
    od, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc", tol=1e-11)
    batch = pkg.workload.make_batch(veh, cfg, 40, 0xB200 + 2 + 7919 * 3, track, pkg.workload.load_laps(), mode=mode)
    inp = pkg.workload.instance(batch, 39)
    k = _emu_solve(emu, pkg, od, veh, cfg, inp)
    d = od.step(inp, impl="dense")
    p = od.step(inp, impl="port")
    assert k["status"] == 0 and d["status"] == 0 and d["kkt"] < 1e-9 and p["status"] == 0
    assert k["iters"] <= 17, k["iters"]
    assert abs(p["iters"] - k["iters"]) <= 2          # the port carries the same rule
    assert max(relerr(k["X"], d["X"]), relerr(k["U"], d["U"]), relerr(k["dU"], d["dU"])) < 1e-8


@pytest.mark.parametrize("integrator", [0, 1])
def test_staged_linearisation_equals_the_plain_form_bit_for_bit(emu, pkg, integrator):
    """The linearisation kernel accumulates the stage tangents in its (strided) output column and overwrites the previous
    stage's tangents in place (lmpc_linearise_staged); the model header's plain form keeps three 6 x 8 arrays.  Same
    products, same sums, same order: identical bits (compiled without contraction here; A[:, 0] = e0 exactly)."""
    from racing_lmpc_ros2_b200 import binding as Bd
    veh = dict(pkg.configs.BARC_VEHICLE, integrator=integrator)
    vs = Bd.fill_struct(Bd.VehicleParams(), veh)
    rng = np.random.default_rng(5)
    for _ in range(100):
        x = np.array([rng.uniform(0, 10), rng.uniform(-.3, .3), rng.uniform(-.3, .3), rng.uniform(0.5, 2.5), rng.uniform(-.2, .2), rng.uniform(-1, 1)])
        u = np.array([rng.uniform(-1, 1), rng.uniform(-.3, .3)])
        kap, dt = rng.uniform(-1, 1), rng.uniform(0.01, 0.1)
        A = np.zeros(36); B = np.zeros(12); g = np.zeros(6); xn = np.zeros(6)
        emu.emu_linearise(C.byref(vs), _p(x), _p(u), C.c_double(kap), C.c_double(dt), _p(A), _p(B), _p(g), _p(xn))
        for stride in (1, 65):
            out = np.full(54 * stride, np.nan)
            emu.emu_linearise_staged(C.byref(vs), _p(x), _p(u), C.c_double(kap), C.c_double(dt), _p(out), stride)
            got = out[::stride][:54]
            assert np.array_equal(got[:36], A) and np.array_equal(got[36:48], B) and np.array_equal(got[48:], g)
        assert np.array_equal(A.reshape(6, 6, order="F")[:, 0], np.eye(6)[0])


def test_emulated_qp_kernel_tolerance_below_the_floor_of_double_precision(emu, pkg):
    """A caller's tol of 1e-15 cannot be met by the primal residual (its own rounding floor is 1e-13): the interior point
    hands over to the polish once the complementarity is below 1e-13 instead of iterating on noise until max_iter (found
    when the boundary slack became a free variable: the last iterations then converge quadratically, mu 6e-10 -> 5e-16 ->
    3e-27, and ran past every test)."""
    od, veh, cfg, track, mode = make_oracle(pkg, "barc_lmpc", tol=1e-11)
    batch = pkg.workload.make_batch(veh, cfg, 6, 0x7071, track, pkg.workload.load_laps(), mode=mode)
    worst, its = 0.0, []
    for b in range(6):
        inp = pkg.workload.instance(batch, b)
        d = od.step(inp, impl="dense")
        if not (d["status"] == 0 and d["polished"] == 1 and d["kkt"] < 1e-9):
            continue
        k = _emu_solve(emu, pkg, od, veh, dict(cfg, tol=1e-15), inp)
        assert k["status"] == 0, (b, k["status"], k["iters"])
        worst = max(worst, relerr(k["X"], d["X"]), relerr(k["U"], d["U"]))
        its.append(k["iters"])
    assert len(its) >= 4 and max(its) <= 16 and worst < 1e-9, (its, worst)
