"""The header-only C++ adapter (racing-lmpc-ros2_b200/cpp/racing_mpc_b200.hpp: RacingMPC::solve(in, out, stats) with
the reference's key strings, racing_mpc.hpp:46-58) compiled with g++ against the C-ABI library and driven by
tests/cpp_adapter_test.cpp.  Without a GPU the constructor must throw (no CPU fallback); with one, a tick solved
through the adapter must equal the same tick through the Python mirror of the interface."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_case

PKG = os.path.join(ROOT, "racing-lmpc-ros2_b200")


@pytest.fixture(scope="module")
def adapter_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("adapter") / "cpp_adapter_test")
    lib = os.path.join(PKG, "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(PKG, "cpp"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_adapter_test.cpp"), "-o", exe, "-L", lib, "-llmpc_b200",
                           "-Wl,-rpath," + lib])
    return exe


def _write_tick(pkg, path, name="barc_lmpc", seed=0xADA):
    from racing_lmpc_ros2_b200 import binding as Bd
    veh, cfg, track, mode = make_case(pkg, name, None, None)
    laps = pkg.workload.load_laps()
    batch = pkg.workload.make_batch(veh, cfg, 1, seed, track, laps, mode=mode)
    inp = pkg.workload.instance(batch, 0)
    N = cfg["N"]
    with open(path, "wb") as f:
        f.write(bytes(Bd.fill_struct(Bd.MpcConfig(), cfg)))
        f.write(bytes(Bd.fill_struct(Bd.VehicleParams(), veh)))
        f.write(struct.pack("i", len(laps)))
        for l in laps:
            n = l["x"].shape[0]
            f.write(struct.pack("i", n)); f.write(struct.pack("d", track["length"]))
            for key in ("x", "u", "k", "t"):
                f.write(np.ascontiguousarray(l[key], dtype=np.float64).tobytes())
        f.write(struct.pack("d", float(inp["total_length"])))
        f.write(np.ascontiguousarray(inp["x_ic"]).tobytes()); f.write(np.ascontiguousarray(inp["u_ic"]).tobytes())
        f.write(struct.pack("d", 0.0))                                   # t_ic
        for key, n in (("X_ref", 6 * N), ("U_ref", 2 * (N - 1)), ("T_ref", N - 1), ("bound_left", N), ("bound_right", N),
                       ("curvatures", N), ("vel_ref", N)):
            a = np.ascontiguousarray(inp[key], dtype=np.float64)
            assert a.size == n, key
            f.write(a.tobytes())
    return veh, cfg, track, batch


def test_adapter_compiles_and_refuses_without_a_gpu(pkg, adapter_exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    _write_tick(pkg, str(tmp_path / "tick.bin"))
    r = subprocess.run([adapter_exe, str(tmp_path / "tick.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "CTOR_THROW" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["qp", "full"])
def test_adapter_tick_equals_python_mirror(pkg, adapter_exe, tmp_path, mode):
    """mode "qp": RacingMPC(config, model, false) -> lmpc_solve_batch; "full": full_dynamics = true -> lmpc_solve_sqp_batch
    (what the node uses for its first tick, racing_mpc_node.cpp:53-56,299-314)."""
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    veh, cfg, track, batch = _write_tick(pkg, str(tmp_path / "tick.bin"))
    r = subprocess.run([adapter_exe, str(tmp_path / "tick.bin")] + (["full"] if mode == "full" else []),
                       capture_output=True, text=True, timeout=300)
    m = BatchedRacingMPC(veh, cfg, max_batch=1)
    for l in pkg.workload.load_laps():
        m.add_lap(l["x"], l["u"], l["k"], l["t"], track["length"])
    out = m.solve(batch) if mode == "qp" else m.solve_sqp(batch, max_sqp_iter=30, tol=1e-9)
    if out["status"][0] != 0:            # not solved: the adapter must omit X_optm too (racing_mpc.cpp:358-371)
        assert r.returncode == 1 and "FAIL: not solved" in r.stdout, (r.returncode, r.stdout)
        assert mode == "full"            # the plain tick of this seed is known to solve
        return
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    m1 = re.search(r"OK iters=(\S+) cost=(\S+) x1=(\S+)", r.stdout)
    assert m1 and "OK2" in r.stdout, r.stdout
    assert abs(float(m1.group(2)) - out["cost"][0]) <= 1e-10 * max(1.0, abs(out["cost"][0]))
    assert abs(float(m1.group(3)) - out["X_optm"][0][cfg["N"] - 1][3]) <= 1e-10
    if mode == "qp":
        assert int(float(m1.group(1))) == int(out["iters"][0])
    # out["ss_x"]: the unpadded query result (6 x found columns), racing_mpc.cpp:249-257
    ms = re.search(r"SS cols=(\d+) found_rows=(\d+)", r.stdout)
    assert ms and int(ms.group(1)) == m.ss_tick_count() == 96 and int(ms.group(2)) == 6, r.stdout
    # the model surface: factory key, discrete_dynamics / _jacobian (g = xip1 - A x - B u), control maps, identity state map
    mm = re.search(r"MODEL xip1_3=(\S+) g_resid=(\S+) fd=(\S+) fb=(\S+) back=(\S+) same_state=(\d)", r.stdout)
    assert mm, r.stdout
    inp = pkg.workload.instance(batch, 0)
    xn = m.discrete_dynamics(inp["X_ref"][0], inp["U_ref"][0], inp["curvatures"][0], inp["T_ref"][0])
    assert abs(float(mm.group(1)) - xn[0][3]) <= 1e-11 and float(mm.group(2)) < 1e-12 and mm.group(6) == "1"
    ul = inp["U_ref"][0][0]
    fd, fb = ul / (1 + np.exp(-ul)), ul / (1 + np.exp(ul))          # single_track_planar_model.cpp:395-400
    assert abs(float(mm.group(3)) - fd) < 1e-11 * max(1, abs(fd)) and abs(float(mm.group(4)) - fb) < 1e-11 * max(1, abs(fb))
    assert abs(float(mm.group(5)) - (fd if abs(fd) > abs(fb) else fb)) < 1e-11
    # create_warm_start (racing_mpc.cpp:374-430) on a synthetic path: speed ramp, omega = v / R, force from the segment's
    # acceleration, pure-pursuit steering; its two exceptions
    mw = re.search(r"WARM vx_last=(\S+) omega1=(\S+) u0=(\S+) steer=(\S+)", r.stdout)
    N = cfg["N"]
    v = np.linspace(1.0, 2.0, N)
    d0 = np.hypot(0.5, 0.01)
    f0 = veh["mass"] * (v[1] ** 2 - v[0] ** 2) / (2 * d0)
    assert mw and abs(float(mw.group(1)) - 2.0) < 1e-11 and abs(float(mw.group(2)) - v[1] / 12.0) < 1e-11      # printed with 12 digits
    assert abs(float(mw.group(3)) - f0) < 1e-10 * abs(f0) and abs(float(mw.group(4)) - np.arctan(veh["wheel_base"] / 12.0)) < 1e-11
    assert "WARM_ERRORS 1 1" in r.stdout, r.stdout


def test_integration_snippet_compiles_against_stub_casadi(pkg, tmp_path):
    """INTEGRATION.md section 3 (the reference-side binding: config conversion + the casadi::DMDict overload of solve)
    compiled with -DLMPC_HAVE_CASADI against tests/stub_casadi; on a box without a GPU the constructor refuses."""
    import torch
    exe = str(tmp_path / "snippet")
    lib = os.path.join(PKG, "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-DLMPC_HAVE_CASADI", "-I", os.path.join(ROOT, "tests", "stub_casadi"),
                           "-I", os.path.join(PKG, "cpp"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_integration_snippet.cpp"), "-o", exe, "-L", lib, "-llmpc_b200", "-Wl,-rpath," + lib])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "MISSING_KEY_THROWS" in r.stdout, (r.returncode, r.stdout, r.stderr)
    else:
        assert r.returncode == 3 and "CTOR_THROW" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_control_maps_match_the_reference_formulas(pkg):
    """a5: to_base_control / from_base_control through the C ABI against the formulas of single_track_planar_model.cpp:390-407
    (logistic split without the x1000; the larger-magnitude force comes back), bit for bit in fp64 up to exp's last ulp."""
    from racing_lmpc_ros2_b200.solver import BatchedRacingMPC
    veh, cfg, track, mode = make_case(pkg, "barc_tracking", None, None)
    m = BatchedRacingMPC(veh, cfg, max_batch=4)
    rng = np.random.default_rng(5)
    u = np.column_stack([np.concatenate([rng.uniform(-3, 3, 500), [0.0, 1e-300, -1e-3, 40.0, -40.0]]), rng.uniform(-0.4, 0.4, 505)])
    ub = m.to_base_control(u)
    fd, fb = u[:, 0] / (1 + np.exp(-u[:, 0])), u[:, 0] / (1 + np.exp(u[:, 0]))
    assert np.abs(ub[:, 0] - fd).max() <= 4e-16 * max(1, np.abs(fd).max()) and np.abs(ub[:, 1] - fb).max() <= 4e-16 * max(1, np.abs(fb).max())
    assert np.array_equal(ub[:, 2], u[:, 1])
    back = m.from_base_control(ub)
    assert np.array_equal(back[:, 0], np.where(np.abs(ub[:, 0]) > np.abs(ub[:, 1]), ub[:, 0], ub[:, 1])) and np.array_equal(back[:, 1], u[:, 1])
    # the quirk App. D records: small commands come back halved (u_lon * sigma(u_lon) ~ u_lon / 2)
    small = np.array([[0.01, 0.0]])
    assert abs(m.from_base_control(m.to_base_control(small))[0, 0] - 0.01 / (1 + np.exp(-0.01))) < 1e-18
