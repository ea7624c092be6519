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
