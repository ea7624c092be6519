"""Track interpolation functions (RacingTrajectory, racing_trajectory.cpp:25-236).

Three independent constructions of the reference's degree-3 not-a-knot interpolants must agree:
  product : csrc/lmpc_track.cuh -- piecewise-polynomial form from a tridiagonal solve (run here through the
            CPU build of the same source, tests/emu; on the GPU in test_gpu_parity.py)
  oracle  : oracle/oracle_track.py -- B-spline collocation + Cox-de Boor
  scipy   : scipy.interpolate.make_interp_spline(k=3) (not-a-knot by default)
plus the reference's own test idea (test_racing_trajectory.cpp): frenet <-> global round trips."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "liblmpc_emu.so")
DP = C.POINTER(C.c_double)
TRACKS = ["barc_center", "barc_optm", "putnam_optm"]


def _p(a):
    return a.ctypes.data_as(DP)


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU_DIR, "emu_core.cpp")] + [
        os.path.join(ROOT, "racing-lmpc-ros2_b200", "csrc", f)
        for f in ("lmpc_qp_core.cuh", "lmpc_ss_core.cuh", "lmpc_model.cuh", "lmpc_warp.cuh", "lmpc_host_params.h",
                  "lmpc_track.cuh")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", EMU_SO, srcs[0]])
    L = C.CDLL(EMU_SO)
    L.emu_track_length.restype = C.c_double
    return L


def table(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "tracks.npz"))
    return np.ascontiguousarray(z[f"{name}_table"], dtype=np.float64)


def emu_eval(emu, tb, s):
    assert emu.emu_track_build(tb.shape[0], tb.shape[1], _p(tb)) == 0
    s = np.ascontiguousarray(s, dtype=np.float64)
    out = np.zeros((len(s), 7))
    emu.emu_track_eval(len(s), _p(s), _p(out))
    return dict(left=out[:, 0], right=out[:, 1], curvature=out[:, 2], vel=out[:, 3], x=out[:, 4], y=out[:, 5], yaw=out[:, 6])


@pytest.mark.parametrize("name", TRACKS)
def test_three_constructions_agree(emu, name):
    from oracle_track import OracleTrack
    from scipy.interpolate import make_interp_spline
    tb = table(name)
    o = OracleTrack(tb)
    L = o.L
    rng = np.random.default_rng(1)
    s = np.concatenate([rng.uniform(-1.5 * L, 2.5 * L, 300), tb[:5, 6], [0.0, L, L / 2, -L, 2 * L, 1e-9, L - 1e-9]])
    a = emu_eval(emu, tb, s)
    assert emu.emu_track_m() == tb.shape[0] + 7 and emu.emu_track_length() == L
    b = o.eval(s)
    sm = o.wrap(s)
    sx, sy = make_interp_spline(o.grid, o.x_i(o.grid), k=3), make_interp_spline(o.grid, o.y_i(o.grid), k=3)
    scale = max(1.0, np.abs(tb[:, 0:2]).max())
    for key, tol in (("left", 1e-10), ("right", 1e-10), ("vel", 1e-10), ("x", 1e-10 * scale), ("y", 1e-10 * scale)):
        assert np.abs(a[key] - b[key]).max() < tol * max(1.0, np.abs(b[key]).max()), key
    assert np.abs(a["x"] - sx(sm)).max() < 1e-10 * scale and np.abs(a["y"] - sy(sm)).max() < 1e-10 * scale
    dyaw = np.arctan2(np.sin(a["yaw"] - b["yaw"]), np.cos(a["yaw"] - b["yaw"]))
    assert np.abs(dyaw).max() < 1e-8
    # second derivatives of a cubic interpolant amplify round-off by 1/h^2; curvature is O(1/track radius)
    assert np.abs(a["curvature"] - b["curvature"]).max() < 1e-7 * max(1.0, np.abs(b["curvature"]).max())
    dx, dy, d2x, d2y = sx(sm, 1), sy(sm, 1), sx(sm, 2), sy(sm, 2)
    ref = dx * d2y - dy * d2x / np.sqrt((dx * dx + dy * dy) ** 3)     # racing_trajectory.cpp:108-110 as written
    assert np.abs(a["curvature"] - ref).max() < 1e-7 * max(1.0, np.abs(ref).max())
    # the interpolants pass through the table rows
    at = emu_eval(emu, tb, tb[1:, 6])
    assert np.abs(at["x"] - tb[1:, 0]).max() < 1e-9 * scale and np.abs(at["vel"] - tb[1:, 4]).max() < 1e-9 * max(1, tb[:, 4].max())
    assert np.abs(at["left"] - np.linalg.norm(tb[1:, 0:2] - tb[1:, 9:11], axis=1)).max() < 1e-9


@pytest.mark.parametrize("name", TRACKS)
def test_frenet_global_round_trip_and_oracle(emu, name):
    """test_racing_trajectory.cpp's check (frenet -> global -> frenet is the identity), and the projection against
    the oracle's bracketing search."""
    from oracle_track import OracleTrack
    tb = table(name)
    o = OracleTrack(tb)
    L = o.L
    assert emu.emu_track_build(tb.shape[0], tb.shape[1], _p(tb)) == 0
    rng = np.random.default_rng(2)
    n = 200
    hw = 0.3 if L < 100 else 3.0
    f = np.column_stack([rng.uniform(0.01 * L, 0.99 * L, n), rng.uniform(-hw, hw, n), rng.uniform(-0.5, 0.5, n)])
    g = np.zeros((n, 3)); f2 = np.zeros((n, 3))
    emu.emu_track_f2g(n, _p(f), _p(g))
    go = o.frenet_to_global(f)
    scale = max(1.0, np.abs(tb[:, 0:2]).max())
    assert np.abs(g[:, :2] - go[:, :2]).max() < 1e-9 * scale and np.abs(np.sin(g[:, 2] - go[:, 2])).max() < 1e-8
    emu.emu_track_g2f(n, _p(g), _p(f2))
    assert np.abs(f2[:, 0] - f[:, 0]).max() < 1e-7 * max(1.0, L / 100) and np.abs(f2[:, 1] - f[:, 1]).max() < 1e-8
    assert np.abs(f2[:, 2] - f[:, 2]).max() < 1e-7
    fo = o.global_to_frenet(g[:24])
    assert np.abs(fo[:, 0] - f2[:24, 0]).max() < 1e-7 * max(1.0, L / 100) and np.abs(fo[:, 1:] - f2[:24, 1:]).max() < 1e-7
