"""Pins the CPU oracle's model: sympy symbolic Jacobian of f, central differences of the discrete map,
structural facts of the linearisation, align_abscissa, and the recorded-lap sanity fixture."""
import numpy as np
import pytest

from conftest import make_oracle


def _sympy_f():
    import sympy as sp
    x = sp.symbols("s ey ephi vx vy om", real=True)
    u = sp.symbols("ul de", real=True)
    kap = sp.Symbol("kappa", real=True)
    prm = sp.symbols("m Jzz l cgr h fr kd kb rho Af cd clf clr mu Bf Cf Br Cr", real=True)
    m, Jzz, l, cgr, h, fr, kd, kb, rho, Af, cd, clf, clr, mu, Bf, Cf, Br, Cr = prm
    s, ey, ephi, vx, vy, om = x
    ul, de = u
    g = sp.Float(9.8)
    lr = cgr * l
    lf = l - lr
    fd = ul * (sp.tanh(ul) / 2 + sp.Rational(1, 2)) * 1000
    fb = ul * (sp.tanh(-ul) / 2 + sp.Rational(1, 2)) * 1000
    vsq = vx * vx
    Fxf = kd * fd / 2 + kb * fb / 2 - fr * m * g * lr / l / 2
    Fxr = (1 - kd) * fd / 2 + (1 - kb) * fb / 2 - fr * m * g * lf / l / 2
    ax = (fd + fb - cd * Af * vsq / 2 - fr * m * g) / m
    Fzf = m * g * lr / (lf + lr) / 2 - h / (lf + lr) * m * ax / 2 + clf * rho * Af * vsq / 4
    Fzr = m * g * lf / (lf + lr) / 2 + h / (lf + lr) * m * ax / 2 + clr * rho * Af * vsq / 4
    af = de - sp.atan((lf * om + vy) / (vx + sp.Float(1e-3)))
    ar = sp.atan((lr * om - vy) / (vx + sp.Float(1e-3)))
    Fyf = mu * Fzf * sp.sin(Cf * sp.atan(Bf * af))
    Fyr = mu * Fzr * sp.sin(Cr * sp.atan(Br * ar))
    omd = (-(2 * Fyr) * lr + ((2 * Fyf) * sp.cos(de) + (2 * Fxf) * sp.sin(de)) * lf) / Jzz
    vxd = ((2 * Fxr) + (2 * Fxf) * sp.cos(de) - (2 * Fyf) * sp.sin(de) - cd * rho * Af * vsq / 2) / m + om * vy
    vyd = ((2 * Fyr) + (2 * Fyf) * sp.cos(de) + (2 * Fxf) * sp.sin(de)) / m - om * vx
    sd = (vx * sp.cos(ephi) - vy * sp.sin(ephi)) / (1 - ey * kap)
    eyd = vx * sp.sin(ephi) + vy * sp.cos(ephi)
    phd = om - kap * sd
    f = sp.Matrix([sd, eyd, phd, vxd, vyd, omd])
    J = f.jacobian(sp.Matrix(list(x) + list(u)))
    args = list(x) + list(u) + [kap] + list(prm)
    return sp.lambdify(args, f, "numpy"), sp.lambdify(args, J, "numpy")


@pytest.fixture(scope="module")
def sym():
    return _sympy_f()


def _prm(veh):
    return [veh[k] for k in ("mass", "moi", "wheel_base", "cg_ratio", "cg_height", "fr", "kd", "kb", "air_density",
                             "frontal_area", "drag_coeff", "cl_f", "cl_r", "mu", "Bf", "Cf", "Br", "Cr")]


def _rand_point(rng, name):
    if name == "iac_tracking":
        x = np.array([rng.uniform(0, 2800), rng.uniform(-3, 3), rng.uniform(-.1, .1), rng.uniform(5, 90), rng.uniform(-3, 3), rng.uniform(-.5, .5)])
        u = np.array([rng.uniform(-10, 5), rng.uniform(-.2, .2)]); k = rng.uniform(-.05, .05)
    else:
        x = np.array([rng.uniform(0, 17), rng.uniform(-.3, .3), rng.uniform(-.3, .3), rng.uniform(.2, 3), rng.uniform(-.5, .5), rng.uniform(-2, 2)])
        u = np.array([rng.uniform(-.01, .01), rng.uniform(-.3, .3)]); k = rng.uniform(-1, 1)
    return x, u, k


@pytest.mark.parametrize("name", ["barc_lmpc", "iac_tracking"])
def test_dynamics_match_sympy(pkg, sym, name):
    """f(x,u,k) of the oracle == an independent sympy transcription of the reference's formulas."""
    o, veh, cfg, _, _ = make_oracle(pkg, name, with_laps=False)
    f_sym, _ = sym
    rng = np.random.default_rng(3)
    for _ in range(50):
        x, u, k = _rand_point(rng, name)
        ref = np.asarray(f_sym(*x, *u, k, *_prm(veh)), dtype=float).ravel()
        got = o.dynamics(x, u, k)
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("name", ["barc_lmpc", "iac_tracking"])
def test_euler_jacobian_matches_sympy(pkg, sym, name):
    """With the Euler integrator A = I + dt df/dx, B = dt df/du: the dual-number Jacobian of the oracle
    must equal sympy's symbolic Jacobian (<= 1e-9 relative)."""
    from oracle import Oracle
    _, veh, cfg, _, _ = make_oracle(pkg, name, with_laps=False)
    veh = dict(veh, integrator=1)
    o = Oracle(veh, cfg)
    _, J_sym = sym
    rng = np.random.default_rng(4)
    dt = 0.025
    for _ in range(30):
        x, u, k = _rand_point(rng, name)
        J = np.asarray(J_sym(*x, *u, k, *_prm(veh)), dtype=float)
        A, B, g, xn = o.linearise(x, u, k, dt)
        Aref = np.eye(6) + dt * J[:, :6]
        Bref = dt * J[:, 6:]
        sc = max(1.0, np.abs(Aref).max(), np.abs(Bref).max())
        assert np.abs(A - Aref).max() <= 1e-9 * sc
        assert np.abs(B - Bref).max() <= 1e-9 * sc


@pytest.mark.parametrize("name", ["barc_lmpc", "iac_tracking"])
def test_rk4_jacobian_central_differences_and_structure(pkg, name):
    o, veh, cfg, _, _ = make_oracle(pkg, name, with_laps=False)
    rng = np.random.default_rng(5)
    dt = 0.025
    for _ in range(20):
        x, u, k = _rand_point(rng, name)
        A, B, g, xn = o.linearise(x, u, k, dt)
        assert np.allclose(xn, o.discrete_dynamics(x, u, k, dt), rtol=0, atol=1e-13 * max(1, np.abs(xn).max()))
        # g = x+ - A x - B u  (single_track_planar_model.cpp:379)
        assert np.allclose(g, xn - A @ x - B @ u, atol=1e-9 * max(1, np.abs(x).max()))
        # f does not depend on s: first column of A is e0
        assert np.array_equal(A[:, 0], np.eye(6)[:, 0])
        for j in range(8):
            h = 1e-6 * max(1.0, abs(x[j]) if j < 6 else abs(u[j - 6]))
            xp, xm, up, um = x.copy(), x.copy(), u.copy(), u.copy()
            if j < 6:
                xp[j] += h; xm[j] -= h
            else:
                up[j - 6] += h; um[j - 6] -= h
            fd = (o.discrete_dynamics(xp, up, k, dt) - o.discrete_dynamics(xm, um, k, dt)) / (2 * h)
            col = A[:, j] if j < 6 else B[:, j - 6]
            assert np.abs(fd - col).max() <= 1e-6 * max(1.0, np.abs(col).max())


def test_align_abscissa(pkg):
    o, *_ = make_oracle(pkg, "barc_lmpc", with_laps=False)
    L = 17.0
    assert o.align_abscissa(1.0, 2.0, L) == pytest.approx(1.0)
    assert o.align_abscissa(16.5, 0.5, L) == pytest.approx(-0.5)
    assert o.align_abscissa(0.5, 16.5, L) == pytest.approx(17.5)
    assert o.align_abscissa(35.0, 1.0, L) == pytest.approx(1.0)
    assert o.align_abscissa(3.0, 3.0, L) == pytest.approx(3.0)   # sign(0) = 0
    rng = np.random.default_rng(0)
    for _ in range(200):
        s1, s2 = rng.uniform(-60, 60), rng.uniform(-60, 60)
        a = o.align_abscissa(s1, s2, L)
        assert abs(a - s2) <= L / 2 + 1e-9
        assert abs(((a - s1) / L) - round((a - s1) / L)) < 1e-9


def test_recorded_laps_are_a_loose_dynamics_fixture(pkg, laps, barc_track):
    """one RK4 step reproduces the recorded samples only loosely (plant ran at another
    rate / frame); s and e_y must still agree to ~1e-3 m in the median."""
    o, *_ = make_oracle(pkg, "barc_lmpc", with_laps=False)
    lap = laps[2]
    ds, de = [], []
    for j in range(0, lap["x"].shape[0] - 1, 7):
        dt = lap["t"][j + 1] - lap["t"][j]
        xn = o.discrete_dynamics(lap["x"][j], lap["u"][j], lap["k"][j], dt)
        ds.append(abs(xn[0] - lap["x"][j + 1, 0])); de.append(abs(xn[1] - lap["x"][j + 1, 1]))
    assert np.median(ds) < 2e-3 and np.median(de) < 5e-4
