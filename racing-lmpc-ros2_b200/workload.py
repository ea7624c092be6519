"""Synthetic, seeded MPC-tick batches of the shape BASELINE.json names.

Host-side numpy only: this builds INPUTS (initial states, references, bounds); nothing here is
on the solve path.  Fixtures come from tests/golden/*.npz (recorded BARC laps, track tables).
Every array is instance-major: X_ref[b] is the reference's 6 x N column-major DM, i.e. (N, 6)
row-major here.
"""
import os
import numpy as np

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_laps():
    z = np.load(os.path.join(_GOLD, "barc_ss_laps.npz"))
    return [dict(x=z[f"lap{i}_x"], u=z[f"lap{i}_u"], k=z[f"lap{i}_k"].ravel(), t=z[f"lap{i}_t"].ravel())
            for i in (1, 2, 3)]


def synthesise_laps(laps, n_laps, seed=0xB200 + 4, sigma_ey=0.02, sigma_vx=0.05):
    """BASELINE configs[3]'s "50-lap learned safe set" (SURVEY.md 8d): the recorded laps cycled, each copy moved by ONE
    Gaussian offset per lap in e_y and v_x (sigma 0.02 m / 0.05 m/s, fixed seed) -> n_laps laps of ~440 samples, ~66 k
    tripled points for 50.  The offset is per lap, not per sample: consecutive samples of a lap stay consistent with the
    car's dynamics, so the one-step model error the error-dynamics regression learns from keeps the size it has in the
    recorded data (v_x 3e-3 rms per step).  Round 1 drew the noise per sample, which made that error 8e-2 -- the regression
    then corrected (A, B, g) with noise and 13 % of the QPs of the 50-lap configuration failed."""
    rng = np.random.Generator(np.random.Philox(seed))
    out = []
    for j in range(n_laps):
        src = laps[j % len(laps)]
        x = src["x"].copy()
        x[:, 1] += rng.standard_normal() * sigma_ey
        x[:, 3] += rng.standard_normal() * sigma_vx
        out.append(dict(x=x, u=src["u"].copy(), k=src["k"].copy(), t=src["t"].copy()))
    return out


def synthesise_track_laps(track, vehicle, dt, n_laps=3, seed=0xB200 + 6, vel_scale=0.9):
    """Laps for a track without recorded ones (IAC LMPC on Putnam): the table's line driven at vel_scale x its speed
    profile, sampled every dt from s = 0 to L; e_y / e_psi / v_y small seeded noise, omega = v kappa, steering holding the
    local curvature.  Lap j is a little faster than lap j-1 (what a learning run records)."""
    rng = np.random.Generator(np.random.Philox(seed))
    L = track["length"]
    out = []
    for j in range(n_laps):
        sc = vel_scale * (1.0 + 0.02 * j)
        s, t, xs, us, ks, ts = 0.0, 0.0, [], [], [], []
        while s < L:
            v = float(track_lookup(track, s, "speed")) * sc
            k = float(track_lookup(track, s, "curvature"))
            xs.append([s, 0.0, 0.0, v, 0.0, v * k]); us.append([0.5, float(np.arctan(vehicle["wheel_base"] * k))]); ks.append(k); ts.append(t)
            s += v * dt; t += dt
        x = np.array(xs); n = x.shape[0]
        x[:, 1] += rng.standard_normal(n) * 0.1; x[:, 2] += rng.standard_normal(n) * 0.005; x[:, 4] += rng.standard_normal(n) * 0.02
        out.append(dict(x=x, u=np.array(us), k=np.array(ks), t=np.array(ts)))
    return out


def load_track(name):
    z = np.load(os.path.join(_GOLD, "tracks.npz"))
    return dict(s=z[f"{name}_s"], speed=z[f"{name}_speed"], curvature=z[f"{name}_curvature"],
                left=z[f"{name}_left"], right=z[f"{name}_right"], length=float(z[f"{name}_length"]))


def track_lookup(track, s, key):
    """Periodic linear interpolation of a track-table column at abscissa s (any lap)."""
    L = track["length"]
    sm = np.mod(s, L)
    xs = np.concatenate([track["s"], [L]])
    ys = np.concatenate([track[key], track[key][:1]])
    return np.interp(sm, xs, ys)


def dynamics_np(p, x, u, kappa):
    """Vectorised restatement of the single-track model (input synthesis only).
    x (...,6), u (...,2), kappa (...)."""
    g = 9.8
    ey, phi, vx, vy, om = x[..., 1], x[..., 2], x[..., 3], x[..., 4], x[..., 5]
    ul, de = u[..., 0], u[..., 1]
    m, Jzz, l = p["mass"], p["moi"], p["wheel_base"]
    lr = p["cg_ratio"] * l
    lf = l - lr
    fd = ul * (0.5 * np.tanh(ul) + 0.5) * 1000.0
    fb = ul * (0.5 * np.tanh(-ul) + 0.5) * 1000.0
    vsq = vx * vx
    Fxf = 0.5 * p["kd"] * fd + 0.5 * p["kb"] * fb - 0.5 * p["fr"] * m * g * lr / l
    Fxr = 0.5 * (1 - p["kd"]) * fd + 0.5 * (1 - p["kb"]) * fb - 0.5 * p["fr"] * m * g * lf / l
    ax = (fd + fb - 0.5 * p["drag_coeff"] * p["frontal_area"] * vsq - p["fr"] * m * g) / m
    rA = p["air_density"] * p["frontal_area"]
    Fzf = 0.5 * m * g * lr / l - 0.5 * p["cg_height"] / l * m * ax + 0.25 * p["cl_f"] * rA * vsq
    Fzr = 0.5 * m * g * lf / l + 0.5 * p["cg_height"] / l * m * ax + 0.25 * p["cl_r"] * rA * vsq
    af = de - np.arctan((lf * om + vy) / (vx + 1e-3))
    ar = np.arctan((lr * om - vy) / (vx + 1e-3))
    Fyf = p["mu"] * Fzf * np.sin(p["Cf"] * np.arctan(p["Bf"] * af))
    Fyr = p["mu"] * Fzr * np.sin(p["Cr"] * np.arctan(p["Br"] * ar))
    cd, sd = np.cos(de), np.sin(de)
    omd = (-2 * Fyr * lr + (2 * Fyf * cd + 2 * Fxf * sd) * lf) / Jzz
    vxd = (2 * Fxr + 2 * Fxf * cd - 2 * Fyf * sd - 0.5 * p["drag_coeff"] * rA * vsq) / m + om * vy
    vyd = (2 * Fyr + 2 * Fyf * cd + 2 * Fxf * sd) / m - om * vx
    sd_ = (vx * np.cos(phi) - vy * np.sin(phi)) / (1 - ey * kappa)
    eyd = vx * np.sin(phi) + vy * np.cos(phi)
    phd = om - kappa * sd_
    return np.stack([sd_, eyd, phd, vxd, vyd, omd], axis=-1)


def rk4_np(p, x, u, kappa, dt):
    k1 = dynamics_np(p, x, u, kappa)
    k2 = dynamics_np(p, x + 0.5 * dt * k1, u, kappa)
    k3 = dynamics_np(p, x + 0.5 * dt * k2, u, kappa)
    k4 = dynamics_np(p, x + dt * k3, u, kappa)
    return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def make_batch(vehicle, config, B, seed, track, laps=None, dt=0.025, mode="barc",
               vel_scale=0.9, speed=None):
    """Seeded batch of B MPC ticks.

    mode "barc": x* sampled from the newest recorded lap, x_ic = x* + clipped Gaussian noise,
      U_ref = the next N-1 recorded controls, X_ref = RK4 rollout from x_ic, curvatures/bounds/vel_ref
      from the track table at X_ref[:,0] (periodic linear interpolation).
    mode "track": x* sampled along the track table at its speed profile (IAC / Putnam tracking).
    """
    rng = np.random.Generator(np.random.Philox(seed))
    N = int(config["N"])
    L = track["length"]
    xmin = np.array(config["x_min"], dtype=float)
    xmax = np.array(config["x_max"], dtype=float)
    if mode == "barc":
        lap = laps[-1]
        n = lap["x"].shape[0]
        j0 = rng.integers(0, n, size=B)
        xs = lap["x"][j0]
        sig = np.array([0.05, 0.02, 0.02, 0.05, 0.02, 0.05])
        x_ic = xs + rng.standard_normal((B, 6)) * sig
        idx = (j0[:, None] + np.arange(N - 1)[None, :]) % n
        U_ref = lap["u"][idx]                      # (B, N-1, 2)
        u_ic = lap["u"][j0]
    else:
        s0 = rng.uniform(0.0, L, size=B)
        v0 = track_lookup(track, s0, "speed") * vel_scale if speed is None else np.full(B, speed)
        x_ic = np.zeros((B, 6))
        x_ic[:, 0] = s0
        x_ic[:, 1] = rng.standard_normal(B) * 0.2
        x_ic[:, 2] = rng.standard_normal(B) * 0.01
        x_ic[:, 3] = v0 + rng.standard_normal(B) * 0.5
        x_ic[:, 4] = rng.standard_normal(B) * 0.05
        x_ic[:, 5] = v0 * track_lookup(track, s0, "curvature") + rng.standard_normal(B) * 0.005
        # steering that roughly holds the local curvature, mild longitudinal command
        U_ref = np.zeros((B, N - 1, 2))
        U_ref[:, :, 1] = np.arctan(vehicle["wheel_base"] * track_lookup(track, s0, "curvature"))[:, None]
        U_ref[:, :, 0] = 0.5
        u_ic = U_ref[:, 0, :].copy()
    with np.errstate(invalid="ignore"):
        lo = np.where(np.isfinite(xmin), xmin + 1e-3 * np.maximum(1.0, np.abs(xmin)), -np.inf)
        hi = np.where(np.isfinite(xmax), xmax - 1e-3 * np.maximum(1.0, np.abs(xmax)), np.inf)
    x_ic = np.clip(x_ic, lo, hi)
    # keep e_y inside the (margin-shrunk) track
    mrg = config["margin"] + vehicle["chassis_b"] / 2.0 + 0.02
    bl0 = track_lookup(track, x_ic[:, 0], "left") - mrg
    br0 = track_lookup(track, x_ic[:, 0], "right") + mrg
    x_ic[:, 1] = np.clip(x_ic[:, 1], np.minimum(br0, bl0 - 1e-3), np.maximum(bl0, br0 + 1e-3))
    X_ref = np.zeros((B, N, 6))
    X_ref[:, 0] = x_ic
    kap = np.zeros((B, N))
    for i in range(N - 1):
        kap[:, i] = track_lookup(track, X_ref[:, i, 0], "curvature")
        X_ref[:, i + 1] = rk4_np(vehicle, X_ref[:, i], U_ref[:, i], kap[:, i], dt)
    kap[:, N - 1] = track_lookup(track, X_ref[:, N - 1, 0], "curvature")
    s_all = X_ref[:, :, 0]
    batch = dict(
        x_ic=x_ic, u_ic=np.ascontiguousarray(u_ic), X_ref=X_ref, U_ref=np.ascontiguousarray(U_ref),
        T_ref=np.full((B, N - 1), dt), bound_left=track_lookup(track, s_all, "left"),
        bound_right=track_lookup(track, s_all, "right"), curvatures=kap,
        vel_ref=track_lookup(track, s_all, "speed") * vel_scale,
        total_length=np.full(B, L))
    # the node wraps s into [0, L) on the measured state; X_ref may run past L -- solve() re-aligns it
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in batch.items()}


def instance(batch, b):
    """One instance of a batch as the dict the oracle / adapter take."""
    d = {k: v[b] for k, v in batch.items()}
    d["total_length"] = float(batch["total_length"][b])
    return d
