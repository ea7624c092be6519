"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's error-dynamics regression and lap recorder.

  regress()   SSTrajectory::query(const RegQuery&)   racing_trajectory/src/safe_set.cpp:56-114
              SafeSetManager::query(const RegQuery&)  racing_trajectory/src/safe_set.cpp:182-245
  Recorder    SafeSetRecorder::step                   racing_trajectory/src/safe_set.cpp:278-322

Parity unpinned: the reference never calls the regression (no call site outside safe_set.cpp), has no test for it and it
cannot run as written.  Deviations from the text, the same ones the CUDA path states in
csrc/lmpc_reg_core.cpp's header: full-state model prediction (the reference passes sliced states to f.map, :217-220),
scalar target y = x_{p+1}[out] - f_d(...)[out] (the reference slices reg_in_state_idxs, :226), dt_p = t_{p+1} - t_p > 0
(the reference stores the negative, :130-135), `sign` selectable (-1 = b = -M'Ky as written, :229).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np


def lap_points(orc, lap):
    """Samples with a successor (safe_set.cpp:68-77) and their one-step model error."""
    x = np.asarray(lap["x"], dtype=float); u = np.asarray(lap["u"], dtype=float)
    k = np.asarray(lap["k"], dtype=float).ravel(); t = np.asarray(lap["t"], dtype=float).ravel()
    n = x.shape[0]
    err = np.zeros((n - 1, 6))
    for p in range(n - 1):
        err[p] = x[p + 1] - orc.discrete_dynamics(x[p], u[p], k[p], t[p + 1] - t[p])
    return np.hstack([x[:-1], u[:-1]]), err


def regress(points, out_idx, in_x, in_u, dist_max, xq, uq, A, B, C, ridge=1e-3, sign=1.0):
    """points: list over laps (oldest first) of (Z [m,8], E [m,6]) from lap_points.  A (6,6), B (6,2), C (6) nominal.
    Returns corrected copies and the number of samples used per regression."""
    A = np.array(A, dtype=float); B = np.array(B, dtype=float); C = np.array(C, dtype=float)
    zq = np.concatenate([np.asarray(xq, dtype=float), np.asarray(uq, dtype=float)])
    used = []
    h = float(dist_max)
    for r, o in enumerate(out_idx):
        sel = list(in_x[r]) + [6 + c for c in in_u[r]]
        Ms, ys, ds = [], [], []
        for Z, E in points:   # every stored lap (:186-193), concatenated (:196-203)
            z = Z[:, sel]
            d = np.sqrt(((z - zq[sel]) ** 2).sum(axis=1))   # :81-85
            m = d < h                                        # :87
            Ms.append(z[m]); ys.append(E[m, o]); ds.append(d[m])
        z = np.vstack(Ms); y = np.concatenate(ys); d = np.concatenate(ds)
        used.append(len(d))
        if len(d) == 0:       # :203-205
            continue
        K = 0.75 / h * (1.0 - (d / h) ** 2) ** 2             # :222-223
        M = np.hstack([z, np.ones((len(d), 1))])             # :227
        Q = M.T @ (K[:, None] * M) + ridge * np.eye(M.shape[1])   # :228
        b = sign * (M.T @ (K * y))                           # :229 (sign = -1 as written)
        R = np.linalg.solve(Q, b)                            # :231
        nx = len(in_x[r])
        A[o, list(in_x[r])] += R[:nx]                        # :239
        B[o, list(in_u[r])] += R[nx:-1]                      # :240
        C[o] += R[-1]                                        # :241
    return A, B, C, np.array(used)


class Recorder:
    """SafeSetRecorder (safe_set.cpp:246-322) with add_lap replaced by a list of completed laps."""

    def __init__(self):
        self.last_x_valid = False; self.initialized = False; self.lap_count = 0
        self.x = []; self.u = []; self.k = []; self.t = []
        self.laps = []

    def step(self, x, u, k, t, L):
        x = np.asarray(x, dtype=float); u = np.asarray(u, dtype=float)
        if not self.last_x_valid:                    # :282-286
            self.x = [x]; self.last_x_valid = True
            return False
        added = False
        if self.x[-1][0] - x[0] > 0.5 * L:           # :290
            if self.initialized:                     # :293-309
                self.laps.append(dict(x=np.array(self.x), u=np.array(self.u), k=np.array(self.k), t=np.array(self.t)))
                added = True
            else:
                self.initialized = True
            self.lap_count += 1
            self.x = [x]; self.u = [u]; self.t = [t]; self.k = [k]   # :312-315
        else:
            self.x.append(x); self.u.append(u); self.t.append(t); self.k.append(k)   # :317-320
        return added
