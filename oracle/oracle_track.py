"""CPU ORACLE for the track interpolation functions -- TEST INFRASTRUCTURE ONLY (see oracle/lmpc_oracle.h).

Restates RacingTrajectory (reference src/vehicle_dynamics_models/racing_trajectory/src/racing_trajectory.cpp:25-236)
in numpy.  The interpolants are built the way CasADi's `interpolant(..., "bspline")` does it (third party, not under
/root/reference; recalled: BSplineInterpolant, degree 3, knots = not_a_knot(grid), coefficients from the collocation
system B c = values), i.e. as B-splines evaluated by the Cox-de Boor recursion -- deliberately NOT the
piecewise-polynomial / tridiagonal construction of the product (csrc/lmpc_track.cuh).  Both describe the same unique
function; tests/test_track.py additionally compares with scipy.interpolate.make_interp_spline.

PARITY UNPINNED for this row as for the rest: CasADi is absent, and the reference's test for this class
(test_racing_trajectory.cpp) asserts only round trips.
"""
import numpy as np


def align_abscissa(s1, s2, total):
    """lmpc_utils/utils.hpp:35-41"""
    s1 = np.asarray(s1, dtype=float)
    k = np.abs(s2 - s1) + total / 2.0
    l = k - np.fmod(np.abs(s2 - s1) + total / 2.0, total)
    return s1 + l * np.sign(s2 - s1)


def align_yaw(y1, y2):
    """lmpc_utils/utils.hpp:25-31"""
    d = y1 - y2
    return np.arctan2(np.sin(d), np.cos(d)) + y2


def not_a_knot(x, k=3):
    """CasADi BSplineInterpolant::not_a_knot for odd degree: x0 repeated k+1 times, interior grid points without the
    (k-1)/2 next to each end, x_end repeated k+1 times."""
    m = (k - 1) // 2
    return np.concatenate([[x[0]] * (k + 1), x[m + 1:len(x) - m - 1], [x[-1]] * (k + 1)])


def bspline_basis(t, k, x, nu=0):
    """All B-spline basis functions of degree k on knots t at points x (len(x), len(t)-k-1); nu-th derivative.
    Cox-de Boor recursion; the derivative through B'_{i,k} = k (B_{i,k-1}/(t_{i+k}-t_i) - B_{i+1,k-1}/(t_{i+k+1}-t_{i+1}))."""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    t = np.asarray(t, dtype=float)
    nt = len(t)
    # degree 0: indicator of [t_i, t_{i+1}), the last non-empty interval closed on the right
    nonempty = t[:-1] < t[1:]
    last = np.max(np.nonzero(nonempty)[0])
    xc = x[:, None]
    B = ((xc >= t[None, :-1]) & (xc < t[None, 1:]) & nonempty[None, :]).astype(float)
    B[:, last] = np.maximum(B[:, last], (x == t[last + 1]).astype(float))
    for d in range(1, k + 1):
        nd = nt - 1 - d
        a = t[d:d + nd] - t[:nd]
        b = t[d + 1:d + 1 + nd] - t[1:1 + nd]
        ia = np.divide(1.0, a, out=np.zeros_like(a), where=a > 0)
        ib = np.divide(1.0, b, out=np.zeros_like(b), where=b > 0)
        if d > k - nu:          # the last nu levels differentiate
            B = d * (B[:, :nd] * ia - B[:, 1:nd + 1] * ib)
        else:
            B = (xc - t[None, :nd]) * ia * B[:, :nd] + (t[None, d + 1:d + 1 + nd] - xc) * ib * B[:, 1:nd + 1]
    return B


class BSpline1D:
    def __init__(self, grid, values, k=3):
        self.k = k
        self.t = not_a_knot(np.asarray(grid, dtype=float), k)
        A = bspline_basis(self.t, k, grid)
        self.c = np.linalg.solve(A, np.asarray(values, dtype=float))

    def __call__(self, x, nu=0):
        return bspline_basis(self.t, self.k, x, nu) @ self.c


class OracleTrack:
    """table: (n, >=13) array in TrajectoryIndex column order (racing_trajectory.hpp:37-56)."""

    def __init__(self, table):
        tb = np.asarray(table, dtype=float)
        self.table = tb
        self.L = float(tb[0, 7])                                             # :29
        L = self.L
        ext = np.vstack([tb, tb[:4]])                                        # :45
        ext[-4:, 6] += L                                                     # :48-50
        ext = np.vstack([ext[-7:-4], ext])                                   # :53
        ext[:3, 6] -= L                                                      # :56
        s = ext[:, 6]
        self.grid = s
        p = ext[:, 0:2]
        t_left = np.linalg.norm(p - ext[:, 9:11], axis=1)                    # :64-71
        t_right = -np.linalg.norm(p - ext[:, 11:13], axis=1)                 # :72-79
        self.left_i = BSpline1D(s, t_left)
        self.right_i = BSpline1D(s, t_right)
        self.x_i = BSpline1D(s, ext[:, 0])
        self.y_i = BSpline1D(s, ext[:, 1])
        self.vel_i = BSpline1D(s, ext[:, 4])

    def wrap(self, s):
        return align_abscissa(s, self.L / 2.0, self.L)                      # :97

    def eval(self, s):
        """dict of left, right, curvature, vel, x, y, yaw at abscissae s (racing_trajectory.cpp:96-118)."""
        sm = np.atleast_1d(self.wrap(s))
        dx, dy = self.x_i(sm, 1), self.y_i(sm, 1)
        d2x, d2y = self.x_i(sm, 2), self.y_i(sm, 2)
        curv = dx * d2y - dy * d2x / np.sqrt((dx ** 2 + dy ** 2) ** 3)       # :108-110, as written
        return dict(left=self.left_i(sm), right=self.right_i(sm), curvature=curv, vel=self.vel_i(sm),
                    x=self.x_i(sm), y=self.y_i(sm), yaw=np.arctan2(dy, dx))

    def frenet_to_global(self, f):
        f = np.atleast_2d(np.asarray(f, dtype=float))
        e = self.eval(self.wrap(f[:, 0]))                                    # :124-127 (wrapped twice)
        return np.column_stack([e["x"] - np.sin(e["yaw"]) * f[:, 1], e["y"] + np.cos(e["yaw"]) * f[:, 1],
                                align_yaw(e["yaw"] + f[:, 2], 0.0)])

    def global_to_frenet(self, g):
        """:138-186,204-236.  The scalar minimisation of |r(s) - p|^2 is done by bracketing around the nearest way
        point and golden-section search (a different method from the product's Newton iteration), polished by
        bisection on the derivative."""
        g = np.atleast_2d(np.asarray(g, dtype=float))
        out = np.zeros((len(g), 3))
        way = self.table
        n = len(way)
        for q, (px, py, phi) in enumerate(g):
            idx = int(np.argmin((way[:, 0] - px) ** 2 + (way[:, 1] - py) ** 2))
            s0 = float(self.wrap(way[idx, 6]))
            h = 2.0 * self.L / n

            def grad(s):
                e = self.eval(self.wrap(np.array([s])))
                sm = self.wrap(self.wrap(np.array([s])))
                return float((e["x"][0] - px) * self.x_i(sm, 1)[0] + (e["y"][0] - py) * self.y_i(sm, 1)[0])

            a, b = s0 - h, s0 + h
            ga, gb = grad(a), grad(b)
            tries = 0
            while ga * gb > 0 and tries < 8:       # widen until the derivative changes sign
                a -= h; b += h; ga, gb = grad(a), grad(b); tries += 1
            for _ in range(200):
                mid = 0.5 * (a + b)
                gm = grad(mid)
                if ga * gm <= 0:
                    b, gb = mid, gm
                else:
                    a, ga = mid, gm
                if b - a < 1e-14 * max(1.0, abs(mid)):
                    break
            s = float(self.wrap(0.5 * (a + b)))
            e = self.eval(np.array([s]))
            x0, y0, yaw0 = e["x"][0], e["y"][0], e["yaw"][0]
            sg = np.sign(np.cos(yaw0) * (py - y0) - np.sin(yaw0) * (px - x0))   # lateral_sign, utils.hpp:72-80
            out[q] = [s, np.hypot(px - x0, py - y0) * sg, align_yaw(phi, yaw0) - yaw0]
        return out
