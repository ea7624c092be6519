"""ctypes binding of the CPU ORACLE (oracle/liblmpc_oracle.so).

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package never does.
PARITY UNPINNED: see oracle/lmpc_oracle.h for why and for how the oracle is validated instead.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblmpc_oracle.so")
_SRCS = ["oracle_model.c", "oracle_safeset.c", "oracle_qp_dense.c", "oracle_port.c", "oracle_osqp.c",
         "lmpc_oracle.h", "oracle_internal.h", "Makefile"]


def build(force=False):
    """Compile the oracle with gcc (plain C, no dependencies)."""
    stale = force or not os.path.exists(_LIB_PATH)
    if not stale:
        t = os.path.getmtime(_LIB_PATH)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SRCS
                    if os.path.exists(os.path.join(_HERE, s)))
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liblmpc_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Vehicle(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "mass", "moi", "wheel_base", "cg_ratio", "cg_height", "fr", "chassis_b", "kd", "kb",
        "air_density", "frontal_area", "drag_coeff", "cl_f", "cl_r", "mu", "Bf", "Cf", "Br", "Cr",
        "Fd_max", "Fb_max", "Td", "Tb", "max_steer", "max_steer_rate")] + [
        ("integrator", C.c_int), ("pad_", C.c_int)]


class Config(C.Structure):
    _fields_ = [("N", C.c_int), ("learning", C.c_int), ("margin", C.c_double),
                ("q_contour", C.c_double), ("q_heading", C.c_double), ("q_vel", C.c_double),
                ("q_vy", C.c_double), ("q_vyaw", C.c_double), ("q_boundary", C.c_double),
                ("R", C.c_double * 4), ("R_d", C.c_double * 4),
                ("x_max", C.c_double * 6), ("x_min", C.c_double * 6),
                ("u_max", C.c_double * 2), ("u_min", C.c_double * 2),
                ("convex_hull_slack", C.c_double * 6),
                ("num_ss_pts", C.c_int), ("num_ss_pts_per_lap", C.c_int),
                ("max_lap_stored", C.c_int), ("max_iter", C.c_int), ("tol", C.c_double)]


class StepIn(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in (
        "x_ic", "u_ic", "X_ref", "U_ref", "T_ref", "bound_left", "bound_right", "curvatures",
        "vel_ref")] + [("total_length", C.c_double), ("ss_query_point", C.POINTER(C.c_double))]


class StepOut(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in ("X", "U", "dU", "lambda_", "ss_x", "ss_cost")] + [
        ("cost", C.c_double), ("sigma_b", C.c_double), ("sigma_h", C.c_double * 6),
        ("kkt", C.c_double), ("status", C.c_int), ("iters", C.c_int), ("polished", C.c_int)]


def fill_struct(st, d):
    for name, _ in st._fields_:
        key = name
        if key not in d:
            continue
        val = d[key]
        cur = getattr(st, name)
        if hasattr(cur, "__len__"):
            arr = np.asarray(val, dtype=np.float64).ravel()
            for i in range(len(cur)):
                cur[i] = float(arr[i])
        else:
            setattr(st, name, val)
    return st


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.orc_dynamics.argtypes = [C.POINTER(Vehicle), dp, dp, C.c_double, dp]
        L.orc_discrete_dynamics.argtypes = [C.POINTER(Vehicle), dp, dp, C.c_double, C.c_double, dp]
        L.orc_linearise.argtypes = [C.POINTER(Vehicle), dp, dp, C.c_double, C.c_double, dp, dp, dp, dp]
        L.orc_align_abscissa.restype = C.c_double
        L.orc_align_abscissa.argtypes = [C.c_double] * 3
        L.orc_ss_create.restype = C.c_void_p
        L.orc_ss_create.argtypes = [C.c_int]
        L.orc_ss_destroy.argtypes = [C.c_void_p]
        L.orc_ss_add_lap.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, C.c_double]
        L.orc_ss_load.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.orc_ss_num_laps.argtypes = [C.c_void_p]
        L.orc_ss_query.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
        L.orc_ss_query_padded.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
        for fn in (L.orc_step_dense, L.orc_step_port):
            fn.argtypes = [C.POINTER(Vehicle), C.POINTER(Config), C.c_void_p, C.POINTER(StepIn),
                           C.POINTER(StepOut)]
        L.orc_step_batch.argtypes = [C.POINTER(Vehicle), C.POINTER(Config), C.c_void_p, C.c_int] + \
            [dp] * 10 + [dp] * 5 + [C.POINTER(C.c_int), C.POINTER(C.c_int), dp, C.c_int, C.c_int]
        L.orc_step_sqp.argtypes = [C.POINTER(Vehicle), C.POINTER(Config), C.c_void_p, C.POINTER(StepIn),
                                   C.POINTER(StepOut), C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), dp]
        L.orc_step_osqp.argtypes = [C.POINTER(Vehicle), C.POINTER(Config), C.c_void_p, C.POINTER(StepIn), C.POINTER(StepOut),
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, dp]
        L.orc_check_candidate.restype = C.c_double
        L.orc_check_candidate.argtypes = [C.POINTER(Vehicle), C.POINTER(Config), C.c_void_p,
                                          C.POINTER(StepIn), dp, dp, dp, dp, dp, dp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """CPU oracle bound to one vehicle + MPC configuration (dicts as in configs.py)."""

    def __init__(self, vehicle, config):
        self.L = lib()
        self.veh = fill_struct(Vehicle(), vehicle)
        self.cfg = fill_struct(Config(), config)
        self.N = int(config["N"])
        self.K = int(config["num_ss_pts"])
        self.learning = bool(config["learning"])
        self.ss = self.L.orc_ss_create(int(config["max_lap_stored"]))

    def __del__(self):
        try:
            self.L.orc_ss_destroy(self.ss)
        except Exception:
            pass

    # ---- model -------------------------------------------------------------------------
    def dynamics(self, x, u, kappa):
        x, u = _f64(x), _f64(u)
        out = np.zeros(6)
        self.L.orc_dynamics(C.byref(self.veh), _p(x), _p(u), float(kappa), _p(out))
        return out

    def discrete_dynamics(self, x, u, kappa, dt):
        x, u = _f64(x), _f64(u)
        out = np.zeros(6)
        self.L.orc_discrete_dynamics(C.byref(self.veh), _p(x), _p(u), float(kappa), float(dt), _p(out))
        return out

    def linearise(self, x, u, kappa, dt):
        """Returns A (6x6), B (6x2), g (6), x_next (6)."""
        x, u = _f64(x), _f64(u)
        A = np.zeros(36); Bm = np.zeros(12); g = np.zeros(6); xn = np.zeros(6)
        self.L.orc_linearise(C.byref(self.veh), _p(x), _p(u), float(kappa), float(dt), _p(A), _p(Bm), _p(g), _p(xn))
        return A.reshape(6, 6).T.copy(), Bm.reshape(2, 6).T.copy(), g, xn

    def align_abscissa(self, s1, s2, total):
        return self.L.orc_align_abscissa(float(s1), float(s2), float(total))

    # ---- safe set ----------------------------------------------------------------------
    def add_lap(self, x, u, k, t, L):
        x, u, k, t = _f64(x), _f64(u), _f64(k).ravel(), _f64(t).ravel()
        assert x.ndim == 2 and x.shape[1] == 6
        return self.L.orc_ss_add_lap(self.ss, x.shape[0], _p(x), _p(u), _p(k), _p(t), float(L))

    def load_lap(self, prefix, L):
        return self.L.orc_ss_load(self.ss, prefix.encode(), float(L))

    def num_laps(self):
        return self.L.orc_ss_num_laps(self.ss)

    def ss_query(self, s, ey, max_total=None, per_lap=None):
        K = self.K if max_total is None else max_total
        per = self.cfg.num_ss_pts_per_lap if per_lap is None else per_lap
        sx = np.zeros((K, 6)); sj = np.zeros(K)
        cnt = self.L.orc_ss_query(self.ss, float(s), float(ey), K, per, _p(sx), _p(sj))
        return sx[:cnt], sj[:cnt]

    def ss_query_padded(self, s, ey):
        sx = np.zeros((self.K, 6)); sc = np.zeros(self.K)
        cnt = self.L.orc_ss_query_padded(self.ss, float(s), float(ey), self.K,
                                         self.cfg.num_ss_pts_per_lap, _p(sx), _p(sc))
        return sx, sc, cnt

    # ---- one tick ----------------------------------------------------------------------
    def _mk_in(self, inp):
        keep = {}
        si = StepIn()
        for key in ("x_ic", "u_ic", "X_ref", "U_ref", "T_ref", "bound_left", "bound_right",
                    "curvatures", "vel_ref"):
            keep[key] = _f64(inp[key]).ravel()
            setattr(si, key, _p(keep[key]))
        si.total_length = float(inp["total_length"])
        if inp.get("ss_query_point") is not None:
            keep["ss_query_point"] = _f64(inp["ss_query_point"]).ravel()
            si.ss_query_point = _p(keep["ss_query_point"])
        return si, keep

    def step(self, inp, impl="dense"):
        """inp: dict with the reference's keys; X_ref is (N,6) row-major (== 6xN column-major),
        U_ref (N-1,2).  Returns dict with X (N,6), U, dU (N-1,2), lambda, cost, status, ..."""
        N, K = self.N, self.K
        si, keep = self._mk_in(inp)
        X = np.zeros((N, 6)); U = np.zeros((N - 1, 2)); dU = np.zeros((N - 1, 2))
        lam = np.zeros(K); ssx = np.zeros((K, 6)); ssc = np.zeros(K)
        so = StepOut()
        so.X, so.U, so.dU, so.lambda_, so.ss_x, so.ss_cost = _p(X), _p(U), _p(dU), _p(lam), _p(ssx), _p(ssc)
        fn = self.L.orc_step_dense if impl == "dense" else self.L.orc_step_port
        st = fn(C.byref(self.veh), C.byref(self.cfg), self.ss, C.byref(si), C.byref(so))
        return dict(X=X, U=U, dU=dU, lam=lam, ss_x=ssx, ss_cost=ssc, cost=so.cost, sigma_b=so.sigma_b,
                    sigma_h=np.array(list(so.sigma_h)), kkt=so.kkt, status=st, iters=so.iters,
                    polished=so.polished)

    def step_osqp(self, inp, with_var_rows=True, rho_interval=100, polish=True, warm=True, eps=0.0, max_iter=0):
        """The tick's QP as the REFERENCE's stack solves it (oracle_osqp.c): scaled variables, OSQP's ADMM at its default
        eps 1e-3, polish.  Returns X, U, dU, lam and info (iterations, solved, polish accepted, residuals)."""
        N, K = self.N, self.K
        si, keep = self._mk_in(inp)
        X = np.zeros((N, 6)); U = np.zeros((N - 1, 2)); dU = np.zeros((N - 1, 2)); lam = np.zeros(max(K, 1))
        so = StepOut()
        so.X, so.U, so.dU, so.lambda_ = _p(X), _p(U), _p(dU), _p(lam)
        info = np.zeros(8)
        st = self.L.orc_step_osqp(C.byref(self.veh), C.byref(self.cfg), self.ss, C.byref(si), C.byref(so), int(bool(with_var_rows)),
                                  int(rho_interval), int(bool(polish)), int(bool(warm)), float(eps), int(max_iter), _p(info))
        return dict(X=X, U=U, dU=dU, lam=lam, status=st, iters=int(info[0]), solved=info[1] == 0, polished=bool(info[2]),
                    pri_res=info[3], dua_res=info[4], pol_pri_res=info[5], pol_dua_res=info[6], rho=info[7])

    def step_sqp(self, inp, max_sqp_iter=20, tol=1e-9, impl="port"):
        """Full-dynamics variant (racing_mpc.cpp:67-84): SQP to convergence.  Returns the step() dict plus
        sqp_iters and defect (max nonlinear-dynamics violation of the returned trajectory)."""
        N, K = self.N, self.K
        si, keep = self._mk_in(inp)
        X = np.zeros((N, 6)); U = np.zeros((N - 1, 2)); dU = np.zeros((N - 1, 2))
        lam = np.zeros(max(K, 1)); ssx = np.zeros((max(K, 1), 6)); ssc = np.zeros(max(K, 1))
        so = StepOut()
        so.X, so.U, so.dU, so.lambda_, so.ss_x, so.ss_cost = _p(X), _p(U), _p(dU), _p(lam), _p(ssx), _p(ssc)
        its = C.c_int(0); dfc = C.c_double(0.0)
        st = self.L.orc_step_sqp(C.byref(self.veh), C.byref(self.cfg), self.ss, C.byref(si), C.byref(so),
                                 int(max_sqp_iter), float(tol), 0 if impl == "port" else 1, C.byref(its),
                                 C.cast(C.byref(dfc), C.POINTER(C.c_double)))
        return dict(X=X, U=U, dU=dU, lam=lam, cost=so.cost, status=st, iters=so.iters, sqp_iters=its.value,
                    defect=dfc.value, kkt=so.kkt)

    def check_candidate(self, inp, X, U, dU, lam=None):
        si, keep = self._mk_in(inp)
        X, U, dU = _f64(X), _f64(U), _f64(dU)
        lam = _f64(lam if lam is not None else np.zeros(max(self.K, 1)))
        cost = C.c_double(); inf = C.c_double()
        self.L.orc_check_candidate(C.byref(self.veh), C.byref(self.cfg), self.ss, C.byref(si), _p(X), _p(U),
                                   _p(dU), _p(lam), C.cast(C.byref(cost), C.POINTER(C.c_double)),
                                   C.cast(C.byref(inf), C.POINTER(C.c_double)))
        return cost.value, inf.value

    def step_batch(self, batch, impl="port", nthreads=1):
        """batch: dict of instance-major arrays (see racing_lmpc_ros2_b200.workload)."""
        B = batch["x_ic"].shape[0]
        N, K = self.N, self.K
        arrs = [_f64(batch[k]) for k in ("x_ic", "u_ic", "X_ref", "U_ref", "T_ref", "bound_left",
                                         "bound_right", "curvatures", "vel_ref", "total_length")]
        X = np.zeros((B, N, 6)); U = np.zeros((B, N - 1, 2)); dU = np.zeros((B, N - 1, 2))
        lam = np.zeros((B, max(K, 1))); cost = np.zeros(B); kkt = np.zeros(B)
        status = np.zeros(B, dtype=np.int32); iters = np.zeros(B, dtype=np.int32)
        nfail = self.L.orc_step_batch(
            C.byref(self.veh), C.byref(self.cfg), self.ss, B, *[_p(a) for a in arrs],
            _p(X), _p(U), _p(dU), _p(lam), _p(cost), status.ctypes.data_as(C.POINTER(C.c_int)),
            iters.ctypes.data_as(C.POINTER(C.c_int)), _p(kkt), 0 if impl == "port" else 1, int(nthreads))
        return dict(X=X, U=U, dU=dU, lam=lam, cost=cost, status=status, iters=iters, kkt=kkt, nfail=nfail)
