"""Host-side mirror of the reference's RacingMPC / SafeSetManager / BaseVehicleModel surface for B
ticks at a time, on top of the C ABI (include/lmpc_b200.h).

    mpc = BatchedRacingMPC(vehicle_dict, config_dict, max_batch=1024)   # RacingMPC(config, model)
    mpc.add_lap(x, u, k, t, L)                                          # SafeSetManager::add_lap
    out = mpc.solve(batch)                                              # RacingMPC::solve, B ticks

`batch` uses the reference's input keys (racing_mpc.cpp:215-228); each value is instance-major
(X_ref: (B, N, 6) == B column-major 6 x N DMs).  numpy arrays take the HOST path (H2D / D2H inside the
call); torch CUDA tensors take the DEVICE path (zero copies, work enqueued on the stream passed to
set_stream(), else on torch's current stream).  Output keys: X_optm, U_optm, dU_optm, convex_combi_optm, ss_x, ss_j, cost, status, iters.
There is no CPU fallback: construction raises without the CUDA library and a GPU.
"""
import ctypes as C

import numpy as np

from . import binding as B

IN_KEYS = ("x_ic", "u_ic", "X_ref", "U_ref", "T_ref", "bound_left", "bound_right", "curvatures",
           "vel_ref", "total_length")


class LmpcError(RuntimeError):
    pass


def _check(lib, h, rc, what):
    if rc != 0:
        msg = lib.lmpc_status_string(rc).decode()
        detail = lib.lmpc_last_error(h).decode() if h else ""
        raise LmpcError(f"{what}: {msg} ({rc}) {detail}")


class BatchedRacingMPC:
    def __init__(self, vehicle, config, max_batch=1024, device=0, lib_path=None):
        self.lib = B.load_library(lib_path)
        self.vehicle = dict(vehicle)
        self.config = dict(config)
        self.N = int(config["N"])
        self.K = int(config["num_ss_pts"])
        self.learning = bool(config["learning"])
        self.max_batch = int(max_batch)
        self.device = int(device)
        self._veh = B.fill_struct(B.VehicleParams(), vehicle)
        self._cfg = B.fill_struct(B.MpcConfig(), config)
        self._h = C.c_void_p()
        rc = self.lib.lmpc_create(C.byref(self._cfg), C.byref(self._veh), self.device, self.max_batch,
                                  C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            _check(self.lib, None, rc, "lmpc_create")
        self._solved = False
        self._stream_explicit = False     # set_stream() was called: the device path then leaves the handle's stream alone
        self._pending_status = None       # status tensor of the last device-path solve (read lazily by solved())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.lmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def set_stream(self, stream=None):
        """Enqueue on `stream` (a torch.cuda.Stream, a raw cudaStream_t int, or None = default)."""
        ptr = 0
        if stream is not None:
            ptr = int(getattr(stream, "cuda_stream", stream))
        self._stream_explicit = stream is not None
        _check(self.lib, self._h, self.lib.lmpc_set_stream(self._h, C.c_void_p(ptr)), "lmpc_set_stream")

    def synchronize(self):
        _check(self.lib, self._h, self.lib.lmpc_synchronize(self._h), "lmpc_synchronize")

    def set_timing(self, enable=True):
        _check(self.lib, self._h, self.lib.lmpc_set_timing(self._h, 1 if enable else 0), "lmpc_set_timing")

    def kernel_ms(self):
        """(ms_linearise, ms_ss_query, ms_qp) summed over the solves recorded since set_timing, and their count."""
        ms = (C.c_double * 3)()
        n = C.c_int()
        _check(self.lib, self._h, self.lib.lmpc_get_kernel_ms(self._h, ms, C.byref(n)), "lmpc_get_kernel_ms")
        return (ms[0], ms[1], ms[2]), n.value

    def measure_fp64_peak(self):
        """Sustained DFMA rate of this device in TFLOP/s (denominator of the fp64-pipe fraction)."""
        v = C.c_double()
        _check(self.lib, self._h, self.lib.lmpc_measure_fp64_peak(self._h, C.byref(v)), "lmpc_measure_fp64_peak")
        return v.value

    @property
    def launch_count(self):
        return int(self.lib.lmpc_launch_count(self._h))

    def solved(self):
        """RacingMPC::solved(): latches true after the first successful solve (status SOLVED or SOLVED_INACCURATE of
        at least one instance).  After a device-path solve this reads that solve's status tensor (synchronises)."""
        if not self._solved and self._pending_status is not None:
            st = self._pending_status.cpu().numpy()
            self._pending_status = None
            if ((st == 0) | (st == 5)).any():
                self._solved = True
        return self._solved

    # ------------------------------------------------------------------ safe set
    def add_lap(self, x, u, k, t, total_length):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        k = np.ascontiguousarray(k, dtype=np.float64).ravel()
        t = np.ascontiguousarray(t, dtype=np.float64).ravel()
        assert x.ndim == 2 and x.shape[1] == 6
        rc = self.lib.lmpc_safe_set_add_lap(self._h, x.shape[0], x.ctypes.data, u.ctypes.data, k.ctypes.data,
                                            t.ctypes.data, float(total_length))
        _check(self.lib, self._h, rc, "lmpc_safe_set_add_lap")

    def load_lap(self, prefix, total_length):
        _check(self.lib, self._h, self.lib.lmpc_safe_set_load(self._h, str(prefix).encode(), float(total_length)),
               "lmpc_safe_set_load")

    def clear_safe_set(self):
        _check(self.lib, self._h, self.lib.lmpc_safe_set_clear(self._h), "lmpc_safe_set_clear")

    def num_laps(self):
        return int(self.lib.lmpc_safe_set_num_laps(self._h))

    def ss_query(self, queries, max_total=None, per_lap=None):
        """SafeSetManager::query for B points (B, 2) -> ss_x (B, count, 6), ss_j (B, count)."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 2)
        Bn = q.shape[0]
        mt = int(self.K if max_total is None else max_total)
        pl = int(self.config["num_ss_pts_per_lap"] if per_lap is None else per_lap)
        sx = np.zeros((Bn, mt, 6)); sj = np.zeros((Bn, mt)); cnt = np.zeros(Bn, dtype=np.int32)
        rc = self.lib.lmpc_safe_set_query_batch(self._h, Bn, q.ctypes.data, mt, pl, sx.ctypes.data, sj.ctypes.data,
                                                cnt.ctypes.data, B.LMPC_MEM_HOST)
        _check(self.lib, self._h, rc, "lmpc_safe_set_query_batch")
        c = int(cnt[0]) if Bn else 0
        return sx[:, :c], sj[:, :c]

    # ------------------------------------------------------------------ lap recorder (SafeSetRecorder)
    def recorder_config(self, to_file=False, prefix=""):
        _check(self.lib, self._h, self.lib.lmpc_recorder_config(self._h, int(bool(to_file)), str(prefix).encode()), "lmpc_recorder_config")

    def recorder_step(self, x, u, k, t, total_length):
        """SafeSetRecorder::step: one tick's (x_ic, u_ic, curvatures[0], t_ic); True when a completed lap was added."""
        x = np.ascontiguousarray(x, dtype=np.float64).ravel(); u = np.ascontiguousarray(u, dtype=np.float64).ravel()
        added = C.c_int32(0)
        rc = self.lib.lmpc_recorder_step(self._h, x.ctypes.data, u.ctypes.data, float(k), float(t), float(total_length), C.addressof(added))
        _check(self.lib, self._h, rc, "lmpc_recorder_step")
        return bool(added.value)

    def recorder_lap_count(self):
        return int(self.lib.lmpc_recorder_lap_count(self._h))

    # ------------------------------------------------------------------ error-dynamics regression (RegQuery)
    def regress(self, spec, xq, uq, A, Bm, Cv):
        """SafeSetManager::query(RegQuery) for n items: returns the corrected (A (n,6,6), B (n,6,2), C (n,6)) and the number
        of samples within dist_max per (item, regression).  A / B are row-indexed here ([item, row, col])."""
        xq = np.ascontiguousarray(xq, dtype=np.float64).reshape(-1, 6); uq = np.ascontiguousarray(uq, dtype=np.float64).reshape(-1, 2)
        n = xq.shape[0]
        Ac = np.ascontiguousarray(np.asarray(A, dtype=np.float64).reshape(n, 6, 6).transpose(0, 2, 1))   # column-major per item
        Bc = np.ascontiguousarray(np.asarray(Bm, dtype=np.float64).reshape(n, 6, 2).transpose(0, 2, 1))
        Cc = np.array(Cv, dtype=np.float64).reshape(n, 6).copy()
        npts = np.zeros((n, spec.n_out), dtype=np.int32)
        rc = self.lib.lmpc_safe_set_regress_batch(self._h, n, C.byref(spec), xq.ctypes.data, uq.ctypes.data, Ac.ctypes.data,
                                                  Bc.ctypes.data, Cc.ctypes.data, npts.ctypes.data, B.LMPC_MEM_HOST)
        _check(self.lib, self._h, rc, "lmpc_safe_set_regress_batch")
        return Ac.transpose(0, 2, 1).copy(), Bc.transpose(0, 2, 1).copy(), Cc, npts

    def set_error_dynamics(self, spec=None):
        """Apply the regression to every stage's (A, B, g) inside solve / solve_sqp / closed_loop (None disables)."""
        rc = self.lib.lmpc_set_error_dynamics(self._h, C.byref(spec) if spec is not None else None)
        _check(self.lib, self._h, rc, "lmpc_set_error_dynamics")

    # ------------------------------------------------------------------ track (RacingTrajectory)
    def set_track(self, table):
        """table: (n, >= 13) trajectory-file rows in TrajectoryIndex column order (racing_trajectory.hpp:37-56)."""
        tb = np.ascontiguousarray(table, dtype=np.float64)
        _check(self.lib, self._h, self.lib.lmpc_track_set(self._h, tb.shape[0], tb.shape[1], tb.ctypes.data), "lmpc_track_set")

    def load_track(self, path):
        _check(self.lib, self._h, self.lib.lmpc_track_load(self._h, str(path).encode()), "lmpc_track_load")

    def track_length(self):
        v = C.c_double()
        _check(self.lib, self._h, self.lib.lmpc_track_total_length(self._h, C.byref(v)), "lmpc_track_total_length")
        return v.value

    def track_eval(self, s):
        """The interpolation functions at abscissae s -> dict of left, right, curvature, vel, x, y, yaw."""
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        out = np.zeros((len(s), 7))
        _check(self.lib, self._h, self.lib.lmpc_track_eval_batch(self._h, len(s), s.ctypes.data, out.ctypes.data, B.LMPC_MEM_HOST),
               "lmpc_track_eval_batch")
        return dict(left=out[:, 0], right=out[:, 1], curvature=out[:, 2], vel=out[:, 3], x=out[:, 4], y=out[:, 5], yaw=out[:, 6])

    def frenet_to_global(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1, 3)
        g = np.zeros_like(f)
        _check(self.lib, self._h, self.lib.lmpc_frenet_to_global_batch(self._h, len(f), f.ctypes.data, g.ctypes.data, B.LMPC_MEM_HOST),
               "lmpc_frenet_to_global_batch")
        return g

    def global_to_frenet(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64).reshape(-1, 3)
        f = np.zeros_like(g)
        _check(self.lib, self._h, self.lib.lmpc_global_to_frenet_batch(self._h, len(g), g.ctypes.data, f.ctypes.data, B.LMPC_MEM_HOST),
               "lmpc_global_to_frenet_batch")
        return f

    # ------------------------------------------------------------------ closed loop
    @staticmethod
    def loop_options(dt, step_mode="step", delay_step=0, plant_dt=None, plant_substeps=1, speed_limit=1e9, speed_scale=1.0,
                     max_vel_ref_diff=1.0):
        o = B.LoopOptions()
        o.step_mode = {"step": 0, "continuous": 1}[step_mode]
        o.delay_step = int(delay_step); o.plant_substeps = int(plant_substeps)
        o.dt = float(dt); o.plant_dt = float(dt if plant_dt is None else plant_dt)
        o.speed_limit = float(speed_limit); o.speed_scale = float(speed_scale); o.max_vel_ref_diff = float(max_vel_ref_diff)
        return o

    def prepare(self, opt, x, u_prev, X_last, U_last):
        """RacingMPCNode's input preparation (racing_mpc_node.cpp:236-292) for B agents -> the solve's input dict
        (torch CUDA tensors in, torch CUDA tensors out)."""
        import torch
        Bn, N = int(x.shape[0]), self.N
        dev = x.device
        shapes = dict(x_ic=(Bn, 6), u_ic=(Bn, 2), X_ref=(Bn, N, 6), U_ref=(Bn, N - 1, 2), T_ref=(Bn, N - 1), bound_left=(Bn, N),
                      bound_right=(Bn, N), curvatures=(Bn, N), vel_ref=(Bn, N), total_length=(Bn,))
        out = {k: torch.empty(shp, dtype=torch.float64, device=dev) for k, shp in shapes.items()}
        rc = self.lib.lmpc_prepare_batch(self._h, Bn, C.byref(opt), x.data_ptr(), u_prev.data_ptr(), X_last.data_ptr(),
                                         U_last.data_ptr(), *[out[k].data_ptr() for k in IN_KEYS])
        _check(self.lib, self._h, rc, "lmpc_prepare_batch")
        return out

    def closed_loop(self, opt, ticks, x, u_prev, X_last, U_last, lap_count=None, log=True):
        """`ticks` MPC ticks of B agents entirely on the device (prepare -> solve -> plant).  numpy in / numpy out:
        returns dict(x, u_prev, X_last, U_last, lap_count, fail_count[, log_x (ticks, B, 6), log_u (ticks, B, 2)])."""
        x = np.array(x, dtype=np.float64, order="C"); u_prev = np.array(u_prev, dtype=np.float64, order="C")
        X_last = np.array(X_last, dtype=np.float64, order="C"); U_last = np.array(U_last, dtype=np.float64, order="C")
        Bn = x.shape[0]
        laps = np.zeros(Bn, dtype=np.int32) if lap_count is None else np.array(lap_count, dtype=np.int32)
        fails = np.zeros(Bn, dtype=np.int32)
        lx = np.zeros((ticks, Bn, 6)) if log else None
        lu = np.zeros((ticks, Bn, 2)) if log else None
        rc = self.lib.lmpc_closed_loop_run(self._h, Bn, int(ticks), C.byref(opt), x.ctypes.data, u_prev.ctypes.data,
                                           X_last.ctypes.data, U_last.ctypes.data, laps.ctypes.data, fails.ctypes.data,
                                           lx.ctypes.data if log else None, lu.ctypes.data if log else None, B.LMPC_MEM_HOST)
        _check(self.lib, self._h, rc, "lmpc_closed_loop_run")
        out = dict(x=x, u_prev=u_prev, X_last=X_last, U_last=U_last, lap_count=laps, fail_count=fails)
        if log:
            out["log_x"] = lx; out["log_u"] = lu
        return out

    # ------------------------------------------------------------------ per-agent safe sets (device-side learning)
    def agents_create(self, Bn, max_lap_samples=1024):
        """Give each of Bn agents its own safe set (seeded with the current laps) and lap recorder on the device."""
        _check(self.lib, self._h, self.lib.lmpc_agents_create(self._h, int(Bn), int(max_lap_samples)), "lmpc_agents_create")
        self._agents = (int(Bn), int(max_lap_samples))

    def agents_destroy(self):
        _check(self.lib, self._h, self.lib.lmpc_agents_destroy(self._h), "lmpc_agents_destroy")

    def closed_loop_agents(self, opt, ticks, x, u_prev, X_last, U_last, lap_count=None, t0=0.0, log=True):
        """closed_loop() in which every agent records its own laps and queries its own safe set (lmpc_closed_loop_run_agents).
        Adds log_rec (ticks, B, 10): what each agent's recorder was fed (x_ic, u_ic, curvatures[0], t_ic)."""
        x = np.array(x, dtype=np.float64, order="C"); u_prev = np.array(u_prev, dtype=np.float64, order="C")
        X_last = np.array(X_last, dtype=np.float64, order="C"); U_last = np.array(U_last, dtype=np.float64, order="C")
        Bn = x.shape[0]
        laps = np.zeros(Bn, dtype=np.int32) if lap_count is None else np.array(lap_count, dtype=np.int32)
        fails = np.zeros(Bn, dtype=np.int32)
        lx = np.zeros((ticks, Bn, 6)) if log else None
        lu = np.zeros((ticks, Bn, 2)) if log else None
        lr = np.zeros((ticks, Bn, 10)) if log else None
        rc = self.lib.lmpc_closed_loop_run_agents(self._h, Bn, int(ticks), C.byref(opt), x.ctypes.data, u_prev.ctypes.data,
                                                  X_last.ctypes.data, U_last.ctypes.data, laps.ctypes.data, fails.ctypes.data,
                                                  lx.ctypes.data if log else None, lu.ctypes.data if log else None,
                                                  lr.ctypes.data if log else None, float(t0), B.LMPC_MEM_HOST)
        _check(self.lib, self._h, rc, "lmpc_closed_loop_run_agents")
        out = dict(x=x, u_prev=u_prev, X_last=X_last, U_last=U_last, lap_count=laps, fail_count=fails)
        if log:
            out["log_x"] = lx; out["log_u"] = lu; out["log_rec"] = lr
        return out

    def agents_status(self):
        Bn = self._agents[0]
        lc = np.zeros(Bn, dtype=np.int32); st = np.zeros(Bn, dtype=np.int32); fl = np.zeros(Bn, dtype=np.int32)
        _check(self.lib, self._h, self.lib.lmpc_agents_status(self._h, lc.ctypes.data, st.ctypes.data, fl.ctypes.data), "lmpc_agents_status")
        return dict(lap_count=lc, stored=st, flags=fl)

    def agents_get_lap(self, agent, which=0):
        """The agent's which-th newest stored lap (0 = newest) as (n, 6) states, or None."""
        n = C.c_int32()
        _check(self.lib, self._h, self.lib.lmpc_agents_get_lap(self._h, int(agent), int(which), 0, C.byref(n), None), "lmpc_agents_get_lap")
        if n.value == 0:
            return None
        x = np.zeros((n.value, 6))
        _check(self.lib, self._h, self.lib.lmpc_agents_get_lap(self._h, int(agent), int(which), n.value, C.byref(n), x.ctypes.data), "lmpc_agents_get_lap")
        return x

    def agents_query(self, queries, max_total=None, per_lap=None):
        """SafeSetManager::query of every agent on its own laps: (B, 2) -> list of (ss_x (count_b, 6), ss_j (count_b,))."""
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 2)
        Bn = q.shape[0]
        mt = int(self.K if max_total is None else max_total)
        pl = int(self.config["num_ss_pts_per_lap"] if per_lap is None else per_lap)
        sx = np.zeros((Bn, mt, 6)); sj = np.zeros((Bn, mt)); cnt = np.zeros(Bn, dtype=np.int32)
        rc = self.lib.lmpc_agents_query_batch(self._h, q.ctypes.data, mt, pl, sx.ctypes.data, sj.ctypes.data, cnt.ctypes.data)
        _check(self.lib, self._h, rc, "lmpc_agents_query_batch")
        return [(sx[b, :cnt[b]], sj[b, :cnt[b]]) for b in range(Bn)]

    # ------------------------------------------------------------------ model
    def discrete_dynamics(self, x, u, kappa, dt):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 6)
        n = x.shape[0]
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(n, 2)
        kappa = np.ascontiguousarray(np.broadcast_to(kappa, (n,)), dtype=np.float64)
        dt = np.ascontiguousarray(np.broadcast_to(dt, (n,)), dtype=np.float64)
        xn = np.zeros((n, 6))
        rc = self.lib.lmpc_discrete_dynamics_batch(self._h, n, x.ctypes.data, u.ctypes.data, kappa.ctypes.data,
                                                   dt.ctypes.data, xn.ctypes.data, B.LMPC_MEM_HOST)
        _check(self.lib, self._h, rc, "lmpc_discrete_dynamics_batch")
        return xn

    def to_base_control(self, u):
        """BaseVehicleModel::to_base_control (single_track_planar_model.cpp:390-400): (n, 2) -> (n, 3) = (Fd, Fb, delta)."""
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 2)
        ub = np.zeros((u.shape[0], 3))
        _check(self.lib, self._h, self.lib.lmpc_to_base_control_batch(self._h, u.shape[0], u.ctypes.data, ub.ctypes.data, B.LMPC_MEM_HOST), "lmpc_to_base_control_batch")
        return ub

    def from_base_control(self, ub):
        """BaseVehicleModel::from_base_control (:401-407): (n, 3) -> (n, 2)."""
        ub = np.ascontiguousarray(ub, dtype=np.float64).reshape(-1, 3)
        u = np.zeros((ub.shape[0], 2))
        _check(self.lib, self._h, self.lib.lmpc_from_base_control_batch(self._h, ub.shape[0], ub.ctypes.data, u.ctypes.data, B.LMPC_MEM_HOST), "lmpc_from_base_control_batch")
        return u

    def ss_tick_count(self):
        c = C.c_int32()
        _check(self.lib, self._h, self.lib.lmpc_safe_set_tick_count(self._h, C.byref(c)), "lmpc_safe_set_tick_count")
        return c.value

    def linearise(self, x, u, kappa, dt):
        """discrete_dynamics_jacobian: returns A (n,6,6), B (n,6,2), g (n,6), x_next (n,6)."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 6)
        n = x.shape[0]
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(n, 2)
        kappa = np.ascontiguousarray(np.broadcast_to(kappa, (n,)), dtype=np.float64)
        dt = np.ascontiguousarray(np.broadcast_to(dt, (n,)), dtype=np.float64)
        A = np.zeros((n, 36)); Bm = np.zeros((n, 12)); g = np.zeros((n, 6)); xn = np.zeros((n, 6))
        rc = self.lib.lmpc_linearise_batch(self._h, n, x.ctypes.data, u.ctypes.data, kappa.ctypes.data, dt.ctypes.data,
                                           A.ctypes.data, Bm.ctypes.data, g.ctypes.data, xn.ctypes.data, B.LMPC_MEM_HOST)
        _check(self.lib, self._h, rc, "lmpc_linearise_batch")
        return A.reshape(n, 6, 6).transpose(0, 2, 1).copy(), Bm.reshape(n, 2, 6).transpose(0, 2, 1).copy(), g, xn

    # ------------------------------------------------------------------ solve
    def _shapes(self, Bn):
        N, K = self.N, max(self.K, 1)
        return dict(X_optm=(Bn, N, 6), U_optm=(Bn, N - 1, 2), dU_optm=(Bn, N - 1, 2), convex_combi_optm=(Bn, K),
                    ss_x=(Bn, K, 6), ss_j=(Bn, K), cost=(Bn,))

    def _arena(self, nbytes, pinned):
        if pinned:
            import torch
            return torch.empty(nbytes, dtype=torch.uint8).pin_memory().numpy()
        return np.zeros(nbytes, dtype=np.uint8)

    def alloc_host_outputs(self, Bn, pinned=False):
        """Output buffers for the host path (optionally CUDA-pinned through torch), carved from ONE arena in the order of
        lmpc_batch_out so that the library returns them with two copies (trajectories + multipliers, cost + status +
        iterations) plus the safe-set columns on the side stream."""
        shp = self._shapes(Bn)
        order = ("X_optm", "U_optm", "dU_optm", "convex_combi_optm", "ss_x", "ss_j", "cost")
        nb = sum(int(np.prod(shp[k])) * 8 for k in order) + 8 * Bn
        arena = self._arena(nb, pinned)
        out, o = {}, 0
        for k in order:
            n = int(np.prod(shp[k])) * 8
            out[k] = arena[o:o + n].view(np.float64).reshape(shp[k])
            o += n
        out["status"] = arena[o:o + 4 * Bn].view(np.int32); o += 4 * Bn
        out["iters"] = arena[o:o + 4 * Bn].view(np.int32)
        return out

    def alloc_host_inputs(self, batch, pinned=False):
        """Copy a batch (dict of numpy arrays, the reference's input keys) into ONE arena in the order of lmpc_batch_in:
        the host path then uploads it with a single copy."""
        keys = list(IN_KEYS) + (["U_optm_ref"] if "U_optm_ref" in batch else [])
        arrs = {k: np.ascontiguousarray(batch[k], dtype=np.float64) for k in keys}
        arena = self._arena(sum(a.nbytes for a in arrs.values()), pinned)
        out, o = {}, 0
        for k in keys:
            a = arrs[k]
            out[k] = arena[o:o + a.nbytes].view(np.float64).reshape(a.shape)
            out[k][...] = a
            o += a.nbytes
        return out

    def solve(self, batch, out=None):
        first = batch["x_ic"]
        if isinstance(first, np.ndarray):
            return self._solve_host(batch, out)
        return self._solve_device(batch, out)

    def _solve_host(self, batch, out=None):
        Bn = int(np.asarray(batch["x_ic"]).shape[0])
        keep = {k: np.ascontiguousarray(batch[k], dtype=np.float64) for k in IN_KEYS}
        warm = batch.get("U_optm_ref", None)
        if warm is not None:
            keep["U_warm"] = np.ascontiguousarray(warm, dtype=np.float64)
        bi = B.BatchIn()
        for k in IN_KEYS:
            setattr(bi, k, keep[k].ctypes.data)
        bi.U_warm = keep["U_warm"].ctypes.data if "U_warm" in keep else None
        if out is None:
            out = self.alloc_host_outputs(Bn)
        bo = B.BatchOut()
        for k in ("X_optm", "U_optm", "dU_optm", "convex_combi_optm", "ss_x", "ss_j", "cost", "status", "iters"):
            setattr(bo, k, out[k].ctypes.data)
        rc = self.lib.lmpc_solve_batch(self._h, Bn, C.byref(bi), C.byref(bo), B.LMPC_MEM_HOST)
        _check(self.lib, self._h, rc, "lmpc_solve_batch")
        if ((out["status"] == 0) | (out["status"] == 5)).any():
            self._solved = True
        return out

    def solve_sqp(self, batch, max_sqp_iter=100, tol=1e-9, out=None):
        """RacingMPC(config, model, full_dynamics=True).solve (racing_mpc.cpp:67-84,162-166): the problem with the
        nonlinear dynamics constraint, solved by SQP on the tick's kernels (host buffers).  Adds `sqp_iters` (QP solves
        per instance) and `defect` (max nonlinear-dynamics violation of the returned trajectory) to the outputs.  status 0
        means the SQP's step test passed; an instance still moving after max_sqp_iter passes gets status 6 (SQP_MAX_ITER)."""
        Bn = int(np.asarray(batch["x_ic"]).shape[0])
        keep = {k: np.ascontiguousarray(batch[k], dtype=np.float64) for k in IN_KEYS}
        warm = batch.get("U_optm_ref", None)
        if warm is not None:
            keep["U_warm"] = np.ascontiguousarray(warm, dtype=np.float64)
        bi = B.BatchIn()
        for k in IN_KEYS:
            setattr(bi, k, keep[k].ctypes.data)
        bi.U_warm = keep["U_warm"].ctypes.data if "U_warm" in keep else None
        if out is None:
            out = self.alloc_host_outputs(Bn)
        out["sqp_iters"] = np.zeros(Bn, dtype=np.int32)
        out["defect"] = np.zeros(Bn)
        bo = B.BatchOut()
        for k in ("X_optm", "U_optm", "dU_optm", "convex_combi_optm", "ss_x", "ss_j", "cost", "status", "iters"):
            setattr(bo, k, out[k].ctypes.data)
        rc = self.lib.lmpc_solve_sqp_batch(self._h, Bn, C.byref(bi), C.byref(bo), int(max_sqp_iter), float(tol),
                                           out["sqp_iters"].ctypes.data, out["defect"].ctypes.data, B.LMPC_MEM_HOST)
        _check(self.lib, self._h, rc, "lmpc_solve_sqp_batch")
        if (out["status"] == 0).any():
            self._solved = True
        return out

    def alloc_device_outputs(self, Bn, device=None):
        """Device output buffers.  X_optm, U_optm, dU_optm, cost and status are views into ONE allocation, out["slab"]
        (array-major: [X | U | dU | cost | status(int32)]), so that a multi-GPU caller gathers the trajectories of
        all ranks with a single collective on that buffer and no packing kernel (distributed.unpack_flat_slab)."""
        import torch
        dev = device or torch.device("cuda", self.device)
        shp = self._shapes(Bn)
        out = {}
        gathered = ("X_optm", "U_optm", "dU_optm", "cost")
        n64 = sum(int(np.prod(shp[k])) for k in gathered)
        slab = torch.zeros(n64 + (Bn + 1) // 2, dtype=torch.float64, device=dev)
        o = 0
        for k in gathered:
            n = int(np.prod(shp[k]))
            out[k] = slab[o:o + n].view(shp[k])
            o += n
        out["status"] = slab[o:].view(torch.int32)[:Bn]
        out["slab"] = slab
        for k, sh in shp.items():
            if k not in out:
                out[k] = torch.empty(sh, dtype=torch.float64, device=dev)
        out["iters"] = torch.empty(Bn, dtype=torch.int32, device=dev)
        return out

    # ------------------------------------------------------------------ multi-GPU result exchange (peer memory)
    def gather_init(self, world, rank, Bn, sets=2):
        """Allocates the gather buffer ([sets][world][slab]) and returns (ipc_handle bytes, slab length in doubles)."""
        buf = (C.c_ubyte * B.LMPC_IPC_HANDLE_BYTES)()
        sb = C.c_size_t()
        rc = self.lib.lmpc_gather_init(self._h, int(world), int(rank), int(Bn), int(sets), buf, C.byref(sb))
        _check(self.lib, self._h, rc, "lmpc_gather_init")
        self._gather = dict(world=int(world), rank=int(rank), B=int(Bn), sets=int(sets), slab_doubles=sb.value // 8)
        return bytes(buf), sb.value // 8

    def gather_connect(self, handles):
        """handles: the `world` IPC handles (bytes, rank order) every rank obtained from gather_init."""
        blob = b"".join(handles)
        assert len(blob) == B.LMPC_IPC_HANDLE_BYTES * self._gather["world"]
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        _check(self.lib, self._h, self.lib.lmpc_gather_connect(self._h, arr), "lmpc_gather_connect")

    def gather_buffer_ptr(self, s):
        p = C.c_void_p(); n = C.c_size_t()
        _check(self.lib, self._h, self.lib.lmpc_gather_buffer(self._h, int(s), C.byref(p), C.byref(n)), "lmpc_gather_buffer")
        return p.value, n.value

    def gather_wait(self, seq):
        _check(self.lib, self._h, self.lib.lmpc_gather_wait(self._h, int(seq)), "lmpc_gather_wait")

    def gather_error(self):
        e = C.c_int32()
        _check(self.lib, self._h, self.lib.lmpc_gather_error(self._h, C.byref(e)), "lmpc_gather_error")
        return e.value

    def solve_gather(self, batch, out, s, wait=True, gathered_host=None):
        """lmpc_solve_gather_batch: like solve(), with X_optm / U_optm / dU_optm / cost / status produced in this rank's
        block of gathered set `s` and mirrored into every peer's buffer by the QP kernel.  Device tensors or numpy
        (host path; gathered_host: optional numpy array of world * slab doubles receiving the whole set).  Returns the
        sequence number of the exchange (for gather_wait when wait=False)."""
        first = batch["x_ic"]
        host = isinstance(first, np.ndarray)
        ptr = (lambda a: a.ctypes.data) if host else (lambda a: a.data_ptr())
        if not host and not self._stream_explicit:
            import torch
            cur = torch.cuda.current_stream(first.device).cuda_stream
            _check(self.lib, self._h, self.lib.lmpc_set_stream(self._h, C.c_void_p(int(cur))), "lmpc_set_stream")
        bi = B.BatchIn()
        for k in IN_KEYS:
            setattr(bi, k, ptr(batch[k]))
        warm = batch.get("U_optm_ref", None)
        bi.U_warm = ptr(warm) if warm is not None else None
        bo = B.BatchOut()
        for k in ("convex_combi_optm", "ss_x", "ss_j", "iters"):
            setattr(bo, k, ptr(out[k]))
        if host and gathered_host is None:
            for k in ("X_optm", "U_optm", "dU_optm", "cost", "status"):
                setattr(bo, k, ptr(out[k]))
        seq = C.c_uint64()
        rc = self.lib.lmpc_solve_gather_batch(self._h, int(first.shape[0]), C.byref(bi), C.byref(bo), int(s), 1 if wait else 0,
                                              gathered_host.ctypes.data if gathered_host is not None else None, C.byref(seq),
                                              B.LMPC_MEM_HOST if host else B.LMPC_MEM_DEVICE)
        _check(self.lib, self._h, rc, "lmpc_solve_gather_batch")
        return seq.value

    def _solve_device(self, batch, out=None):
        """torch CUDA tensors in, torch CUDA tensors out; asynchronous.  Work is enqueued on the stream given to
        set_stream(); when none was given, on torch's CURRENT stream of the tensors' device (so inputs produced on a
        non-default torch stream are ordered before the kernels)."""
        import torch
        Bn = int(batch["x_ic"].shape[0])
        if not self._stream_explicit:
            cur = torch.cuda.current_stream(batch["x_ic"].device).cuda_stream
            _check(self.lib, self._h, self.lib.lmpc_set_stream(self._h, C.c_void_p(int(cur))), "lmpc_set_stream")
        for k in IN_KEYS:
            t = batch[k]
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
                raise LmpcError(f"device path needs contiguous float64 CUDA tensors ({k})")
        if out is None:
            out = self.alloc_device_outputs(Bn, batch["x_ic"].device)
        bi = B.BatchIn()
        for k in IN_KEYS:
            setattr(bi, k, batch[k].data_ptr())
        warm = batch.get("U_optm_ref", None)
        bi.U_warm = warm.data_ptr() if warm is not None else None
        bo = B.BatchOut()
        for k in ("X_optm", "U_optm", "dU_optm", "convex_combi_optm", "ss_x", "ss_j", "cost", "status", "iters"):
            setattr(bo, k, out[k].data_ptr())
        rc = self.lib.lmpc_solve_batch(self._h, Bn, C.byref(bi), C.byref(bo), B.LMPC_MEM_DEVICE)
        _check(self.lib, self._h, rc, "lmpc_solve_batch")
        self._pending_status = out["status"]    # not latched here: the solve has not run yet (see solved())
        return out
