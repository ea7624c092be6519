"""ctypes mirror of include/lmpc_b200.h (PODs + function prototypes) and the library loader.

The CUDA library is built in-tree (csrc/liblmpc_b200.so, see build.py).  Loading it never touches
the GPU; every compute entry point needs one (lmpc_create returns LMPC_ERR_NO_DEVICE otherwise).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblmpc_b200.so")

LMPC_MEM_HOST, LMPC_MEM_DEVICE = 0, 1
STATUS_NAMES = {0: "SOLVED", 1: "MAX_ITER", 2: "INFEASIBLE_IC", 3: "NO_SAFE_SET", 4: "NUMERIC", 5: "SOLVED_INACCURATE", 6: "SQP_MAX_ITER"}


class VehicleParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "mass", "moi", "wheel_base", "cg_ratio", "cg_height", "fr", "chassis_b", "kd", "kb",
        "air_density", "frontal_area", "drag_coeff", "cl_f", "cl_r", "mu", "Bf", "Cf", "Br", "Cr",
        "Fd_max", "Fb_max", "Td", "Tb", "max_steer", "max_steer_rate")] + [
        ("integrator", C.c_int32), ("pad_", C.c_int32)]


class MpcConfig(C.Structure):
    _fields_ = [("N", C.c_int32), ("learning", C.c_int32), ("margin", C.c_double),
                ("q_contour", C.c_double), ("q_heading", C.c_double), ("q_vel", C.c_double),
                ("q_vy", C.c_double), ("q_vyaw", C.c_double), ("q_boundary", C.c_double),
                ("R", C.c_double * 4), ("R_d", C.c_double * 4),
                ("x_max", C.c_double * 6), ("x_min", C.c_double * 6),
                ("u_max", C.c_double * 2), ("u_min", C.c_double * 2),
                ("convex_hull_slack", C.c_double * 6),
                ("num_ss_pts", C.c_int32), ("num_ss_pts_per_lap", C.c_int32),
                ("max_lap_stored", C.c_int32), ("max_iter", C.c_int32), ("tol", C.c_double)]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)


class BatchIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "x_ic", "u_ic", "X_ref", "U_ref", "T_ref", "bound_left", "bound_right", "curvatures",
        "vel_ref", "total_length", "U_warm")]


class BatchOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "X_optm", "U_optm", "dU_optm", "convex_combi_optm", "ss_x", "ss_j", "cost", "status", "iters")]


class LoopOptions(C.Structure):
    _fields_ = [("step_mode", C.c_int32), ("delay_step", C.c_int32), ("plant_substeps", C.c_int32), ("pad_", C.c_int32),
                ("dt", C.c_double), ("plant_dt", C.c_double), ("speed_limit", C.c_double), ("speed_scale", C.c_double),
                ("max_vel_ref_diff", C.c_double)]


class RegSpec(C.Structure):
    _fields_ = [("n_out", C.c_int32), ("out_idx", C.c_int32 * 6), ("n_in_x", C.c_int32 * 6), ("in_x", (C.c_int32 * 6) * 6),
                ("n_in_u", C.c_int32 * 6), ("in_u", (C.c_int32 * 2) * 6), ("dist_max", C.c_double), ("ridge", C.c_double),
                ("sign", C.c_double)]


def make_reg_spec(out_idx, in_x, in_u, dist_max, ridge=1e-3, sign=1.0):
    """RegQuery's index lists (safe_set.hpp:61-76): out_idx[r] the output state of regression r, in_x[r] / in_u[r] its
    input states / controls."""
    sp = RegSpec()
    sp.n_out = len(out_idx)
    for r, o in enumerate(out_idx):
        sp.out_idx[r] = int(o)
        sp.n_in_x[r] = len(in_x[r]); sp.n_in_u[r] = len(in_u[r])
        for q, c in enumerate(in_x[r]):
            sp.in_x[r][q] = int(c)
        for q, c in enumerate(in_u[r]):
            sp.in_u[r][q] = int(c)
    sp.dist_max = float(dist_max); sp.ridge = float(ridge); sp.sign = float(sign)
    return sp


def fill_struct(st, d):
    for name, _ in st._fields_:
        if name not in d:
            continue
        cur = getattr(st, name)
        if hasattr(cur, "__len__"):
            arr = np.asarray(d[name], dtype=np.float64).ravel()
            for i in range(len(cur)):
                cur[i] = float(arr[i])
        else:
            setattr(st, name, d[name])
    return st


EXPORTS = [
    "lmpc_version", "lmpc_status_string", "lmpc_create", "lmpc_destroy", "lmpc_set_stream",
    "lmpc_last_error", "lmpc_launch_count", "lmpc_safe_set_add_lap", "lmpc_safe_set_load",
    "lmpc_safe_set_clear", "lmpc_safe_set_num_laps", "lmpc_safe_set_query_batch",
    "lmpc_discrete_dynamics_batch", "lmpc_linearise_batch", "lmpc_solve_batch", "lmpc_solve_sqp_batch", "lmpc_synchronize",
    "lmpc_set_timing", "lmpc_get_kernel_ms", "lmpc_measure_fp64_peak",
    "lmpc_track_set", "lmpc_track_load", "lmpc_track_total_length", "lmpc_track_eval_batch",
    "lmpc_frenet_to_global_batch", "lmpc_global_to_frenet_batch", "lmpc_closed_loop_run", "lmpc_prepare_batch",
    "lmpc_recorder_config", "lmpc_recorder_step", "lmpc_recorder_lap_count", "lmpc_safe_set_regress_batch", "lmpc_set_error_dynamics",
    "lmpc_to_base_control_batch", "lmpc_from_base_control_batch", "lmpc_model_create", "lmpc_safe_set_tick_count",
    "lmpc_agents_create", "lmpc_agents_destroy", "lmpc_closed_loop_run_agents", "lmpc_agents_get_lap", "lmpc_agents_status", "lmpc_agents_query_batch",
    "lmpc_gather_init", "lmpc_gather_connect", "lmpc_gather_buffer", "lmpc_solve_gather_batch", "lmpc_gather_wait", "lmpc_gather_error",
]
LMPC_IPC_HANDLE_BYTES = 64

_lib = None


def load_library(path=None):
    """dlopen the C-ABI library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("LMPC_B200_LIB") or LIB_PATH   # LMPC_B200_LIB: an alternate build of the same ABI
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: the CUDA library is not built (run `python -c 'import __graft_entry__ as g; "
            "g.build()'`).  There is no CPU fallback for the solve path.")
    L = C.CDLL(p)
    vp = C.c_void_p
    L.lmpc_version.restype = C.c_int
    L.lmpc_status_string.restype = C.c_char_p
    L.lmpc_status_string.argtypes = [C.c_int]
    L.lmpc_create.argtypes = [C.POINTER(MpcConfig), C.POINTER(VehicleParams), C.c_int, C.c_int, C.POINTER(vp)]
    L.lmpc_destroy.argtypes = [vp]
    L.lmpc_set_stream.argtypes = [vp, vp]
    L.lmpc_last_error.restype = C.c_char_p
    L.lmpc_last_error.argtypes = [vp]
    L.lmpc_launch_count.restype = C.c_int64
    L.lmpc_launch_count.argtypes = [vp]
    L.lmpc_safe_set_add_lap.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_double]
    L.lmpc_safe_set_load.argtypes = [vp, C.c_char_p, C.c_double]
    L.lmpc_safe_set_clear.argtypes = [vp]
    L.lmpc_safe_set_num_laps.argtypes = [vp]
    L.lmpc_safe_set_query_batch.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int]
    L.lmpc_discrete_dynamics_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int]
    L.lmpc_linearise_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    L.lmpc_solve_batch.argtypes = [vp, C.c_int, C.POINTER(BatchIn), C.POINTER(BatchOut), C.c_int]
    L.lmpc_solve_sqp_batch.argtypes = [vp, C.c_int, C.POINTER(BatchIn), C.POINTER(BatchOut), C.c_int, C.c_double, vp, vp, C.c_int]
    L.lmpc_synchronize.argtypes = [vp]
    L.lmpc_set_timing.argtypes = [vp, C.c_int]
    L.lmpc_get_kernel_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.lmpc_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.lmpc_track_set.argtypes = [vp, C.c_int, C.c_int, vp]
    L.lmpc_track_load.argtypes = [vp, C.c_char_p]
    L.lmpc_track_total_length.argtypes = [vp, C.POINTER(C.c_double)]
    L.lmpc_track_eval_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    L.lmpc_frenet_to_global_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    L.lmpc_global_to_frenet_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    L.lmpc_closed_loop_run.argtypes = [vp, C.c_int, C.c_int, C.POINTER(LoopOptions)] + [vp] * 8 + [C.c_int]
    L.lmpc_prepare_batch.argtypes = [vp, C.c_int, C.POINTER(LoopOptions)] + [vp] * 14
    L.lmpc_recorder_config.argtypes = [vp, C.c_int, C.c_char_p]
    L.lmpc_recorder_step.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_double, vp]
    L.lmpc_recorder_lap_count.argtypes = [vp]
    L.lmpc_safe_set_regress_batch.argtypes = [vp, C.c_int, C.POINTER(RegSpec), vp, vp, vp, vp, vp, vp, C.c_int]
    L.lmpc_set_error_dynamics.argtypes = [vp, C.POINTER(RegSpec)]
    L.lmpc_to_base_control_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    L.lmpc_from_base_control_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    L.lmpc_model_create.argtypes = [C.POINTER(VehicleParams), C.c_int, C.POINTER(vp)]
    L.lmpc_safe_set_tick_count.argtypes = [vp, C.POINTER(C.c_int32)]
    L.lmpc_agents_create.argtypes = [vp, C.c_int, C.c_int]
    L.lmpc_agents_destroy.argtypes = [vp]
    L.lmpc_closed_loop_run_agents.argtypes = [vp, C.c_int, C.c_int, C.POINTER(LoopOptions)] + [vp] * 9 + [C.c_double, C.c_int]
    L.lmpc_agents_get_lap.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), vp]
    L.lmpc_agents_status.argtypes = [vp, vp, vp, vp]
    L.lmpc_agents_query_batch.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.lmpc_gather_init.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_size_t)]
    L.lmpc_gather_connect.argtypes = [vp, vp]
    L.lmpc_gather_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.lmpc_solve_gather_batch.argtypes = [vp, C.c_int, C.POINTER(BatchIn), C.POINTER(BatchOut), C.c_int, C.c_int, vp, C.POINTER(C.c_uint64), C.c_int]
    L.lmpc_gather_wait.argtypes = [vp, C.c_uint64]
    L.lmpc_gather_error.argtypes = [vp, C.POINTER(C.c_int32)]
    if path is None:
        _lib = L
    return L
