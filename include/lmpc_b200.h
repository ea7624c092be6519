/*
 * lmpc_b200.h -- C ABI of the B200-native batched LMPC solve (liblmpc_b200.so).
 *
 * Drop-in boundary for the per-tick solve of MPC-Berkeley/Racing-LMPC-ROS2.  Each entry
 * point names the reference interface it replaces (paths relative to the reference's src/):
 *
 *   lmpc_create / lmpc_destroy          RacingMPC::RacingMPC(config, model, full_dynamics=false)
 *                                       mpc/racing_mpc/include/racing_mpc/racing_mpc.hpp:46-49,
 *                                       vehicle_model_factory.cpp:31-50 ("single_track_planar_model")
 *   lmpc_solve_batch                    RacingMPC::solve(in, out, stats)   racing_mpc.hpp:52,
 *                                       racing_mpc.cpp:209-372  (B independent ticks per call)
 *   lmpc_solve_sqp_batch                RacingMPC(config, model, full_dynamics=true)::solve -- the one-off IPOPT
 *                                       solve of the nonlinear-dynamics problem, racing_mpc.cpp:67-84,162-166,
 *                                       racing_mpc_node.cpp:299-314
 *   lmpc_linearise_batch                BaseVehicleModel::discrete_dynamics_jacobian() {x,u,k,dt}->{A,B,g}
 *                                       single_track_planar_model.cpp:377-387
 *   lmpc_discrete_dynamics_batch        BaseVehicleModel::discrete_dynamics() {x,u,k,dt}->{xip1}
 *                                       single_track_planar_model.cpp:370-375
 *   lmpc_safe_set_add_lap               SafeSetManager::add_lap(x,u,k,t,total_length)   safe_set.hpp:119-121
 *   lmpc_safe_set_load                  SafeSetRecorder::load(from_files,total_length)  safe_set.cpp:260-276
 *   lmpc_safe_set_query_batch           SafeSetManager::query(const SSQuery&) -> SSResult  safe_set.cpp:153-180
 *   lmpc_safe_set_regress_batch         SafeSetManager::query(const RegQuery&) -> RegResult  safe_set.cpp:56-114,182-245
 *   lmpc_set_error_dynamics             (the same regression applied to every stage's A, B, g inside the tick)
 *   lmpc_recorder_step / _config        SafeSetRecorder::step / SafeSetRecorder(manager, to_file, prefix)
 *                                       safe_set.cpp:246-258,278-322 (called from RacingMPC::solve, racing_mpc.cpp:245-246)
 *   lmpc_track_set / _load / _eval_batch RacingTrajectory and its interpolation functions  racing_trajectory.cpp:25-119
 *   lmpc_frenet_to_global_batch,        RacingTrajectory::frenet_to_global / global_to_frenet
 *   lmpc_global_to_frenet_batch           racing_trajectory.cpp:121-236
 *   lmpc_prepare_batch                  RacingMPCNode::on_step_timer's input preparation  racing_mpc_node.cpp:236-292
 *   lmpc_closed_loop_run                prepare + solve + RacingSimulator::step, B agents  racing_simulator.cpp:97-113
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; functions return an lmpc_status code
 *     and never throw across the ABI.
 *   - every array is fp64, instance-major: element [b] of a batch is the reference's dense
 *     column-major casadi::DM for that tick (X_ref[b] = 6 x N column-major = N rows of 6).
 *   - `memspace` says where the caller's buffers live.  LMPC_MEM_HOST buffers are staged through
 *     device memory owned by the handle (H2D, kernels, D2H, then a stream synchronise).  Pinned host buffers make the
 *     copies asynchronous; when the outputs of lmpc_solve_batch are one pinned arena in the order of lmpc_batch_out
 *     (X_optm, U_optm, dU_optm, convex_combi_optm, ss_x, ss_j, cost, status, iters back to back) the QP kernel stores
 *     the results straight into it (zero-copy) and no D2H copy of them follows;
 *     LMPC_MEM_DEVICE buffers are used in place, work is enqueued on the handle's stream and the
 *     call returns without synchronising.
 *   - a handle is thread-compatible, not thread-safe (same contract as RacingMPC::solve, which
 *     the node serialises under a mutex, racing_mpc_node.cpp:158).
 *   - there is NO CPU fallback: without a CUDA device lmpc_create returns LMPC_ERR_NO_DEVICE.
 */
#ifndef LMPC_B200_H_
#define LMPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMPC_NX 6
#define LMPC_NU 2
#define LMPC_MAX_N 128        /* horizon cap of the kernels (N is a runtime parameter below it; the shipped files go to n: 80,
                               * iac_car_tracking_mpc.param.yaml:7).  lmpc_create also rejects a horizon whose working set
                               * exceeds the 227 KB of shared memory a CTA may use. */
#define LMPC_MAX_SS_PTS 128   /* num_ss_pts cap */

/* Per-call / per-handle error codes. */
enum lmpc_status {
  LMPC_OK = 0,
  LMPC_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
  LMPC_ERR_NO_DEVICE = -2,    /* no CUDA device or driver */
  LMPC_ERR_CUDA = -3,         /* CUDA runtime error (see lmpc_last_error) */
  LMPC_ERR_ALLOC = -4,
  LMPC_ERR_IO = -5,           /* safe-set file missing / malformed */
  LMPC_ERR_CAPACITY = -6      /* batch larger than max_batch */
};

/* Per-instance solve status (lmpc_batch_out.status).  The reference signals failure by omitting
 * "X_optm" from the output dict (racing_mpc.cpp:358-371); here a failed instance gets a code and
 * never poisons the batch. */
enum lmpc_instance_status {
  LMPC_SOLVED = 0,
  LMPC_MAX_ITER = 1,
  LMPC_INFEASIBLE_IC = 2,     /* x_ic violates the hard box on x_0 (racing_mpc.cpp:147,199-201) */
  LMPC_NO_SAFE_SET = 3,       /* learning mode with an empty safe set */
  LMPC_NUMERIC = 4,           /* factorisation broke down before convergence */
  LMPC_SOLVED_INACCURATE = 5, /* the interior point converged (tol, then 1e-2 tol) but the active-set polish could not
                               * certify an active set: the trajectory is the interior-point iterate (typically 1e-7..1e-4
                               * from the optimum).  OSQP's "solved inaccurate" / "polish unsuccessful".  Outputs are
                               * written; the adapter publishes them (they are far inside the reference's own eps 1e-3). */
  LMPC_SQP_MAX_ITER = 6       /* lmpc_solve_sqp_batch only: the last QP solved but the SQP iteration hit max_sqp_iter
                               * before its step test passed -- the trajectory violates the nonlinear dynamics (IPOPT's
                               * "Maximum_Iterations_Exceeded" under error_on_fail, racing_mpc.cpp:71) */
};

enum lmpc_memspace { LMPC_MEM_HOST = 0, LMPC_MEM_DEVICE = 1 };

/* SingleTrackPlanarModel constants (BaseVehicleModelConfig + SingleTrackPlanarModelConfig,
 * base_vehicle_model_config.hpp:30-153, single_track_planar_model.hpp:31-43).  Only
 * simplify_lon_control=true / use_frenet=true is supported (what every launch file ships). */
typedef struct lmpc_vehicle_params {
  double mass, moi, wheel_base, cg_ratio, cg_height, fr, chassis_b;
  double kd;                 /* powertrain.kd */
  double kb;                 /* front_brake.bias */
  double air_density, frontal_area, drag_coeff, cl_f, cl_r;
  double mu;
  double Bf, Cf, Br, Cr;     /* pacejka_b / pacejka_c front, rear */
  double Fd_max, Fb_max, Td, Tb;
  double max_steer, max_steer_rate;
  int32_t integrator;        /* 0 = rk4, 1 = euler (base_vehicle_model IntegratorType) */
  int32_t pad_;
} lmpc_vehicle_params;

/* RacingMPCConfig (racing_mpc_config.hpp:37-82) -- the fields the QP reads, plus the
 * interior-point options of this implementation (max_iter, tol). */
typedef struct lmpc_mpc_config {
  int32_t N;                 /* horizon (states x_0..x_{N-1}) */
  int32_t learning;          /* 1 = LMPC cost + safe-set terminal set, 0 = tracking cost */
  double margin;
  double q_contour, q_heading, q_vel, q_vy, q_vyaw, q_boundary;
  double R[4], R_d[4];
  double x_max[6], x_min[6], u_max[2], u_min[2];   /* +-INFINITY (or |v| >= 1e19) = unbounded */
  double convex_hull_slack[6];
  int32_t num_ss_pts, num_ss_pts_per_lap, max_lap_stored;
  int32_t max_iter;          /* interior-point iteration cap (default 30 when <= 0) */
  double tol;                /* interior-point stage tolerance: primal step per channel relative to
                              * max(1, |channel|) (default 1e-7 when <= 0; complementarity floor
                              * 1e-4 * tol).  The active-set polish that follows (the counterpart of
                              * OSQP's polish=true) then lands on the optimum to ~1e-12. */
} lmpc_mpc_config;

typedef struct lmpc_handle lmpc_handle;

/* Inputs of B ticks: the keys RacingMPC::solve reads (racing_mpc.cpp:215-228). */
typedef struct lmpc_batch_in {
  const double* x_ic;          /* [B][6] */
  const double* u_ic;          /* [B][2] */
  const double* X_ref;         /* [B][N][6]   linearisation states (abscissa aligned inside) */
  const double* U_ref;         /* [B][N-1][2] linearisation controls */
  const double* T_ref;         /* [B][N-1]    "T_ref"/"T_optm_ref" */
  const double* bound_left;    /* [B][N] */
  const double* bound_right;   /* [B][N] */
  const double* curvatures;    /* [B][N] */
  const double* vel_ref;       /* [B][N] */
  const double* total_length;  /* [B] */
  const double* U_warm;        /* [B][N-1][2] optional "U_optm_ref" start; NULL => U_ref */
} lmpc_batch_in;

/* Outputs: the keys RacingMPC::solve writes (racing_mpc.cpp:256-257,347-352) plus cost/status.
 * Any pointer may be NULL to skip that output. */
typedef struct lmpc_batch_out {
  double* X_optm;              /* [B][N][6] */
  double* U_optm;              /* [B][N-1][2] */
  double* dU_optm;             /* [B][N-1][2] */
  double* convex_combi_optm;   /* [B][num_ss_pts]   (learning) */
  double* ss_x;                /* [B][num_ss_pts][6] safe-set columns used by the QP (padded) */
  double* ss_j;                /* [B][num_ss_pts]    their raw cost-to-go J (what the reference's out["ss_j"] holds,
                                * racing_mpc.cpp:257; the shift J - J[0] of :280 happens inside the QP kernel) */
  double* cost;                /* [B] objective value (the reference never returns it) */
  int32_t* status;             /* [B] lmpc_instance_status */
  int32_t* iters;              /* [B] interior-point iterations (stats["iter_count"]) */
} lmpc_batch_out;

int lmpc_version(void);
const char* lmpc_status_string(int status);

int lmpc_create(const lmpc_mpc_config* config, const lmpc_vehicle_params* vehicle, int device_ordinal,
                int max_batch, lmpc_handle** out);
int lmpc_destroy(lmpc_handle* h);
/* cudaStream_t to enqueue on (NULL = the legacy default stream). */
int lmpc_set_stream(lmpc_handle* h, void* cuda_stream);
const char* lmpc_last_error(const lmpc_handle* h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches evidence) */
int64_t lmpc_launch_count(const lmpc_handle* h);

/* ---- safe set (host-side ingestion, device-resident slab) ---- */
int lmpc_safe_set_add_lap(lmpc_handle* h, int n, const double* x /*[n][6]*/, const double* u /*[n][2]*/,
                          const double* k /*[n]*/, const double* t /*[n]*/, double total_length);
int lmpc_safe_set_load(lmpc_handle* h, const char* file_prefix, double total_length);
int lmpc_safe_set_clear(lmpc_handle* h);
int lmpc_safe_set_num_laps(const lmpc_handle* h);
/* k-NN query of B points (s, e_y).  ss_x [B][max_total][6], ss_j [B][max_total] (raw J), count [B].
 * Columns beyond count[b] are left untouched. */
int lmpc_safe_set_query_batch(lmpc_handle* h, int B, const double* query_s_ey /*[B][2]*/, int max_total,
                              int max_per_lap, double* ss_x, double* ss_j, int32_t* count, int memspace);

/* ---- lap recorder (SafeSetRecorder, safe_set.cpp:246-258,278-322): host-side segmentation of a stream of ticks into
 *      laps.  RacingMPC::solve feeds it (x_ic, u_ic, curvatures[0], t_ic) every tick (racing_mpc.cpp:245-246).  A lap
 *      ends when the abscissa drops by more than half the track length (:290); the samples before the first wrap are
 *      discarded (:293-309); a completed lap goes to lmpc_safe_set_add_lap and, with to_file, to
 *      <prefix>lap_<k>_{x,u,t,k}.txt ("%.16e", one sample per line: what lmpc_safe_set_load reads back). ---- */
int lmpc_recorder_config(lmpc_handle* h, int to_file, const char* file_prefix);
/* lap_added (optional): set to 1 when this sample closed a lap that was added to the safe set */
int lmpc_recorder_step(lmpc_handle* h, const double* x /*[6]*/, const double* u /*[2]*/, double k, double t,
                       double total_length, int32_t* lap_added);
int lmpc_recorder_lap_count(const lmpc_handle* h);   /* SafeSetRecorder::lap_count_ */

/* ---- error-dynamics regression over the stored laps (RegQuery / RegResult, safe_set.hpp:61-88;
 *      SSTrajectory::query(RegQuery) safe_set.cpp:56-114; SafeSetManager::query(RegQuery) :182-245).
 *      One regression per output state: local weighted ridge regression of the one-step model error on
 *      (x[in_x], u[in_u], 1) over the stored samples within dist_max of the query, added to the nominal A, B, C.
 *      The reference never calls this function and it cannot run as written (csrc/lmpc_reg_core.cuh lists the
 *      deviations); `sign` = -1 reproduces b = -M'Ky as written (:229), +1 adds the error model (LMPC paper). ---- */
#define LMPC_REG_MAX_OUT 6
typedef struct lmpc_reg_spec {
  int32_t n_out;                           /* number of regressions = reg_out_state_idxs.size() */
  int32_t out_idx[LMPC_REG_MAX_OUT];       /* reg_out_state_idxs[r][0] (exactly one per regression, :63-65) */
  int32_t n_in_x[LMPC_REG_MAX_OUT];        /* reg_in_state_idxs[r] */
  int32_t in_x[LMPC_REG_MAX_OUT][LMPC_NX];
  int32_t n_in_u[LMPC_REG_MAX_OUT];        /* reg_in_control_idxs[r] */
  int32_t in_u[LMPC_REG_MAX_OUT][LMPC_NU];
  double dist_max;                         /* RegQuery::dist_max = the kernel bandwidth h */
  double ridge;                            /* 1e-3 (:228) */
  double sign;                             /* see above */
} lmpc_reg_spec;
/* n query items: xq [n][6], uq [n][2]; A [n][36] / B [n][12] (column-major per item) and C [n][6] hold the nominal
 * model on entry and the corrected one on return.  npts (optional) [n][n_out]: samples within dist_max. */
int lmpc_safe_set_regress_batch(lmpc_handle* h, int n, const lmpc_reg_spec* spec, const double* xq, const double* uq,
                                double* A, double* Bm, double* C, int32_t* npts, int memspace);
/* Enables (spec != NULL) or disables (NULL) the regression inside lmpc_solve_batch / lmpc_solve_sqp_batch /
 * lmpc_closed_loop_run: after the linearisation every stage's (A, B, g) is corrected at its linearisation point
 * (abscissa-aligned X_ref_i, U_ref_i); g takes the affine term.  One more kernel per tick. */
int lmpc_set_error_dynamics(lmpc_handle* h, const lmpc_reg_spec* spec);

/* ---- vehicle model ---- */
int lmpc_discrete_dynamics_batch(lmpc_handle* h, int n, const double* x, const double* u,
                                 const double* kappa, const double* dt, double* x_next, int memspace);
/* A [n][36] and B [n][12] column-major per item, g [n][6]; x_next may be NULL. */
int lmpc_linearise_batch(lmpc_handle* h, int n, const double* x, const double* u, const double* kappa,
                         const double* dt, double* A, double* Bm, double* g, double* x_next, int memspace);

/* BaseVehicleModel::to_base_control / from_base_control (single_track_planar_model.cpp:390-417, simplify_lon_control):
 * derived control (u_lon, delta) [n][2] <-> base control (Fd, Fb, delta) [n][3].  to_base: Fd = u_lon / (1 + e^-u_lon),
 * Fb = u_lon / (1 + e^u_lon) (as written: no x1000, unlike the dynamics' tanh split, :214-217); from_base: the
 * larger-magnitude one of (Fd, Fb).  to_base_state / from_base_state are the identity for this model (:411-414). */
int lmpc_to_base_control_batch(lmpc_handle* h, int n, const double* u, double* u_base, int memspace);
int lmpc_from_base_control_batch(lmpc_handle* h, int n, const double* u_base, double* u, int memspace);
/* A handle that serves the model functions only (vehicle_model_factory::load_vehicle_model("single_track_planar_model"),
 * vehicle_model_factory.cpp:31-50: the node builds the model before any MPC configuration exists). */
int lmpc_model_create(const lmpc_vehicle_params* vehicle, int device_ordinal, lmpc_handle** out);
/* Columns the tick's safe-set query finds before padding = out["ss_x"].size2() of the reference (racing_mpc.cpp:249-262). */
int lmpc_safe_set_tick_count(lmpc_handle* h, int32_t* count);

/* ---- the hot path: B independent MPC ticks ---- */
int lmpc_solve_batch(lmpc_handle* h, int B, const lmpc_batch_in* in, const lmpc_batch_out* out,
                     int memspace);
/* Full-dynamics variant (RacingMPC(..., full_dynamics = true), racing_mpc.cpp:67-84,162-166): the same cost and rows
 * with the NONLINEAR discrete dynamics x_{i+1} = f_d(x_i, u_i, curvatures_i, T_i) as equality constraints, which the
 * reference hands to IPOPT once per run (racing_mpc_node.cpp:299-314).  Here: sequential quadratic programming on the
 * tick's own kernels -- linearise at the current trajectory, solve the QP, take the full step, repeat until the step
 * max |new - old| / max(1, |new|) over X and U is below sqp_tol (or max_sqp_iter passes).  A fixed point of this
 * iteration satisfies the KKT conditions of the nonlinear problem.  The safe-set columns are queried once, at the
 * caller's X_ref[:, N-1] (racing_mpc.cpp:249-255), exactly as the reference's solve() does before calling IPOPT.
 * sqp_iters [B] (optional): QP solves spent per instance.  defect [B] (optional): max |x_{i+1} - f_d(x_i, u_i)| of the
 * returned trajectory (the nonlinear constraint violation).  out->status: the status of the last QP of the instance when
 * that QP failed; LMPC_SOLVED (/ _INACCURATE) when the step test passed; LMPC_SQP_MAX_ITER when the instance was still
 * moving after max_sqp_iter passes (IPOPT runs with max_iter 1000 and error_on_fail: racing_mpc.cpp:67-84; 100 passes
 * cover 1023 of 1024 random BARC tracking ticks, mean 14).  The step length of the iteration follows the secant rule of
 * lmpc_sqp_update_kernel (csrc/lmpc_kernels.cuh); an l1-merit backtracking line search was measured against it and
 * rejected (mean 36 passes, 16 % at the cap: DESIGN.md). */
int lmpc_solve_sqp_batch(lmpc_handle* h, int B, const lmpc_batch_in* in, const lmpc_batch_out* out, int max_sqp_iter,
                         double sqp_tol, int32_t* sqp_iters, double* defect, int memspace);
/* ---- track (RacingTrajectory: the degree-3 "bspline" interpolants over the abscissa and what is built from them,
 *      vehicle_dynamics_models/racing_trajectory/src/racing_trajectory.cpp:25-236) ----
 * table: the trajectory file's rows (>= 13 columns in TrajectoryIndex order, racing_trajectory.hpp:37-56), row-major. */
int lmpc_track_set(lmpc_handle* h, int n_rows, int n_cols, const double* table);
int lmpc_track_load(lmpc_handle* h, const char* file);             /* RacingTrajectory(file_name), :188-191 */
int lmpc_track_total_length(const lmpc_handle* h, double* total_length);
/* left_boundary / right_boundary / curvature / velocity / x / y / yaw interpolation functions (:96-118) at n abscissae:
 * out [n][7] in that order. */
int lmpc_track_eval_batch(lmpc_handle* h, int n, const double* s, double* out, int memspace);
/* frenet_to_global (:121-136, 193-202) and global_to_frenet (:138-186, 204-236) for n poses [n][3] = (s, t, xi) / (x, y, phi). */
int lmpc_frenet_to_global_batch(lmpc_handle* h, int n, const double* frenet, double* global, int memspace);
int lmpc_global_to_frenet_batch(lmpc_handle* h, int n, const double* global, double* frenet, int memspace);

/* ---- closed loop on the device: tick preparation + solve + plant, many agents, no host round trip ----
 * What RacingMPCNode::on_step_timer does around the solve (mpc/racing_mpc/src/racing_mpc_node.cpp:236-292, 322-331,
 * 386-401) and what RacingSimulator / RacingSimulatorNode do with the published actuation
 * (simulation/racing_simulator/src/racing_simulator.cpp:97-113, racing_simulator_node.cpp:241-286). */
typedef struct lmpc_loop_options {
  int32_t step_mode;        /* 0 = step: x_ic is the measured state; 1 = continuous: x_ic = f_d(measured, last_u[0])
                             * (RacingMPCStepMode, racing_mpc_node.cpp:238-244) */
  int32_t delay_step;       /* column of the solution that is published (racing_mpc_node.cpp:386-389) */
  int32_t plant_substeps;   /* simulator steps per MPC tick */
  int32_t pad_;
  double dt;                /* MPC sample time: T_ref and the shift-and-extend step (racing_mpc_node.cpp:65,248) */
  double plant_dt;          /* racing_simulator dt (RK4 step of the plant) */
  double speed_limit, speed_scale, max_vel_ref_diff;   /* velocity-reference clipping, racing_mpc_node.cpp:266-287 */
} lmpc_loop_options;
/* Runs `ticks` MPC ticks for B agents.  Agent state, all in/out: x [B][6] plant state (Frenet), u_prev [B][2] the
 * last published (u_a, u_steer) (= u_ic of the next tick), X_last [B][N][6] / U_last [B][N-1][2] the previous solution
 * (last_x_ / last_u_; start them from lmpc_solve_sqp_batch as the node does).  Optional: lap_count [B] (in/out; the
 * simulator node's counter), fail_count [B] (out: ticks whose solve failed -- the shifted reference is kept, as in the
 * node), log_x [ticks][B][6] / log_u [ticks][B][2] (plant state after, and actuation of, every tick). */
int lmpc_closed_loop_run(lmpc_handle* h, int B, int ticks, const lmpc_loop_options* opt, double* x, double* u_prev,
                         double* X_last, double* U_last, int32_t* lap_count, int32_t* fail_count, double* log_x,
                         double* log_u, int memspace);
/* ---- per-agent safe sets and lap recording on the device (Monte-Carlo LMPC in which every agent learns from its OWN
 *      laps).  In the reference each RacingMPC owns a SafeSetManager and a SafeSetRecorder (racing_mpc.hpp:99-100) fed by
 *      every solve (racing_mpc.cpp:245-246, safe_set.cpp:278-322).  lmpc_agents_create gives each of B agents its own
 *      circular buffer of max_lap_stored laps (capacity max_lap_samples each), seeded with the handle's current laps, and a
 *      recorder; lmpc_closed_loop_run_agents is lmpc_closed_loop_run in which every tick feeds the agents' recorders with
 *      (x_ic, u_ic, curvatures(0), t_ic = t0 + tick dt), stores completed laps (process_lap_data, safe_set.cpp:116-137) and
 *      queries each agent's own laps -- all on the device.  log_rec (optional) [ticks][B][10]: what each recorder was fed.
 *      The shared-set entry points (lmpc_solve_batch, lmpc_closed_loop_run with another B) keep using the handle's laps. ---- */
int lmpc_agents_create(lmpc_handle* h, int B, int max_lap_samples);
int lmpc_agents_destroy(lmpc_handle* h);
int lmpc_closed_loop_run_agents(lmpc_handle* h, int B, int ticks, const lmpc_loop_options* opt, double* x, double* u_prev,
                                double* X_last, double* U_last, int32_t* lap_count, int32_t* fail_count, double* log_x,
                                double* log_u, double* log_rec, double t0, int memspace);
/* the agent's which-th newest stored lap (0 = newest): n_out samples, x [n][6] (host buffer of capacity max_n; may be NULL) */
int lmpc_agents_get_lap(lmpc_handle* h, int agent, int which, int max_n, int32_t* n_out, double* x);
/* host arrays [B]: SafeSetRecorder::lap_count_, laps stored, flags (bit 3: a lap overflowed max_lap_samples) */
int lmpc_agents_status(lmpc_handle* h, int32_t* lap_count, int32_t* stored, int32_t* flags);
/* SafeSetManager::query for every agent on its own laps (host buffers): query [B][2] -> ss_x [B][max_total][6], ss_j, count [B] */
int lmpc_agents_query_batch(lmpc_handle* h, const double* query, int max_total, int max_per_lap, double* ss_x, double* ss_j,
                            int32_t* count);

/* The preparation step on its own (DEVICE buffers only): fills the input keys of lmpc_solve_batch. */
int lmpc_prepare_batch(lmpc_handle* h, int B, const lmpc_loop_options* opt, const double* x, const double* u_prev,
                       const double* X_last, const double* U_last, double* x_ic, double* u_ic, double* X_ref, double* U_ref,
                       double* T_ref, double* bound_left, double* bound_right, double* curvatures, double* vel_ref,
                       double* total_length);

/* ---- multi-GPU: sharded batch, every rank receives every rank's converged trajectories (the reference
 *      is a single process and has no counterpart).  One process per GPU of one NVSwitch domain.  The exchange is fused
 *      into the QP kernel: its epilogue stores each instance's X, U, dU, cost, status into every peer's gather buffer
 *      through NVLink peer mappings and publishes a sequence number; receivers wait on local memory.
 *      Set-up: every rank calls lmpc_gather_init (same world, B, sets), the ranks exchange the 64-byte IPC handles by
 *      any means (torch.distributed.all_gather in distributed.py), every rank calls lmpc_gather_connect.
 *      Gathered set layout: [world][slab], slab = [X [B][N][6] | U [B][N-1][2] | dU [B][N-1][2] | cost [B] |
 *      status int32 [B]] padded to a multiple of 256 bytes (slab_bytes). ---- */
#define LMPC_IPC_HANDLE_BYTES 64
int lmpc_gather_init(lmpc_handle* h, int world, int rank, int B, int sets, void* ipc_handle_out /*[64]*/, size_t* slab_bytes);
int lmpc_gather_connect(lmpc_handle* h, const void* ipc_handles_all /*[world][64]*/);
/* device pointer of gathered set `set` ([world][slab]) and the slab length in doubles */
int lmpc_gather_buffer(lmpc_handle* h, int set, double** device_ptr, size_t* doubles_per_rank);
/* lmpc_solve_batch whose trajectory outputs (X_optm, U_optm, dU_optm, cost, status) are produced in this rank's block
 * of gathered set `set` and mirrored to every peer.  wait != 0: also enqueue the wait for this exchange (afterwards the
 * whole set is valid on this rank, stream-ordered); wait == 0: call lmpc_gather_wait(seq) later (e.g. after enqueuing the
 * next solve into another set).  Two sets make back-to-back solves safe (a set is overwritten two exchanges later).
 * DEVICE memspace: out->X_optm/U_optm/dU_optm/cost/status are ignored (may be NULL).  HOST memspace: they receive this
 * rank's results unless gathered_host is given (with wait != 0), which receives the whole set [world][slab] instead. */
int lmpc_solve_gather_batch(lmpc_handle* h, int B, const lmpc_batch_in* in, const lmpc_batch_out* out, int set, int wait,
                            double* gathered_host, uint64_t* seq_out, int memspace);
int lmpc_gather_wait(lmpc_handle* h, uint64_t seq);
/* synchronises; *peer_timed_out = 0, or 1 + the rank a wait gave up on (after about two seconds) */
int lmpc_gather_error(lmpc_handle* h, int32_t* peer_timed_out);

/* Per-kernel device timing (measurement aid): when enabled, CUDA events are recorded on the handle's
 * stream around the three kernels of every lmpc_solve_batch; lmpc_get_kernel_ms synchronises and
 * returns the summed milliseconds {linearise, safe-set query, QP} over the recorded solves. */
int lmpc_set_timing(lmpc_handle* h, int enable);
int lmpc_get_kernel_ms(lmpc_handle* h, double* ms3, int* nsolves);
/* Measurement aid for the fp64 roofline: runs a register-resident chain of independent DFMAs on every
 * SM of the handle's device and returns the sustained rate in TFLOP/s (2 flop per DFMA lane).  This is the
 * denominator of the fp64-pipe fraction bench.py reports; MEASURED_PEAKS.json carries no fp64 figure. */
int lmpc_measure_fp64_peak(lmpc_handle* h, double* tflops);
/* Blocks until everything enqueued on the handle's stream has finished. */
int lmpc_synchronize(lmpc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* LMPC_B200_H_ */
