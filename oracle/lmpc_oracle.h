/*
 * lmpc_oracle.h -- CPU ORACLE for the batched LMPC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C fp64 restatement of the reference's per-tick solve
 * (MPC-Berkeley/Racing-LMPC-ROS2, RacingMPC::solve and everything it evaluates).
 * Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in CasADi/OSQP/CGAL, none of
 * which exist in this environment, and the reference's own tests assert no numbers on
 * this path.  The oracle is instead validated by
 *   - sympy symbolic differentiation and central differences (Jacobians),
 *   - two independent QP methods (dense primal-dual IPM, then active-set polish) plus a
 *     KKT certificate computed from the dense QP data,
 *   - the reference's recorded BARC laps as a loose dynamics fixture.
 *
 * Reference map (paths relative to /root/reference/src):
 *   dynamics f(x,u,k)        vehicle_dynamics_models/single_track_planar_model/src/single_track_planar_model.cpp:195-342
 *   RK4 / Euler              tools/lmpc_utils/src/utils.cpp:88-123
 *   A,B,g linearisation      single_track_planar_model.cpp:377-387
 *   actuator rows            single_track_planar_model.cpp:53-159
 *   align_abscissa           tools/lmpc_utils/include/lmpc_utils/utils.hpp:35-41
 *   QP statement             mpc/racing_mpc/src/racing_mpc.cpp:31-202, 442-543
 *   per-tick data flow       mpc/racing_mpc/src/racing_mpc.cpp:209-372
 *   safe set                 vehicle_dynamics_models/racing_trajectory/src/safe_set.cpp:33-54,116-180,260-276
 *   k-NN                     vehicle_dynamics_models/racing_trajectory/src/trajectory_kd_tree.cpp:27-63
 */
#ifndef LMPC_ORACLE_H_
#define LMPC_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NX 6
#define ORC_NU 2

/* SingleTrackPlanarModel + BaseVehicleModelConfig scalars that the path reads
 * (single_track_planar_model.cpp:224-251, 65-72). */
typedef struct orc_vehicle {
  double mass;            /* chassis.total_mass */
  double moi;             /* chassis.moi (Jzz) */
  double wheel_base;      /* chassis.wheel_base (l) */
  double cg_ratio;        /* chassis.cg_ratio ; lr = cg_ratio*l, lf = l-lr */
  double cg_height;       /* chassis.cg_height */
  double fr;              /* chassis.fr rolling resistance */
  double chassis_b;       /* chassis.b (width used in the boundary margin) */
  double kd;              /* powertrain.kd */
  double kb;              /* front_brake.bias */
  double air_density, frontal_area, drag_coeff, cl_f, cl_r;
  double mu;              /* single_track_planar.mu */
  double Bf, Cf, Br, Cr;  /* pacejka b,c front / rear */
  double Fd_max, Fb_max, Td, Tb;
  double max_steer, max_steer_rate;
  int integrator;         /* 0 = rk4, 1 = euler */
  int pad_;
} orc_vehicle;

/* RacingMPCConfig (racing_mpc_config.hpp:37-82) -- fields the QP reads. */
typedef struct orc_config {
  int N;
  int learning;
  double margin;
  double q_contour, q_heading, q_vel, q_vy, q_vyaw, q_boundary;
  double R[4], R_d[4];              /* 2x2, row/col symmetric */
  double x_max[6], x_min[6], u_max[2], u_min[2];
  double convex_hull_slack[6];
  int num_ss_pts, num_ss_pts_per_lap, max_lap_stored;
  int max_iter;                     /* solver iteration cap (ours, not OSQP's) */
  double tol;                       /* solver tolerance (ours) */
} orc_config;

/* ---- model (oracle_model.c) ------------------------------------------------ */
void orc_dynamics(const orc_vehicle* v, const double x[6], const double u[2], double kappa,
                  double xdot[6]);
void orc_discrete_dynamics(const orc_vehicle* v, const double x[6], const double u[2],
                           double kappa, double dt, double xnext[6]);
/* A (6x6 col-major), B (6x2 col-major), g (6): single_track_planar_model.cpp:377-379 */
void orc_linearise(const orc_vehicle* v, const double x[6], const double u[2], double kappa,
                   double dt, double A[36], double B[12], double g[6], double xnext[6]);
double orc_align_abscissa(double s1, double s2, double total);

/* ---- safe set (oracle_safeset.c) ------------------------------------------- */
typedef struct orc_safe_set orc_safe_set;
orc_safe_set* orc_ss_create(int max_lap_stored);
void orc_ss_destroy(orc_safe_set* ss);
/* x is n rows of 6 (the layout of the *_x.txt files == column-major 6 x n DM). */
int orc_ss_add_lap(orc_safe_set* ss, int n, const double* x, const double* u, const double* k,
                   const double* t, double total_length);
int orc_ss_load(orc_safe_set* ss, const char* prefix, double total_length);
int orc_ss_num_laps(const orc_safe_set* ss);
/* SafeSetManager::query(SSQuery) (safe_set.cpp:153-180).  Returns the number of columns
 * found (<= max_total); ss_x is [count][6], ss_j is [count]. */
int orc_ss_query(const orc_safe_set* ss, double qs, double qey, int max_total, int max_per_lap,
                 double* ss_x, double* ss_j);
/* query + pad/truncate to exactly K columns + J - J[0] (racing_mpc.cpp:263-281).
 * Returns the raw count (0 => no safe set). */
int orc_ss_query_padded(const orc_safe_set* ss, double qs, double qey, int K, int per_lap,
                        double* ss_x, double* ss_cost);

/* ---- one MPC tick (oracle_qp_dense.c / oracle_port.c) ---------------------- */
typedef struct orc_step_in {
  const double* x_ic;        /* 6 */
  const double* u_ic;        /* 2 */
  const double* X_ref;       /* 6 x N col-major */
  const double* U_ref;       /* 2 x (N-1) */
  const double* T_ref;       /* N-1 */
  const double* bound_left;  /* N */
  const double* bound_right; /* N */
  const double* curvatures;  /* N */
  const double* vel_ref;     /* N */
  double total_length;
  const double* ss_query_point; /* optional (s, e_y): safe-set query point; NULL => aligned X_ref(:, N-1) (racing_mpc.cpp:249-255) */
} orc_step_in;

typedef struct orc_step_out {
  double* X;        /* 6 x N */
  double* U;        /* 2 x (N-1) */
  double* dU;       /* 2 x (N-1) */
  double* lambda;   /* K (learning) or NULL */
  double* ss_x;     /* 6 x K used by the QP (may be NULL) */
  double* ss_cost;  /* K (J - J0) (may be NULL) */
  double cost;
  double sigma_b;
  double sigma_h[6];
  double kkt;       /* max KKT residual of the returned point (dense certificate) */
  int status;       /* 0 = solved */
  int iters;
  int polished;     /* dense path: 1 if active-set polish accepted */
} orc_step_out;

enum { ORC_OK = 0, ORC_MAX_ITER = 1, ORC_INFEASIBLE_IC = 2, ORC_NO_SAFE_SET = 3, ORC_NUMERIC = 4, ORC_INACCURATE = 5, ORC_SQP_MAX_ITER = 6 };

/* Dense statement of the reference QP, dense Mehrotra IPM, active-set polish, KKT check. */
int orc_step_dense(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss,
                   const orc_step_in* in, orc_step_out* out);
/* Structure-exploiting (Riccati) IPM: the CPU "port" used as the timed CPU baseline. */
int orc_step_port(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss,
                  const orc_step_in* in, orc_step_out* out);
/* Batch driver (OpenMP over instances) over packed arrays, instance-major; impl 0=port 1=dense.
 * Returns number of instances with status != 0. */
/* The reference's own solver stack restated (oracle_osqp.c): the QP in the reference's scaled variables, OSQP's published
 * ADMM with its defaults (eps 1e-3) and polish.  info[8] = {iterations, 0 solved / 1 max_iter, polish accepted, ADMM primal
 * residual, ADMM dual residual, polished primal residual, polished dual residual, final rho}. */
int orc_step_osqp(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss, const orc_step_in* in, orc_step_out* out,
                  int with_var_rows, int rho_interval, int do_polish, int warm, double eps, int max_iter, double* info);
int orc_step_batch(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss, int B,
                   const double* x_ic, const double* u_ic, const double* X_ref,
                   const double* U_ref, const double* T_ref, const double* bl, const double* br,
                   const double* kap, const double* vref, const double* total_length,
                   double* X, double* U, double* dU, double* lambda, double* cost, int* status,
                   int* iters, double* kkt, int impl, int nthreads);
/* Full-dynamics variant (RacingMPC(..., full_dynamics=true), racing_mpc.cpp:67-84,162-166): the reference gives the
 * problem with the nonlinear dynamics constraint to IPOPT once per run (racing_mpc_node.cpp:299-314).  Restated here as
 * full-step SQP: solve the tick's QP (impl 0 = port, 1 = dense), re-linearise at its solution, repeat until
 * max |new - old| / max(1, |new|) over X and U < tol.  The safe set is queried once at the caller's X_ref(:, N-1).
 * Returns the status of the last QP; *sqp_iters = QP solves, *defect = max |x_{i+1} - f_d(x_i, u_i, k_i, T_i)|. */
int orc_step_sqp(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss, const orc_step_in* in,
                 orc_step_out* out, int max_sqp_iter, double tol, int impl, int* sqp_iters, double* defect);
/* KKT certificate of an arbitrary candidate (X,U,dU[,lambda]) against the dense QP:
 * returns max(primal infeasibility, projected-gradient optimality gap) -- see .c */
double orc_check_candidate(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss,
                           const orc_step_in* in, const double* X, const double* U,
                           const double* dU, const double* lambda, double* cost_out,
                           double* prim_inf_out);

#ifdef __cplusplus
}
#endif
#endif
