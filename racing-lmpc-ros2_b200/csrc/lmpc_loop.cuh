// lmpc_loop.cuh -- the steps either side of the solve, on the device, so that a closed loop of many agents never
// leaves the GPU:
//
//   lmpc_prepare_kernel   what RacingMPCNode::on_step_timer does before mpc_->solve (reference
//                         src/mpc/racing_mpc/src/racing_mpc_node.cpp:236-292): x_ic prediction (continuous mode),
//                         shift-and-extend of the previous solution, track look-ups at its abscissae, velocity
//                         reference clipping -> every key of the solve's input dict
//   lmpc_plant_kernel     what happens after it: to_base_control and the actuation message (:386-401,
//                         single_track_planar_model.cpp:395-407), RacingSimulator::step (src/simulation/racing_simulator/
//                         src/racing_simulator.cpp:97-113: v_x floor, RK4 with the curvature at the car's abscissa,
//                         abscissa wrap) and the simulator node's lap counter (racing_simulator_node.cpp:283-286)
//
// One thread per (agent, horizon column) in prepare, one thread per agent in plant.
#pragma once
#include "lmpc_model.cuh"
#include "lmpc_track.cuh"
#include "../../include/lmpc_b200.h"

struct LmpcLoopParams {
  int step_mode, delay_step, plant_substeps;
  double dt, plant_dt, speed_limit, speed_scale, max_vel_ref_diff;
};

// discrete dynamics with the curvature taken from the track at the state's own abscissa: the node's and the
// simulator's `discrete_dynamics_` (racing_mpc_node.cpp:69-76, racing_simulator.cpp:46-57)
LMPC_HD void lmpc_step_on_track(const LmpcModel& M, const LmpcTrack& T, const double* x, const double* u, double dt, double* xn) {
  LmpcTrackPoint p;
  lmpc_track_eval(T, x[0], &p);
  lmpc_step(M, x, u, p.curvature, dt, xn);
}

// to_base_control followed by the actuation message's choice (racing_mpc_node.cpp:386-401): (u_lon, delta) ->
// (u_a, u_steer); the simulator turns that back into the derived control with from_base_control
// (racing_simulator_node.cpp:245-250, single_track_planar_model.cpp:401-407), which returns u_a itself.
// to_base_control (single_track_planar_model.cpp:390-400, simplify_lon_control): (u_lon, delta) -> (Fd, Fb, delta) with the
// logistic split Fd = u_lon / (1 + e^-u_lon), Fb = u_lon / (1 + e^u_lon) -- as written, WITHOUT the x1000 the dynamics use
LMPC_HD void lmpc_to_base_control(const double* u, double* ub) {
  ub[0] = u[0] * 1.0 / (1.0 + exp(-u[0]));
  ub[1] = u[0] * 1.0 / (1.0 + exp(u[0]));
  ub[2] = u[1];
}
// from_base_control (:401-407): the larger-magnitude one of (Fd, Fb), and the steering angle
LMPC_HD void lmpc_from_base_control(const double* ub, double* u) {
  u[0] = fabs(ub[0]) > fabs(ub[1]) ? ub[0] : ub[1];
  u[1] = ub[2];
}
LMPC_HD void lmpc_actuation(const double* u, double* ua) {
  double ub[3];
  lmpc_to_base_control(u, ub);
  lmpc_from_base_control(ub, ua);
}

// velocity reference of column i (racing_mpc_node.cpp:269-287)
LMPC_HD double lmpc_clip_vel_ref(const LmpcLoopParams& O, double track_vel, double current_speed) {
  const double lo = current_speed - O.max_vel_ref_diff, hi = current_speed + O.max_vel_ref_diff;
  const double ref_speed = track_vel * O.speed_scale;
  const double limit_clipped = fmin(fmax(O.speed_limit, lo), hi);
  if (ref_speed > 0.0) return fmin(fmin(fmax(ref_speed, lo), hi), limit_clipped);
  return limit_clipped;
}

#if !defined(LMPC_EMULATE)
struct LmpcLoopState {
  double* x;         // [B][6]  plant state (Frenet)
  double* u_prev;    // [B][2]  last published (u_a, u_steer) = the next tick's u_ic
  double* X_last;    // [B][N][6]   previous solution (last_x_)
  double* U_last;    // [B][N-1][2] previous solution (last_u_)
  int* lap_count;    // [B]
  int* fail_count;   // [B] ticks whose solve did not succeed
  int* tick;         // [1] device-side tick counter (indexes the log)
};

struct LmpcTickIn {   // the solve's input keys (device), filled by prepare
  double *x_ic, *u_ic, *X_ref, *U_ref, *T_ref, *bl, *br, *kap, *vref, *L;
};

__global__ void lmpc_prepare_kernel(LmpcModel M, LmpcTrack T, LmpcLoopParams O, int B, int N, LmpcLoopState S, LmpcTickIn I) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * N) return;
  const int b = t / N, i = t - b * N, NS = N - 1;
  const double* xl = S.X_last + (6 * (size_t)N) * b;
  const double* ul = S.U_last + (2 * (size_t)NS) * b;
  double xr[6];
  if (i < NS) {
    for (int c = 0; c < 6; c++) xr[c] = xl[6 * (i + 1) + c];                       // :245
    const int j = (i + 1 < NS) ? i + 1 : NS - 1;                                   // :246 (the last control is repeated)
    I.U_ref[(2 * (size_t)NS) * b + 2 * i] = ul[2 * j]; I.U_ref[(2 * (size_t)NS) * b + 2 * i + 1] = ul[2 * j + 1];
    I.T_ref[(size_t)NS * b + i] = O.dt;
  } else {
    lmpc_step_on_track(M, T, xl + 6 * NS, ul + 2 * (NS - 1), O.dt, xr);            // :248-249
  }
  for (int c = 0; c < 6; c++) I.X_ref[(6 * (size_t)N) * b + 6 * i + c] = xr[c];
  LmpcTrackPoint p;
  lmpc_track_eval(T, xr[0], &p);                                                   // :261-265
  I.bl[(size_t)N * b + i] = p.left; I.br[(size_t)N * b + i] = p.right; I.kap[(size_t)N * b + i] = p.curvature;
  I.vref[(size_t)N * b + i] = lmpc_clip_vel_ref(O, p.vel, xr[3]);
  if (i == 0) {
    const double* xm = S.x + 6 * (size_t)b;
    double xi[6];
    if (O.step_mode == 1) lmpc_step_on_track(M, T, xm, ul, O.dt, xi);              // :239 continuous: one step ahead with last_u[0]
    else for (int c = 0; c < 6; c++) xi[c] = xm[c];                                // :241
    for (int c = 0; c < 6; c++) I.x_ic[6 * (size_t)b + c] = xi[c];
    I.u_ic[2 * (size_t)b] = S.u_prev[2 * (size_t)b]; I.u_ic[2 * (size_t)b + 1] = S.u_prev[2 * (size_t)b + 1];
    I.L[b] = T.L;
  }
}

// after the solve: accept or keep the shifted reference (racing_mpc_node.cpp:322-331), publish, step the plant
__global__ void lmpc_plant_kernel(LmpcModel M, LmpcTrack T, LmpcLoopParams O, int B, int N, LmpcLoopState S, LmpcTickIn I,
                                  const double* __restrict__ X_optm, const double* __restrict__ U_optm,
                                  const int* __restrict__ status, double* __restrict__ log_x, double* __restrict__ log_u,
                                  int log_ticks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NS = N - 1;
  const bool ok = status[b] == LMPC_SOLVED || status[b] == LMPC_SOLVED_INACCURATE;
  double* xl = S.X_last + (6 * (size_t)N) * b; double* ul = S.U_last + (2 * (size_t)NS) * b;
  const double* xs = ok ? X_optm + (6 * (size_t)N) * b : I.X_ref + (6 * (size_t)N) * b;
  const double* us = ok ? U_optm + (2 * (size_t)NS) * b : I.U_ref + (2 * (size_t)NS) * b;
  for (int q = 0; q < 6 * N; q++) xl[q] = xs[q];
  for (int q = 0; q < 2 * NS; q++) ul[q] = us[q];
  if (!ok) S.fail_count[b] += 1;
  double ua[2];
  lmpc_actuation(ul + 2 * O.delay_step, ua);
  double x[6];
  for (int c = 0; c < 6; c++) x[c] = S.x[6 * (size_t)b + c];
  int laps = S.lap_count[b];
  for (int k = 0; k < O.plant_substeps; k++) {
    if (fabs(x[3]) < 1e-6) x[3] = copysign(1e-6, x[3]);                            // racing_simulator.cpp:99-103
    double xn[6];
    lmpc_step_on_track(M, T, x, ua, O.plant_dt, xn);
    xn[0] = lmpc_track_wrap(T, xn[0]);                                             // :58-62
    if (x[0] - xn[0] > 0.5 * T.L) laps++;                                          // racing_simulator_node.cpp:283-286
    for (int c = 0; c < 6; c++) x[c] = xn[c];
  }
  for (int c = 0; c < 6; c++) S.x[6 * (size_t)b + c] = x[c];
  S.u_prev[2 * (size_t)b] = ua[0]; S.u_prev[2 * (size_t)b + 1] = ua[1];
  S.lap_count[b] = laps;
  const int tk = *S.tick;
  if (log_x && tk < log_ticks) for (int c = 0; c < 6; c++) log_x[(6 * (size_t)B) * tk + 6 * (size_t)b + c] = x[c];
  if (log_u && tk < log_ticks) { log_u[(2 * (size_t)B) * tk + 2 * (size_t)b] = ua[0]; log_u[(2 * (size_t)B) * tk + 2 * (size_t)b + 1] = ua[1]; }
}

__global__ void lmpc_tick_advance_kernel(int* tick) { *tick += 1; }

// the model's control maps for n items: dir 0 = to_base_control ([n][2] -> [n][3]), 1 = from_base_control ([n][3] -> [n][2])
__global__ void lmpc_control_map_kernel(int n, int dir, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (dir == 0) { const double u[2] = {in[2 * (size_t)i], in[2 * (size_t)i + 1]}; double ub[3]; lmpc_to_base_control(u, ub); for (int c = 0; c < 3; c++) out[3 * (size_t)i + c] = ub[c]; }
  else { const double ub[3] = {in[3 * (size_t)i], in[3 * (size_t)i + 1], in[3 * (size_t)i + 2]}; double u[2]; lmpc_from_base_control(ub, u); out[2 * (size_t)i] = u[0]; out[2 * (size_t)i + 1] = u[1]; }
}

// ---- track interpolation functions for n abscissae / poses
__global__ void lmpc_track_eval_kernel(LmpcTrack T, int n, const double* __restrict__ s, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  LmpcTrackPoint p;
  lmpc_track_eval(T, s[t], &p);
  double* o = out + 7 * (size_t)t;
  o[0] = p.left; o[1] = p.right; o[2] = p.curvature; o[3] = p.vel; o[4] = p.x; o[5] = p.y; o[6] = p.yaw;
}
__global__ void lmpc_frenet_to_global_kernel(LmpcTrack T, int n, const double* __restrict__ f, double* __restrict__ g) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double fi[3] = {f[3 * (size_t)t], f[3 * (size_t)t + 1], f[3 * (size_t)t + 2]}, go[3];
  lmpc_frenet_to_global(T, fi, go);
  for (int c = 0; c < 3; c++) g[3 * (size_t)t + c] = go[c];
}
__global__ void lmpc_global_to_frenet_kernel(LmpcTrack T, int n, const double* __restrict__ g, double* __restrict__ f) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double gi[3] = {g[3 * (size_t)t], g[3 * (size_t)t + 1], g[3 * (size_t)t + 2]}, fo[3];
  lmpc_global_to_frenet(T, gi, fo);
  for (int c = 0; c < 3; c++) f[3 * (size_t)t + c] = fo[c];
}
#endif
