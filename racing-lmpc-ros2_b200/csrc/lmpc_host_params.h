// lmpc_host_params.h -- host-side conversion of the C-ABI PODs into the kernels' parameter blocks.
#pragma once
#include "../../include/lmpc_b200.h"
#include "lmpc_model.cuh"

static inline LmpcModel lmpc_make_model(const lmpc_vehicle_params& v) {
  LmpcModel P;
  P.m = v.mass; P.Jzz = v.moi; P.l = v.wheel_base;
  P.lr = v.cg_ratio * v.wheel_base;            // single_track_planar_model.cpp:229
  P.lf = v.wheel_base - P.lr;                  // :230
  P.fr = v.fr; P.hcog = v.cg_height; P.kd = v.kd; P.kb = v.kb;
  P.rho = v.air_density; P.Af = v.frontal_area; P.cd = v.drag_coeff; P.clf = v.cl_f; P.clr = v.cl_r;
  P.mu = v.mu; P.Bf = v.Bf; P.Cf = v.Cf; P.Br = v.Br; P.Cr = v.Cr;
  P.integrator = v.integrator;
  return P;
}

#include <math.h>
#include <string.h>
#include "lmpc_qp_core.cuh"

static inline bool lmpc_finite_bound(double b) { return isfinite(b) && fabs(b) < 1e19; }

// Validates the configuration and fills the kernel parameter block (row structure, merged boxes,
// shared-memory layout).  Returns LMPC_OK or LMPC_ERR_INVALID.
static inline int lmpc_make_qp_params(const lmpc_mpc_config& c, const lmpc_vehicle_params& v, LmpcQpParams* out, int NW = 1) {
  LmpcQpParams P;
  memset(&P, 0, sizeof P);
  if (c.N < 3 || c.N > LMPC_MAX_N) return LMPC_ERR_INVALID;
  P.N = c.N; P.NS = c.N - 1;
  P.learning = c.learning ? 1 : 0;
  P.K = P.learning ? c.num_ss_pts : 0;
  if (P.learning && (P.K < LMPC_MB || P.K > LMPC_MAX_SS_PTS || c.num_ss_pts_per_lap < 1)) return LMPC_ERR_INVALID;
  P.soft = c.q_boundary > 0.0 ? 1 : 0;                                  // racing_mpc.cpp:528
  double hs = 0.0;
  for (int k = 0; k < 6; k++) hs += c.convex_hull_slack[k] * c.convex_hull_slack[k];
  P.hull_slack = hs > 0.0 ? 1 : 0;                                      // racing_mpc.cpp:493
  P.nh = 0;
  if (P.learning)
    for (int k = 0; k < 6; k++) {
      if (P.hull_slack && c.convex_hull_slack[k] == 0.0) continue;      // free slack component: vacuous row
      P.hidx[P.nh] = k;
      P.Einv[P.nh] = P.hull_slack ? 1.0 / (2.0 * c.convex_hull_slack[k]) : 0.0;
      P.nh++;
    }
  for (int k = 0; k < 6; k++) P.chs[k] = P.hull_slack ? c.convex_hull_slack[k] : 0.0;
  P.nxb = 0;
  for (int k = 0; k < 6; k++) {
    P.xslot[k][0] = P.xslot[k][1] = -1;
    if (lmpc_finite_bound(c.x_max[k])) { P.xslot[k][0] = P.nxb; P.xb_c[P.nxb] = k; P.xb_sg[P.nxb] = 1.0; P.xb_h[P.nxb] = c.x_max[k]; P.nxb++; }
    if (lmpc_finite_bound(c.x_min[k])) { P.xslot[k][1] = P.nxb; P.xb_c[P.nxb] = k; P.xb_sg[P.nxb] = -1.0; P.xb_h[P.nxb] = -c.x_min[k]; P.nxb++; }
  }
  P.RS = P.nxb + 10;
  // merged u box: RacingMPC primal bounds (racing_mpc.cpp:148) and the model's actuator rows
  // (single_track_planar_model.cpp:114,120); rate box (:146-151)
  const double alo[2] = {v.Fb_max / 1000.0, -v.max_steer}, ahi[2] = {v.Fd_max / 1000.0, v.max_steer};
  for (int k = 0; k < 2; k++) { P.ulo[k] = fmax(c.u_min[k], alo[k]); P.uhi[k] = fmin(c.u_max[k], ahi[k]); }
  P.dlo[0] = v.Fb_max / 1000.0 / v.Tb; P.dhi[0] = v.Fd_max / 1000.0 / v.Td;
  P.dlo[1] = -v.max_steer_rate;        P.dhi[1] = v.max_steer_rate;
  for (int k = 0; k < 2; k++) {
    P.ub_act[2 * k] = lmpc_finite_bound(P.uhi[k]); P.ub_act[2 * k + 1] = lmpc_finite_bound(P.ulo[k]);
    P.db_act[2 * k] = lmpc_finite_bound(P.dhi[k]); P.db_act[2 * k + 1] = lmpc_finite_bound(P.dlo[k]);
    if (!(P.ulo[k] < P.uhi[k]) || !(P.dlo[k] < P.dhi[k])) return LMPC_ERR_INVALID;
  }
  P.margin = c.margin + v.chassis_b / 2.0;                              // racing_mpc.cpp:531
  P.qb = c.q_boundary;
  const double w[6] = {0.0, c.q_contour, c.q_heading, c.q_vel, c.q_vy, c.q_vyaw};   // racing_mpc.cpp:459-463
  for (int k = 0; k < 6; k++) { P.qx[k] = w[k]; P.qxN[k] = (k >= 1 && k <= 3) ? 10.0 * w[k] : 0.0; }   // :473-476
  P.Rm[0] = c.R[0]; P.Rm[1] = 0.5 * (c.R[1] + c.R[2]); P.Rm[2] = c.R[3];
  P.Rd[0] = c.R_d[0]; P.Rd[1] = 0.5 * (c.R_d[1] + c.R_d[2]); P.Rd[2] = c.R_d[3];
  P.max_iter = c.max_iter > 0 ? c.max_iter : 30;
  P.tol = c.tol > 0.0 ? c.tol : 1e-7;
  P.NSd = P.N | 1;
  if (NW != 1 && NW != 2 && NW != 4) return LMPC_ERR_INVALID;
  if (P.learning && (P.K + 32 * NW - 1) / (32 * NW) > LMPC_KPL_MAX) return LMPC_ERR_INVALID;
  P.NW = NW;
  // row table: state box rows (slot = their index among the finite bounds), then -- in slot order after them --
  // u box (c0 hi, c0 lo, c1 hi, c1 lo), rate box (same order), boundary (left = upper side of e_y, right = lower side)
  LmpcRowTab& T = P.rowtab;
  for (int k = 0; k < 6; k++)
    for (int r = 0; r < 2; r++) {
      T.xslot[2 * k + r] = P.xslot[k][r];
      T.xbnd[2 * k + r] = P.xslot[k][r] >= 0 ? P.xb_h[P.xslot[k][r]] : 0.0;
    }
  for (int k = 0; k < 2; k++)
    for (int r = 0; r < 2; r++) {
      T.uslot[2 * k + r] = P.ub_act[2 * k + r] ? P.nxb + 2 * k + r : -1;
      T.ubnd[2 * k + r] = r ? -P.ulo[k] : P.uhi[k];
      T.dslot[2 * k + r] = P.db_act[2 * k + r] ? P.nxb + 4 + 2 * k + r : -1;
      T.dbnd[2 * k + r] = r ? -P.dlo[k] : P.dhi[k];
    }
  T.bslot[0] = P.nxb + 8; T.bslot[1] = P.nxb + 9;
  T.ib0 = P.soft ? 0 : 1; T.pad_ = 0;
  // by slot: x rows exist on stages 1..N-2, boundary rows on ib0..N-1, control and rate rows on 0..N-2
  for (int q = 0; q < LMPC_MAX_ROWS; q++) { T.sbnd[q] = 0.0; T.sdesc[q] = 0; }
  auto desc = [](int c8, int lower, int isb, int rate, int i0, int last) { return c8 | (lower << 4) | (isb << 5) | (rate << 6) | (i0 << 8) | (last << 9) | (1 << 10); };   // bit 10: the slot holds a row
  for (int k = 0; k < 6; k++)
    for (int r = 0; r < 2; r++)
      if (T.xslot[2 * k + r] >= 0) { T.sdesc[T.xslot[2 * k + r]] = desc(k, r, 0, 0, 1, 0); T.sbnd[T.xslot[2 * k + r]] = T.xbnd[2 * k + r]; }
  for (int k = 0; k < 2; k++)
    for (int r = 0; r < 2; r++) {
      if (T.uslot[2 * k + r] >= 0) { T.sdesc[T.uslot[2 * k + r]] = desc(6 + k, r, 0, 0, 0, 0); T.sbnd[T.uslot[2 * k + r]] = T.ubnd[2 * k + r]; }
      if (T.dslot[2 * k + r] >= 0) { T.sdesc[T.dslot[2 * k + r]] = desc(6 + k, r, 0, 1, 0, 0); T.sbnd[T.dslot[2 * k + r]] = T.dbnd[2 * k + r]; }
    }
  for (int r = 0; r < 2; r++) T.sdesc[T.bslot[r]] = desc(1, r, 1, 0, T.ib0, 1);
  P.lay = lmpc_layout(P.N, P.RS, NW);
  *out = P;
  return LMPC_OK;
}
