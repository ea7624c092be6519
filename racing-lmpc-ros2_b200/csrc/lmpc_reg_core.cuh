// lmpc_reg_core.cuh -- error-dynamics regression over the stored laps, one warp per query item.
//
// Replaces SafeSetManager::query(const RegQuery&) -> RegResult and SSTrajectory::query(const RegQuery&)
// (reference src/vehicle_dynamics_models/racing_trajectory/src/safe_set.cpp:56-114,182-245; RegQuery / RegResult
// safe_set.hpp:61-88).  For every regression r (one output state o_r, input states S_r, input controls C_r):
//
//   points      all samples p of all stored laps that have a successor (safe_set.cpp:68-77)
//   distance    d_p = || (x_p[S_r], u_p[C_r]) - (x_q[S_r], u_q[C_r]) ||_2, kept when d_p < dist_max   (:81-88)
//   weights     K_p = 0.75 / h (1 - (d_p / h)^2)^2                                                    (:222-223)
//   regressors  m_p = (x_p[S_r], u_p[C_r], 1),  target y_p = x_{p+1}[o_r] - f_d(x_p, u_p, k_p, dt_p)[o_r]   (:225-227)
//   solve       (M' K M + ridge I) R = sign M' K y                                                   (:228-231)
//   update      A[o_r, S_r] += R[0:|S|],  B[o_r, C_r] += R[|S|:-1],  C[o_r] += R[-1]                   (:233-241)
//   no point within dist_max: that regression is skipped (:203-205)
//
// Stated deviations from the code as written (the function is never called in the reference and cannot run as written;
// SURVEY.md 8f #3): the model prediction uses the full stored state (the reference feeds f.map the *sliced* states,
// :217-220, a shape error); y is the scalar error of the output state (the reference slices reg_in_state_idxs, :226);
// dt_p = t_{p+1} - t_p > 0 (the reference stores t_p - t_{p+1}, :130-135); `sign` is a parameter: -1 reproduces
// b = -M'Ky as written (:229), +1 is the local-linear-regression correction of the LMPC paper (error = actual -
// predicted is *added* to the nominal model).  The per-lap sort by distance (:94-108) only permutes the sums.
//
// Both the test d_p < h and the weight are evaluated from d_p^2 (d_p^2 < h^2, K_p = 0.75/h (1 - d_p^2/h^2)^2): the same
// numbers to the last place or two, no square root in the scan.
//
// The model error y_p does not depend on the query: it is computed once per safe-set update by lmpc_reg_prepare
// (thread per point) and kept next to the points.  The scan is exact brute force over the device-resident slab
// (L2-resident: 14 doubles per point), lanes stride the points, partial normal equations live in registers
// (45 + 9 accumulators; 15 + 5 in the size class of at most four inputs), one butterfly all-reduce, then every lane runs
// the same Cholesky.  The CUDA kernel (lmpc_regress_tiled_kernel, lmpc_kernels.cuh) streams the points through
// shared-memory tiles, scans only the window of the sorted points a query can reach, instantiates the scan for the exact
// number of regressors and scans regressions with identical input lists as a pair.
#pragma once
#include "lmpc_warp.cuh"
#include "../../include/lmpc_b200.h"

#define LMPC_REG_D 9         // |S| + |C| + 1 <= 6 + 2 + 1

// One regression, ready for the device: sel[a] in 0..7 = component of (x, u); 8 = the constant 1; 9 = unused (0)
struct LmpcRegRow {
  int out;                 // output state
  int D;                   // |S| + |C| + 1
  int sel[LMPC_REG_D];
  int lead;                // the first regression with the same input lists (itself when there is none before it), and
  int follower;            // for a leader the next regression with its lists, or -1: the two share regressors and weights,
                           // hence M'KM -- the tiled kernel scans once for the pair and keeps a second M'Ky
};
struct LmpcRegPlan {
  int n_out;
  LmpcRegRow row[LMPC_REG_MAX_OUT];
  double h, ridge, sign;
};

// host: validate the caller's index lists and lay them out for the kernel (false = rejected)
static inline bool lmpc_make_reg_plan(const lmpc_reg_spec* sp, LmpcRegPlan* plan) {
  if (!sp || sp->n_out < 1 || sp->n_out > LMPC_REG_MAX_OUT || !(sp->dist_max > 0.0) || !(sp->ridge > 0.0) || !(sp->sign == 1.0 || sp->sign == -1.0)) return false;
  plan->n_out = sp->n_out; plan->h = sp->dist_max; plan->ridge = sp->ridge; plan->sign = sp->sign;
  for (int r = 0; r < sp->n_out; r++) {
    LmpcRegRow& row = plan->row[r];
    const int nx = sp->n_in_x[r], nu = sp->n_in_u[r];
    if (sp->out_idx[r] < 0 || sp->out_idx[r] >= 6 || nx < 0 || nx > 6 || nu < 0 || nu > 2) return false;
    row.out = sp->out_idx[r]; row.D = nx + nu + 1;
    int a = 0;
    for (int q = 0; q < nx; q++) { const int c = sp->in_x[r][q]; if (c < 0 || c >= 6) return false; row.sel[a++] = c; }
    for (int q = 0; q < nu; q++) { const int c = sp->in_u[r][q]; if (c < 0 || c >= 2) return false; row.sel[a++] = 6 + c; }
    row.sel[a++] = 8;
    while (a < LMPC_REG_D) row.sel[a++] = 9;
    row.lead = r; row.follower = -1;
    for (int e = 0; e < r && row.lead == r; e++) {   // pair it with an earlier unpaired leader of identical lists
      const LmpcRegRow& o = plan->row[e];
      bool same = o.lead == e && o.follower < 0 && o.D == row.D;
      for (int q = 0; same && q < LMPC_REG_D; q++) same = o.sel[q] == row.sel[q];
      if (same) { row.lead = e; plan->row[e].follower = r; }
    }
  }
  return true;
}

// Device view of the regression points, by column (a lane per sample reads consecutive doubles): Z [8][ld] = (x, u) of
// every sample with a successor, E [6][ld] its model error; M samples, ld >= M the column stride
struct LmpcRegView {
  const double* Z;
  const double* E;
  int M;
  int ld;
  int sort_dim;   // component of (x, u) the samples are sorted by (ascending), or -1: a query whose regressions all use that
                  // component only visits the samples with |z[sort_dim] - z_q[sort_dim]| < dist_max (binary search)
};

// One lane's share of the scan of `count` points (global memory or a shared-memory tile):
// points lane, lane + 32, ...  Accumulates the upper triangle of M'KM into Q, M'Ky into bv and the number of points
// within dist_max into cnt.  Visiting the points tile by tile (tiles a multiple of 32 long) keeps every lane's order.
// DD: compile-time size of the regression (5 covers the LMPC paper's choice of three states and one control; 9 everything).
// GATHERED = false: Z, E are the by-column arrays of the view (column sel[a] at Z + sel[a] ld).  GATHERED = true: Z is a tile that
// already holds the regression's own columns in order (column a at Z + a ld) and E its output column.
template <int DD, bool GATHERED>
LMPC_DEV void lmpc_reg_scan_lane(const LmpcRegRow& row, double h, double ih, double kc, const double* q, const double* Z,
                                 const double* E, int ld, int count, int lane, double* Q, double* bv, double& cnt) {
  constexpr int D = DD;
  // d < h and (d / h)^2 from the squared distance: no square root in the scan (the weight is the same polynomial in d^2)
  const double h2 = h * h, ih2 = ih * ih;
  for (int p = lane; p < count; p += 32) {
    double m[D], d2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; a++) {
      const int s = row.sel[a];
      m[a] = (s < 8) ? Z[(size_t)(GATHERED ? a : s) * ld + p] : (s == 8 ? 1.0 : 0.0);
      if (s < 8) { const double t = m[a] - q[a]; d2 += t * t; }
    }
    if (d2 < h2) {
      const double u1 = 1.0 - d2 * ih2;
      const double w = kc * u1 * u1;
      const double y = E[(size_t)(GATHERED ? 0 : row.out) * ld + p];
#pragma unroll
      for (int a = 0, k = 0; a < D; a++) {
        const double wa = w * m[a];
        bv[a] += wa * y;
#pragma unroll
        for (int b = a; b < D; b++, k++) Q[k] += wa * m[b];
      }
      cnt += 1.0;
    }
  }
}

// The same sums for a regression of exactly DE regressors (DE - 1 inputs and the constant) over a shared-memory tile that
// holds its inputs by column (column a at tZ + a * LD) and its output's error (tE): no index lists in the loop, the
// constant regressor folded (w * 1 = w).  Q, bv are laid out for the size class DD >= DE (the unused rows stay zero), the
// order of the additions per lane is that of lmpc_reg_scan_lane.
// TWO: a second output over the same regressors (tE2 -> bv2).
template <int DE, int DD, int LD, bool TWO>
LMPC_DEV void lmpc_reg_scan_tile(double h, double ih, double kc, const double* q, const double* tZ, const double* tE, const double* tE2,
                                 int count, int lane, double* Q, double* bv, double* bv2, double& cnt) {
  constexpr int NI = DE - 1;
  const double h2 = h * h, ih2 = ih * ih;
#pragma unroll 2
  for (int p = lane; p < count; p += 32) {
    double m[DE], d2 = 0.0;
#pragma unroll
    for (int a = 0; a < NI; a++) { m[a] = tZ[a * LD + p]; const double t = m[a] - q[a]; d2 += t * t; }
    m[NI] = 1.0;
    if (d2 < h2) {
      const double u1 = 1.0 - d2 * ih2;
      const double w = kc * u1 * u1;
      const double y = tE[p], y2 = TWO ? tE2[p] : 0.0;
#pragma unroll
      for (int a = 0; a < DE; a++) {
        const double wa = (a < NI) ? w * m[a] : w;
        bv[a] += wa * y;
        if (TWO) bv2[a] += wa * y2;
#pragma unroll
        for (int b = a; b < DE; b++) {
          const int k = a * DD - a * (a - 1) / 2 + (b - a);
          if (b < NI) Q[k] += wa * m[b]; else Q[k] += wa;
        }
      }
      cnt += 1.0;
    }
  }
}

// After the scan: all-reduce of the partial normal equations (acc: Q upper triangle row-major, then M'Ky, then the count),
// Cholesky of Q + ridge I (every lane the same), solve, update of A (6x6 column-major), B (6x2 column-major), C (6).
template <int DD>
LMPC_DEV void lmpc_reg_finish(const LmpcRegPlan& plan, const LmpcRegRow& row, LaneVar<double> (&acc)[DD * (DD + 1) / 2 + DD + 1],
                              double* A, double* B, double* C, int* npts_r) {
  constexpr int D = DD, NQ = D * (D + 1) / 2, NV = NQ + D + 1;
  {
    int ops[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) ops[k] = LMPC_RED_SUM;
    group_reduce<1, NV>(acc, ops, nullptr);
  }
  const int used = (int)(acc[NQ + D](0) + 0.5);
  LANES_BEGIN
    if (lane == 0 && npts_r) *npts_r = used;
  LANES_END
  if (used == 0) return;   // safe_set.cpp:203-205
  // ---- Cholesky of Q + ridge I (uniform: every lane holds the reduced sums), unused rows are ridge-only
  double Lc[NQ], R[D];       // lower factor stored in the same packed upper-triangle slots: L[b][a] at (a, b), a <= b
#pragma unroll
  for (int a = 0, k = 0; a < D; a++)
#pragma unroll
    for (int b = a; b < D; b++, k++) Lc[k] = acc[k](0) + (a == b ? plan.ridge : 0.0);
#define LMPC_RQ(a, b) Lc[(a) * D - (a) * ((a) - 1) / 2 + ((b) - (a))]
#pragma unroll
  for (int j = 0; j < D; j++) {
    double dg = LMPC_RQ(j, j);
#pragma unroll
    for (int k = 0; k < j; k++) dg -= LMPC_RQ(k, j) * LMPC_RQ(k, j);
    const double il = 1.0 / sqrt(dg);   // dg >= ridge > 0
    LMPC_RQ(j, j) = il;
#pragma unroll
    for (int i = j + 1; i < D; i++) {
      double a2 = LMPC_RQ(j, i);
#pragma unroll
      for (int k = 0; k < j; k++) a2 -= LMPC_RQ(k, i) * LMPC_RQ(k, j);
      LMPC_RQ(j, i) = a2 * il;
    }
  }
#pragma unroll
  for (int i = 0; i < D; i++) {
    double a2 = plan.sign * acc[NQ + i](0);
#pragma unroll
    for (int k = 0; k < i; k++) a2 -= LMPC_RQ(k, i) * R[k];
    R[i] = a2 * LMPC_RQ(i, i);
  }
#pragma unroll
  for (int i = D - 1; i >= 0; i--) {
    double a2 = R[i];
#pragma unroll
    for (int k = i + 1; k < D; k++) a2 -= LMPC_RQ(i, k) * R[k];
    R[i] = a2 * LMPC_RQ(i, i);
  }
#undef LMPC_RQ
  // ---- lane a adds R[a] where sel[a] points
  LANES_BEGIN
    if (lane < row.D) {
      double ra = R[0];
#pragma unroll
      for (int a = 1; a < D; a++) if (lane == a) ra = R[a];
      const int s = row.sel[lane];
      if (s < 6) A[row.out + 6 * s] += ra;
      else if (s < 8) B[row.out + 6 * (s - 6)] += ra;
      else if (s == 8) C[row.out] += ra;
    }
  LANES_END
}

// One query item over points in global memory (the emulator's and the reference form; the CUDA kernels stream the points
// the same way, lmpc_kernels.cuh).  zq [8] = (x_q, u_q).  A, B, C are updated in place.
// npts (optional): points used per regression [n_out].
template <int DD>
LMPC_DEV void lmpc_regress_warp_d(const LmpcRegPlan& plan, const LmpcRegView& v, const double* zq, double* A, double* B,
                                  double* C, int* npts) {
  constexpr int D = DD, NQ = D * (D + 1) / 2, NV = NQ + D + 1;
  const double h = plan.h, ih = 1.0 / h, kc = 0.75 / h;
  for (int r = 0; r < plan.n_out; r++) {
    const LmpcRegRow& row = plan.row[r];
    LaneVar<double> acc[NV];
    LANES_BEGIN
      double q[D], Q[NQ], bv[D], cnt = 0.0;
#pragma unroll
      for (int a = 0; a < D; a++) { q[a] = (row.sel[a] < 8) ? zq[row.sel[a]] : 0.0; bv[a] = 0.0; }
#pragma unroll
      for (int k = 0; k < NQ; k++) Q[k] = 0.0;
      lmpc_reg_scan_lane<DD, false>(row, h, ih, kc, q, v.Z, v.E, v.ld, v.M, lane, Q, bv, cnt);
#pragma unroll
      for (int k = 0; k < NQ; k++) acc[k](lane) = Q[k];
#pragma unroll
      for (int a = 0; a < D; a++) acc[NQ + a](lane) = bv[a];
      acc[NQ + D](lane) = cnt;
    LANES_END
    lmpc_reg_finish<DD>(plan, row, acc, A, B, C, npts ? npts + r : nullptr);
  }
}
// the size class of a plan: 5 when every regression has at most four inputs, else 9
LMPC_HD int lmpc_reg_size_class(const LmpcRegPlan& plan) {
  int dmax = 0;
  for (int r = 0; r < plan.n_out; r++) dmax = plan.row[r].D > dmax ? plan.row[r].D : dmax;
  return dmax <= 5 ? 5 : LMPC_REG_D;
}
LMPC_DEV void lmpc_regress_warp(const LmpcRegPlan& plan, const LmpcRegView& v, const double* zq, double* A, double* B,
                                double* C, int* npts) {
  if (lmpc_reg_size_class(plan) == 5) lmpc_regress_warp_d<5>(plan, v, zq, A, B, C, npts);
  else lmpc_regress_warp_d<LMPC_REG_D>(plan, v, zq, A, B, C, npts);
}
