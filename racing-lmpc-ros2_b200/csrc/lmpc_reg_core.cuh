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
// The model error y_p does not depend on the query: it is computed once per safe-set update by lmpc_reg_prepare
// (thread per point) and kept next to the points.  The scan is exact brute force over the device-resident slab
// (L2-resident: 14 doubles per point), lanes stride the points, partial normal equations live in registers
// (45 + 9 accumulators), one butterfly all-reduce, then every lane runs the same 9x9 Cholesky.
#pragma once
#include "lmpc_warp.cuh"
#include "../../include/lmpc_b200.h"

#define LMPC_REG_D 9         // |S| + |C| + 1 <= 6 + 2 + 1

// One regression, ready for the device: sel[a] in 0..7 = component of (x, u); 8 = the constant 1; 9 = unused (0)
struct LmpcRegRow {
  int out;                 // output state
  int D;                   // |S| + |C| + 1
  int sel[LMPC_REG_D];
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
  }
  return true;
}

// Device view of the regression points: Z [M][8] = (x, u) of every sample with a successor, E [M][6] its model error
struct LmpcRegView {
  const double* Z;
  const double* E;
  int M;
};

// One query item.  zq [8] = (x_q, u_q).  A (6x6 column-major), B (6x2 column-major), C (6) are updated in place.
// npts (optional): points used per regression [n_out].
LMPC_DEV void lmpc_regress_warp(const LmpcRegPlan& plan, const LmpcRegView& v, const double* zq, double* A, double* B,
                                double* C, int* npts) {
  constexpr int D = LMPC_REG_D, NQ = D * (D + 1) / 2, NV = NQ + D + 1;
  const double h = plan.h, ih = 1.0 / h, kc = 0.75 / h;
  for (int r = 0; r < plan.n_out; r++) {
    const LmpcRegRow& row = plan.row[r];
    LaneVar<double> acc[NV];   // Q upper triangle (a <= b) row-major, then M'Ky, then the point count
    LANES_BEGIN
      double q[D], Q[NQ], bv[D], cnt = 0.0;
#pragma unroll
      for (int a = 0; a < D; a++) { q[a] = (row.sel[a] < 8) ? zq[row.sel[a]] : 0.0; bv[a] = 0.0; }
#pragma unroll
      for (int k = 0; k < NQ; k++) Q[k] = 0.0;
      for (int p = lane; p < v.M; p += 32) {
        const double* zp = v.Z + 8 * (size_t)p;
        double m[D], d2 = 0.0;
#pragma unroll
        for (int a = 0; a < D; a++) {
          const int s = row.sel[a];
          m[a] = (s < 8) ? zp[s] : (s == 8 ? 1.0 : 0.0);
          if (s < 8) { const double t = m[a] - q[a]; d2 += t * t; }
        }
        const double d = sqrt(d2);
        if (d < h) {
          const double t = d * ih, u1 = 1.0 - t * t;
          const double w = kc * u1 * u1;
          const double y = v.E[6 * (size_t)p + row.out];
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
#pragma unroll
      for (int k = 0; k < NQ; k++) acc[k](lane) = Q[k];
#pragma unroll
      for (int a = 0; a < D; a++) acc[NQ + a](lane) = bv[a];
      acc[NQ + D](lane) = cnt;
    LANES_END
    {
      int ops[NV];
#pragma unroll
      for (int k = 0; k < NV; k++) ops[k] = LMPC_RED_SUM;
      group_reduce<1, NV>(acc, ops, nullptr);
    }
    const int used = (int)(acc[NQ + D](0) + 0.5);
    LANES_BEGIN
      if (lane == 0 && npts) npts[r] = used;
    LANES_END
    if (used == 0) continue;   // safe_set.cpp:203-205
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
}
