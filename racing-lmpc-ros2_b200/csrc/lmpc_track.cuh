// lmpc_track.cuh -- the race track as the MPC sees it: cubic interpolants of the trajectory table over the
// abscissa, evaluated on the device.
//
// Replaces RacingTrajectory's CasADi `interpolant("bspline")` functions and what is built from them
// (reference src/vehicle_dynamics_models/racing_trajectory/src/racing_trajectory.cpp:25-186):
//   left / right boundary offset, centre-line x / y, speed profile        :62-94
//   yaw = atan2(y', x'), curvature (the reference's un-parenthesised formula)  :96-110
//   frenet_to_global                                                       :121-136
//   global_to_frenet (nearest waypoint seed :204-224, then minimise |r(s) - p|^2 over s :138-186)
//
// CasADi's 1-D "bspline" interpolant of degree 3 is THE C2 piecewise cubic through the data whose second and
// second-to-last grid points are not knots (BSplineInterpolant::not_a_knot).  That function is unique, so it is
// built here the classical way -- second derivatives from one tridiagonal solve with not-a-knot end conditions --
// and stored in piecewise-polynomial form: per interval j four coefficients of (s - s_j)^k.  The oracle builds
// the same function through a B-spline collocation system instead (oracle/oracle_track.py), tests also compare
// with scipy.interpolate.make_interp_spline.
#pragma once
#include <math.h>
#include <vector>
#include "lmpc_warp.cuh"
#include "lmpc_model.cuh"

enum { LMPC_TRK_LEFT = 0, LMPC_TRK_RIGHT = 1, LMPC_TRK_X = 2, LMPC_TRK_Y = 3, LMPC_TRK_VEL = 4, LMPC_TRK_NF = 5 };

struct LmpcTrack {
  int m;                 // break points (table rows + 7 wrapped rows, racing_trajectory.cpp:45-60)
  int n_way;             // way points (table rows) for the nearest-neighbour seed
  double L;              // total_length (DIST_TO_SF_FWD of row 0)
  const double* brk;     // [m]        abscissa of the padded table
  const double* coef;    // [NF][m-1][4]
  const double* way;     // [n_way][3] x, y, abscissa of the table rows
};

// what the MPC node reads at an abscissa (racing_mpc_node.cpp:261-265) plus the pose of the centre line
struct LmpcTrackPoint { double left, right, curvature, vel, x, y, yaw, dx, dy, d2x, d2y; };

// largest j in [0, m-2] with brk[j] <= s (clamped): binary search, ~8-11 steps
LMPC_HD int lmpc_track_interval(const LmpcTrack& T, double s) {
  int lo = 0, hi = T.m - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (T.brk[mid] <= s) lo = mid; else hi = mid;
  }
  return lo;
}

LMPC_HD double lmpc_cubic(const double* c, double t) { return ((c[3] * t + c[2]) * t + c[1]) * t + c[0]; }
LMPC_HD double lmpc_cubic_d1(const double* c, double t) { return (3.0 * c[3] * t + 2.0 * c[2]) * t + c[1]; }
LMPC_HD double lmpc_cubic_d2(const double* c, double t) { return 6.0 * c[3] * t + 2.0 * c[2]; }

// every interpolation function of the reference wraps its argument first: s_mod = align_abscissa(s, L/2, L)
// (racing_trajectory.cpp:96-97) -- note that s = 0 maps to L, not 0 (fmod(L, L) = 0)
LMPC_HD double lmpc_track_wrap(const LmpcTrack& T, double s) { return lmpc_align_abscissa(s, T.L / 2.0, T.L); }

LMPC_HD void lmpc_track_eval_mod(const LmpcTrack& T, double sm, LmpcTrackPoint* p) {
  const int j = lmpc_track_interval(T, sm);
  const double t = sm - T.brk[j];
  const size_t stride = 4 * (size_t)(T.m - 1);
  const double* c = T.coef + 4 * (size_t)j;
  p->left = lmpc_cubic(c + LMPC_TRK_LEFT * stride, t);
  p->right = lmpc_cubic(c + LMPC_TRK_RIGHT * stride, t);
  p->vel = lmpc_cubic(c + LMPC_TRK_VEL * stride, t);
  const double* cx = c + LMPC_TRK_X * stride; const double* cy = c + LMPC_TRK_Y * stride;
  p->x = lmpc_cubic(cx, t); p->y = lmpc_cubic(cy, t);
  const double dx = lmpc_cubic_d1(cx, t), dy = lmpc_cubic_d1(cy, t), d2x = lmpc_cubic_d2(cx, t), d2y = lmpc_cubic_d2(cy, t);
  p->dx = dx; p->dy = dy; p->d2x = d2x; p->d2y = d2y;
  p->yaw = atan2(dy, dx);
  // racing_trajectory.cpp:108-110 as written: dx d2y - dy d2x / sqrt((dx^2 + dy^2)^3)  (only the second term is divided)
  const double n2 = dx * dx + dy * dy;
  p->curvature = dx * d2y - dy * d2x / sqrt(n2 * n2 * n2);
}

LMPC_HD void lmpc_track_eval(const LmpcTrack& T, double s, LmpcTrackPoint* p) { lmpc_track_eval_mod(T, lmpc_track_wrap(T, s), p); }

// lmpc_utils align_yaw (utils.hpp:25-31)
LMPC_HD double lmpc_align_yaw(double yaw1, double yaw2) {
  const double d = yaw1 - yaw2;
  return atan2(sin(d), cos(d)) + yaw2;
}

// frenet_to_global (racing_trajectory.cpp:121-136): (s, t, xi) -> (x, y, phi)
LMPC_HD void lmpc_frenet_to_global(const LmpcTrack& T, const double* f, double* g) {
  LmpcTrackPoint p;
  // the reference wraps twice (s_mod, then again inside x_intp_): wrapping is idempotent except at s_mod = L -> L
  lmpc_track_eval(T, lmpc_track_wrap(T, f[0]), &p);
  g[0] = p.x - sin(p.yaw) * f[1];
  g[1] = p.y + cos(p.yaw) * f[1];
  g[2] = lmpc_align_yaw(p.yaw + f[2], 0.0);
}

// global_to_frenet (racing_trajectory.cpp:138-186,204-236): seed at the nearest way point, then minimise
// |r(s) - p|^2 over s (the reference: CasADi sqpmethod + qrqp on the same scalar problem), then
// t = |p - r(s)| * lateral_sign, xi = align_yaw(phi, yaw(s)) - yaw(s).
// Newton on g(s) = (r - p).r' with the exact second derivative; falls back to the Gauss-Newton curvature when the
// exact one is not positive; the step is limited to a quarter of the local way-point spacing times 8.
LMPC_HD void lmpc_global_to_frenet(const LmpcTrack& T, const double* g, double* f) {
  const double px = g[0], py = g[1];
  int best = 0; double bd = 1e300;
  for (int i = 0; i < T.n_way; i++) {
    const double ex = T.way[3 * i] - px, ey = T.way[3 * i + 1] - py, d = ex * ex + ey * ey;
    if (d < bd) { bd = d; best = i; }
  }
  double s = lmpc_track_wrap(T, T.way[3 * best + 2]);
  const double smax = 2.0 * T.L / (double)T.n_way;
  LmpcTrackPoint p;
  for (int it = 0; it < 50; it++) {
    lmpc_track_eval_mod(T, lmpc_track_wrap(T, s), &p);
    const double ex = p.x - px, ey = p.y - py;
    const double grad = ex * p.dx + ey * p.dy;
    const double gn = p.dx * p.dx + p.dy * p.dy;
    double hess = gn + ex * p.d2x + ey * p.d2y;
    if (!(hess > 1e-3 * gn)) hess = gn;
    double ds = -grad / hess;
    if (ds > smax) ds = smax; else if (ds < -smax) ds = -smax;
    s += ds;
    if (fabs(ds) < 1e-13 * fmax(1.0, fabs(s))) break;
  }
  s = lmpc_track_wrap(T, s);
  lmpc_track_eval_mod(T, lmpc_track_wrap(T, s), &p);
  const double cr = cos(p.yaw) * (py - p.y) - sin(p.yaw) * (px - p.x);   // lateral_sign (utils.hpp:72-80)
  const double sg = (double)((cr > 0.0) - (cr < 0.0));
  f[0] = s;
  f[1] = hypot(px - p.x, py - p.y) * sg;
  f[2] = lmpc_align_yaw(g[2], p.yaw) - p.yaw;
}

// ------------------------------------------------------------------------------------------------ host builder

struct LmpcTrackHost {
  int m = 0, n_way = 0;
  double L = 0.0;
  std::vector<double> brk, coef, way;
  LmpcTrack view() const { LmpcTrack T; T.m = m; T.n_way = n_way; T.L = L; T.brk = brk.data(); T.coef = coef.data(); T.way = way.data(); return T; }
};

// not-a-knot cubic interpolant of (x_j, y_j), j = 0..m-1 (m >= 4), as m-1 coefficient quadruples
static inline bool lmpc_notaknot_pp(const std::vector<double>& x, const std::vector<double>& y, double* coef) {
  const int m = (int)x.size();
  if (m < 4) return false;
  std::vector<double> h(m - 1), dlt(m - 1), M(m, 0.0);
  for (int j = 0; j < m - 1; j++) { h[j] = x[j + 1] - x[j]; if (!(h[j] > 0.0)) return false; dlt[j] = (y[j + 1] - y[j]) / h[j]; }
  // unknowns M_1..M_{m-2} after eliminating M_0 and M_{m-1} through the not-a-knot conditions
  //   h_1 M_0 - (h_0 + h_1) M_1 + h_0 M_2 = 0,   h_{m-2} M_{m-3} - (h_{m-3} + h_{m-2}) M_{m-2} + h_{m-3} M_{m-1} = 0
  const int nu = m - 2;
  std::vector<double> a(nu, 0.0), b(nu, 0.0), c(nu, 0.0), r(nu, 0.0);
  for (int k = 0; k < nu; k++) {
    const int j = k + 1;
    a[k] = h[j - 1]; b[k] = 2.0 * (h[j - 1] + h[j]); c[k] = h[j]; r[k] = 6.0 * (dlt[j] - dlt[j - 1]);
  }
  if (nu == 2) {
    // m = 4: a single cubic; both end conditions say the same thing -> M is linear in x (handled by the general code
    // below only for m >= 5), so fit the cubic through the four points directly (divided differences)
    const double d01 = dlt[0], d12 = dlt[1], d23 = dlt[2];
    const double d012 = (d12 - d01) / (x[2] - x[0]), d123 = (d23 - d12) / (x[3] - x[1]);
    const double d0123 = (d123 - d012) / (x[3] - x[0]);
    // Newton form at x0 -> power form at x_j
    for (int j = 0; j < 3; j++) {
      const double t0 = x[j] - x[0], t1 = x[j] - x[1], t2 = x[j] - x[2];
      coef[4 * j + 0] = y[j];
      coef[4 * j + 1] = d01 + d012 * (t0 + t1) + d0123 * (t0 * t1 + t0 * t2 + t1 * t2);
      coef[4 * j + 2] = d012 + d0123 * (t0 + t1 + t2);
      coef[4 * j + 3] = d0123;
    }
    return true;
  }
  // first row:  M_0 = ((h_0 + h_1) M_1 - h_0 M_2) / h_1
  b[0] = (h[0] + h[1]) * (h[0] + 2.0 * h[1]) / h[1]; c[0] = (h[1] * h[1] - h[0] * h[0]) / h[1]; a[0] = 0.0;
  // last row:   M_{m-1} = ((h_{m-3} + h_{m-2}) M_{m-2} - h_{m-2} M_{m-3}) / h_{m-3}
  { const double hp = h[m - 3], hq = h[m - 2];
    b[nu - 1] = (hp + hq) * (2.0 * hp + hq) / hp; a[nu - 1] = (hp * hp - hq * hq) / hp; c[nu - 1] = 0.0; }
  // Thomas
  for (int k = 1; k < nu; k++) { const double w = a[k] / b[k - 1]; b[k] -= w * c[k - 1]; r[k] -= w * r[k - 1]; }
  M[nu] = r[nu - 1] / b[nu - 1];
  for (int k = nu - 2; k >= 0; k--) M[k + 1] = (r[k] - c[k] * M[k + 2]) / b[k];
  M[0] = ((h[0] + h[1]) * M[1] - h[0] * M[2]) / h[1];
  M[m - 1] = ((h[m - 3] + h[m - 2]) * M[m - 2] - h[m - 2] * M[m - 3]) / h[m - 3];
  for (int j = 0; j < m - 1; j++) {
    coef[4 * j + 0] = y[j];
    coef[4 * j + 1] = dlt[j] - h[j] * (2.0 * M[j] + M[j + 1]) / 6.0;
    coef[4 * j + 2] = 0.5 * M[j];
    coef[4 * j + 3] = (M[j + 1] - M[j]) / (6.0 * h[j]);
  }
  return true;
}

// table: n rows of ncols (>= 13) columns in the reference's TrajectoryIndex order (racing_trajectory.hpp:37-56)
static inline bool lmpc_track_build(int n, int ncols, const double* table, LmpcTrackHost* out) {
  if (n < 8 || ncols < 13 || !table) return false;
  auto at = [&](int row, int col) { return table[(size_t)row * ncols + col]; };
  const double L = at(0, 7);   // total_length_ = traj_(DIST_TO_SF_FWD, 0)   racing_trajectory.cpp:29
  if (!(L > 0.0)) return false;
  // rows: last 3 (abscissa - L), all n, first 4 (abscissa + L)     racing_trajectory.cpp:45-60
  const int m = n + 7;
  std::vector<int> src(m); std::vector<double> off(m);
  for (int q = 0; q < 3; q++) { src[q] = n - 3 + q; off[q] = -L; }
  for (int q = 0; q < n; q++) { src[3 + q] = q; off[3 + q] = 0.0; }
  for (int q = 0; q < 4; q++) { src[3 + n + q] = q; off[3 + n + q] = L; }
  std::vector<double> s(m), f[LMPC_TRK_NF];
  for (auto& v : f) v.resize(m);
  for (int q = 0; q < m; q++) {
    const int r = src[q];
    s[q] = at(r, 6) + off[q];
    const double px = at(r, 0), py = at(r, 1);
    f[LMPC_TRK_LEFT][q] = hypot(px - at(r, 9), py - at(r, 10));      // :64-71
    f[LMPC_TRK_RIGHT][q] = -hypot(px - at(r, 11), py - at(r, 12));   // :72-79
    f[LMPC_TRK_X][q] = px; f[LMPC_TRK_Y][q] = py; f[LMPC_TRK_VEL][q] = at(r, 4);
  }
  out->m = m; out->n_way = n; out->L = L; out->brk = s;
  out->coef.assign((size_t)LMPC_TRK_NF * 4 * (m - 1), 0.0);
  for (int k = 0; k < LMPC_TRK_NF; k++)
    if (!lmpc_notaknot_pp(s, f[k], out->coef.data() + (size_t)k * 4 * (m - 1))) return false;
  out->way.resize(3 * (size_t)n);
  for (int r = 0; r < n; r++) { out->way[3 * r] = at(r, 0); out->way[3 * r + 1] = at(r, 1); out->way[3 * r + 2] = at(r, 6); }
  return true;
}
