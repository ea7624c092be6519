/*
 * oracle_model.c -- CPU ORACLE (test infrastructure only; see lmpc_oracle.h).
 *
 * Restates the single-track ("dynamic bicycle") model in the Frenet frame exactly as
 * SingleTrackPlanarModel::compile_dynamics writes it
 * (reference src/vehicle_dynamics_models/single_track_planar_model/src/single_track_planar_model.cpp:195-387)
 * and the RK4/Euler integrators of lmpc_utils (src/tools/lmpc_utils/src/utils.cpp:88-123).
 *
 * The reference obtains A = d x+/dx, B = d x+/du by CasADi SX algorithmic
 * differentiation of the integrator map.  The oracle does the same thing generically:
 * forward-mode dual numbers carrying 8 tangents (6 state + 2 control seeds) through the
 * identical expression graph.  (The CUDA product uses hand-derived analytic partials
 * instead, so the two are independent derivations; tests add sympy as a third.)
 */
#include <math.h>
#include <string.h>
#include "lmpc_oracle.h"

#define GRAVITY 9.8 /* single_track_planar_model.cpp:18 */
#define NT 8

typedef struct { double v; double d[NT]; } dual;

static dual dconst(double c) { dual r; r.v = c; memset(r.d, 0, sizeof r.d); return r; }
static dual dadd(dual a, dual b) { for (int i = 0; i < NT; i++) a.d[i] += b.d[i]; a.v += b.v; return a; }
static dual dsub(dual a, dual b) { for (int i = 0; i < NT; i++) a.d[i] -= b.d[i]; a.v -= b.v; return a; }
static dual dmul(dual a, dual b) {
  dual r; r.v = a.v * b.v;
  for (int i = 0; i < NT; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
static dual ddiv(dual a, dual b) {
  dual r; r.v = a.v / b.v;
  for (int i = 0; i < NT; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
static dual dscale(dual a, double c) { for (int i = 0; i < NT; i++) a.d[i] *= c; a.v *= c; return a; }
static dual daddc(dual a, double c) { a.v += c; return a; }
static dual dchain(dual a, double f, double fp) {
  dual r; r.v = f; for (int i = 0; i < NT; i++) r.d[i] = fp * a.d[i]; return r;
}
static dual dsin(dual a) { return dchain(a, sin(a.v), cos(a.v)); }
static dual dcos(dual a) { return dchain(a, cos(a.v), -sin(a.v)); }
static dual datan(dual a) { return dchain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
static dual dtanh(dual a) { double t = tanh(a.v); return dchain(a, t, 1.0 - t * t); }
static dual dneg(dual a) { return dscale(a, -1.0); }

/* x_dot = f(x,u,k): single_track_planar_model.cpp:199-332 with simplify_lon_control and
 * use_frenet both true (the only combination the launch files ship). */
static void f_dual(const orc_vehicle* p, const dual x[6], const dual u[2], double kappa, dual xd[6]) {
  const dual py = x[1], phi = x[2], vx = x[3], vy = x[4], omega = x[5];
  const dual v_sq = dmul(vx, vx);                                   /* :208 */
  const dual ulon = u[0], delta = u[1];
  /* :214-217 */
  const dual fd = dscale(dmul(ulon, daddc(dscale(dtanh(ulon), 0.5), 0.5)), 1000.0);
  const dual fb = dscale(dmul(ulon, daddc(dscale(dtanh(dneg(ulon)), 0.5), 0.5)), 1000.0);
  const double m = p->mass, Jzz = p->moi, l = p->wheel_base;
  const double lr = p->cg_ratio * l, lf = l - lr;                  /* :229-230 */
  const double fr = p->fr, hcog = p->cg_height;
  const double cl_f = p->cl_f, cl_r = p->cl_r, rho = p->air_density, A = p->frontal_area,
               cd = p->drag_coeff, mu = p->mu;
  const double kd = p->kd, kb = p->kb;
  /* :258-263 */
  const dual Fx_f = daddc(dadd(dscale(fd, 0.5 * kd), dscale(fb, 0.5 * kb)), -0.5 * fr * m * GRAVITY * lr / l);
  const dual Fx_r = daddc(dadd(dscale(fd, 0.5 * (1 - kd)), dscale(fb, 0.5 * (1.0 - kb))), -0.5 * fr * m * GRAVITY * lf / l);
  /* :267  (no rho here, as in the reference) */
  const dual ax = dscale(daddc(dsub(dadd(fd, fb), dscale(v_sq, 0.5 * cd * A)), -fr * m * GRAVITY), 1.0 / m);
  /* :270-276 */
  const dual Fz_f = dadd(daddc(dscale(ax, -0.5 * hcog / (lf + lr) * m), 0.5 * m * GRAVITY * lr / (lf + lr)),
                         dscale(v_sq, 0.25 * cl_f * rho * A));
  const dual Fz_r = dadd(daddc(dscale(ax, 0.5 * hcog / (lf + lr) * m), 0.5 * m * GRAVITY * lf / (lf + lr)),
                         dscale(v_sq, 0.25 * cl_r * rho * A));
  /* :280-283 */
  const dual vxe = daddc(vx, 1e-3);
  const dual a_f = dsub(delta, datan(ddiv(dadd(dscale(omega, lf), vy), vxe)));
  const dual a_r = datan(ddiv(dsub(dscale(omega, lr), vy), vxe));
  /* :299-300 */
  const dual Fy_f = dscale(dmul(Fz_f, dsin(dscale(datan(dscale(a_f, p->Bf)), p->Cf))), mu);
  const dual Fy_r = dscale(dmul(Fz_r, dsin(dscale(datan(dscale(a_r, p->Br)), p->Cr))), mu);
  const dual cd_ = dcos(delta), sd_ = dsin(delta);
  /* :309-310 */
  const dual omega_dot = dscale(
      dadd(dscale(Fy_r, -2.0 * lr), dscale(dadd(dmul(dscale(Fy_f, 2.0), cd_), dmul(dscale(Fx_f, 2.0), sd_)), lf)),
      1.0 / Jzz);
  /* :314-319 */
  const dual vx_dot = dadd(
      dscale(dsub(dsub(dadd(dscale(Fx_r, 2.0), dmul(dscale(Fx_f, 2.0), cd_)), dmul(dscale(Fy_f, 2.0), sd_)),
                  dscale(v_sq, 0.5 * cd * rho * A)),
             1.0 / m),
      dmul(omega, vy));
  const dual vy_dot = dsub(
      dscale(dadd(dadd(dscale(Fy_r, 2.0), dmul(dscale(Fy_f, 2.0), cd_)), dmul(dscale(Fx_f, 2.0), sd_)), 1.0 / m),
      dmul(omega, vx));
  /* :322-330 */
  dual px_dot = dsub(dmul(vx, dcos(phi)), dmul(vy, dsin(phi)));
  const dual py_dot = dadd(dmul(vx, dsin(phi)), dmul(vy, dcos(phi)));
  dual phi_dot = omega;
  px_dot = ddiv(px_dot, daddc(dscale(py, -kappa), 1.0));
  phi_dot = dsub(phi_dot, dscale(px_dot, kappa));
  xd[0] = px_dot; xd[1] = py_dot; xd[2] = phi_dot; xd[3] = vx_dot; xd[4] = vy_dot; xd[5] = omega_dot;
}

/* utils.cpp:88-123 : u and kappa held over the step */
static void step_dual(const orc_vehicle* p, const dual x[6], const dual u[2], double kappa, double dt, dual xn[6]) {
  dual k1[6], k2[6], k3[6], k4[6], xt[6];
  f_dual(p, x, u, kappa, k1);
  if (p->integrator == 1) { /* euler_function */
    for (int i = 0; i < 6; i++) xn[i] = dadd(x[i], dscale(k1[i], dt));
    return;
  }
  for (int i = 0; i < 6; i++) xt[i] = dadd(x[i], dscale(k1[i], dt / 2.0));
  f_dual(p, xt, u, kappa, k2);
  for (int i = 0; i < 6; i++) xt[i] = dadd(x[i], dscale(k2[i], dt / 2.0));
  f_dual(p, xt, u, kappa, k3);
  for (int i = 0; i < 6; i++) xt[i] = dadd(x[i], dscale(k3[i], dt));
  f_dual(p, xt, u, kappa, k4);
  for (int i = 0; i < 6; i++) {
    dual s = dadd(dadd(k1[i], dscale(k2[i], 2.0)), dadd(dscale(k3[i], 2.0), k4[i]));
    xn[i] = dadd(x[i], dscale(s, dt / 6.0));
  }
}

static void seed(const double x[6], const double u[2], dual xs[6], dual us[2], int with_tangents) {
  for (int i = 0; i < 6; i++) { xs[i] = dconst(x[i]); if (with_tangents) xs[i].d[i] = 1.0; }
  for (int i = 0; i < 2; i++) { us[i] = dconst(u[i]); if (with_tangents) us[i].d[6 + i] = 1.0; }
}

void orc_dynamics(const orc_vehicle* v, const double x[6], const double u[2], double kappa, double xdot[6]) {
  dual xs[6], us[2], xd[6];
  seed(x, u, xs, us, 0);
  f_dual(v, xs, us, kappa, xd);
  for (int i = 0; i < 6; i++) xdot[i] = xd[i].v;
}

void orc_discrete_dynamics(const orc_vehicle* v, const double x[6], const double u[2], double kappa, double dt,
                           double xnext[6]) {
  dual xs[6], us[2], xn[6];
  seed(x, u, xs, us, 0);
  step_dual(v, xs, us, kappa, dt, xn);
  for (int i = 0; i < 6; i++) xnext[i] = xn[i].v;
}

void orc_linearise(const orc_vehicle* v, const double x[6], const double u[2], double kappa, double dt,
                   double A[36], double B[12], double g[6], double xnext[6]) {
  dual xs[6], us[2], xn[6];
  seed(x, u, xs, us, 1);
  step_dual(v, xs, us, kappa, dt, xn);
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) A[r + 6 * c] = xn[r].d[c];
    for (int c = 0; c < 2; c++) B[r + 6 * c] = xn[r].d[6 + c];
  }
  /* gd = xip1 - (Ad x + Bd u)  (single_track_planar_model.cpp:379) */
  for (int r = 0; r < 6; r++) {
    double acc = 0.0;
    for (int c = 0; c < 6; c++) acc += A[r + 6 * c] * x[c];
    for (int c = 0; c < 2; c++) acc += B[r + 6 * c] * u[c];
    g[r] = xn[r].v - acc;
    if (xnext) xnext[r] = xn[r].v;
  }
}

/* utils.hpp:35-41 */
double orc_align_abscissa(double s1, double s2, double total) {
  const double k = fabs(s2 - s1) + total / 2.0;
  const double l = k - fmod(fabs(s2 - s1) + total / 2.0, total);
  const double d = s2 - s1;
  const double sg = (d > 0.0) - (d < 0.0);
  return s1 + l * sg;
}
