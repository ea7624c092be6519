// lmpc_model.cuh -- single-track ("dynamic bicycle") model in the Frenet frame, its analytic
// partial derivatives, and the RK4/Euler step with its exact Jacobians A, B and affine term g.
//
// Replaces the CasADi SX graphs + algorithmic differentiation + JIT C code generation of
// SingleTrackPlanarModel::compile_dynamics (reference
// src/vehicle_dynamics_models/single_track_planar_model/src/single_track_planar_model.cpp:195-387)
// and lmpc_utils' rk4_function / euler_function (src/tools/lmpc_utils/src/utils.cpp:88-123).
// The derivatives here are hand-derived closed forms pushed through the integrator with the chain
// rule (dk_n = J_n (I + c_n dk_{n-1}) + G_n); the CPU oracle differentiates the same map with
// dual numbers instead, so the two are independent.
//
// State  x = (s, e_y, e_psi, v_x, v_y, omega)   base_vehicle_model.hpp:32-40
// Input  u = (u_lon, delta)                     single_track_planar_model.hpp:62-66 (simplify_lon_control)
#pragma once
#include "lmpc_warp.cuh"

#define LMPC_GRAVITY 9.8  // single_track_planar_model.cpp:18

struct LmpcModel {
  double m, Jzz, l, lr, lf, fr, hcog, kd, kb, rho, Af, cd, clf, clr, mu, Bf, Cf, Br, Cr;
  int integrator;  // 0 rk4, 1 euler
};

// f(x,u,kappa) and, when JAC, the non-zero partials.  Column order of the Jacobian rows:
// 0:e_y 1:e_psi 2:v_x 3:v_y 4:omega 5:u_lon 6:delta   (d/ds is identically zero)
// The transcendental part of f: independent chains (front tyre, rear tyre, controls, heading), kept apart from the algebra.
// (Measured in round 1: giving each chain of an item to its own thread -- four threads per item, two tangent columns each,
// bit-identical results -- made the linearisation kernel slower, 66-77 us against 51 us: the redundant algebra with its
// fifteen IEEE divisions and a second wave of blocks cost more than the shorter chains saved.  Round 2: half- and
// quarter-filled warps, to give every scheduler two or four chains, were slower too -- 75 and 97 us against 46 us: a wave of
// blocks lasts one chain whatever shares the scheduler, and the partial warps only add waves.  What shortens the chain is
// fewer instructions in it: the reciprocals below took it from 46 to 38 us.)
struct LmpcTrig { double th, sd, cdl, sph, cph, a_rf, a_rr, sqf, cqf, sqr, cqr; };
LMPC_HD void lmpc_trig_front(const LmpcModel& P, const double* x, const double* u, LmpcTrig& T) {
  const double ivxe = 1.0 / (x[3] + 1e-3);
  const double rf = (P.lf * x[5] + x[4]) * ivxe;
  T.a_rf = atan(rf);
  const double qf = P.Cf * atan(P.Bf * (u[1] - T.a_rf));
  sincos(qf, &T.sqf, &T.cqf);
}
LMPC_HD void lmpc_trig_rear(const LmpcModel& P, const double* x, LmpcTrig& T) {
  const double ivxe = 1.0 / (x[3] + 1e-3);
  const double rr = (P.lr * x[5] - x[4]) * ivxe;
  T.a_rr = atan(rr);
  const double qr = P.Cr * atan(P.Br * T.a_rr);
  sincos(qr, &T.sqr, &T.cqr);
}
LMPC_HD void lmpc_trig_controls(const double* u, LmpcTrig& T) { T.th = tanh(u[0]); sincos(u[1], &T.sd, &T.cdl); }
LMPC_HD void lmpc_trig_heading(const double* x, LmpcTrig& T) { sincos(x[2], &T.sph, &T.cph); }

template <bool JAC>
LMPC_HD void lmpc_f_algebra(const LmpcModel& P, const double* x, const double* u, double kappa, const LmpcTrig& T, double* xd,
                            double (*J)[7]) {
  const double ey = x[1], vx = x[3], vy = x[4], om = x[5];
  const double ul = u[0], de = u[1];
  const double m = P.m, l = P.l, lr = P.lr, lf = P.lf;
  // the divisions by the mass and the yaw inertia are multiplications by their reciprocals (24 IEEE divisions per evaluation
  // otherwise; a last-place difference from the formulas as written, far inside the parity tolerance)
  const double im = 1.0 / m, iJ = 1.0 / P.Jzz;
  // longitudinal command -> drive / brake force (single_track_planar_model.cpp:214-217)
  const double th = T.th;
  const double fd = ul * (0.5 * th + 0.5) * 1000.0;
  const double fb = ul * (-0.5 * th + 0.5) * 1000.0;  // tanh(-u) = -tanh(u)
  const double vsq = vx * vx;
  // :258-263
  const double Fxf = 0.5 * P.kd * fd + 0.5 * P.kb * fb - 0.5 * P.fr * m * LMPC_GRAVITY * lr / l;
  const double Fxr = 0.5 * (1.0 - P.kd) * fd + 0.5 * (1.0 - P.kb) * fb - 0.5 * P.fr * m * LMPC_GRAVITY * lf / l;
  // :267 (drag without rho here, as the reference writes it)
  const double ax = (fd + fb - 0.5 * P.cd * P.Af * vsq - P.fr * m * LMPC_GRAVITY) * im;
  // :270-276
  const double c4 = 0.5 * P.hcog / (lf + lr) * m;
  const double Fzf = 0.5 * m * LMPC_GRAVITY * lr / (lf + lr) - c4 * ax + 0.25 * P.clf * P.rho * P.Af * vsq;
  const double Fzr = 0.5 * m * LMPC_GRAVITY * lf / (lf + lr) + c4 * ax + 0.25 * P.clr * P.rho * P.Af * vsq;
  // :280-283
  const double ivxe = 1.0 / (vx + 1e-3);
  const double nf = lf * om + vy, nr = lr * om - vy;
  const double rf = nf * ivxe, rr = nr * ivxe;
  const double af = de - T.a_rf;
  const double ar = T.a_rr;
  // :299-300
  const double pf = P.Bf * af, pr = P.Br * ar;
  const double sqf = T.sqf, cqf = T.cqf, sqr = T.sqr, cqr = T.cqr, sd = T.sd, cdl = T.cdl, sph = T.sph, cph = T.cph;
  const double Fyf = P.mu * Fzf * sqf;
  const double Fyr = P.mu * Fzr * sqr;
  // :309-319
  const double latf = 2.0 * Fyf * cdl + 2.0 * Fxf * sd;  // front axle force along body y
  xd[5] = (-2.0 * Fyr * lr + latf * lf) * iJ;
  xd[3] = (2.0 * Fxr + 2.0 * Fxf * cdl - 2.0 * Fyf * sd - 0.5 * P.cd * P.rho * P.Af * vsq) * im + om * vy;
  xd[4] = (2.0 * Fyr + latf) * im - om * vx;
  // :322-330 (Frenet)
  const double iden = 1.0 / (1.0 - ey * kappa);
  const double num = vx * cph - vy * sph;
  const double sdot = num * iden;
  xd[0] = sdot;
  xd[1] = vx * sph + vy * cph;
  xd[2] = om - kappa * sdot;
  if (JAC) {
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 7; c++) J[r][c] = 0.0;
    // ---- kinematic rows
    J[0][0] = num * iden * iden * kappa;        // d sdot / d e_y
    J[0][1] = (-vx * sph - vy * cph) * iden;    // d / d e_psi
    J[0][2] = cph * iden;
    J[0][3] = -sph * iden;
    J[1][1] = vx * cph - vy * sph;
    J[1][2] = sph;
    J[1][3] = cph;
    for (int c = 0; c < 4; c++) J[2][c] = -kappa * J[0][c];
    J[2][4] = 1.0;
    // ---- force partials; index order of the small gradient arrays: [vx, vy, om, ul, de]
    const double sech2 = 1.0 - th * th;
    const double dfd = 1000.0 * ((0.5 * th + 0.5) + ul * 0.5 * sech2);
    const double dfb = 1000.0 * ((-0.5 * th + 0.5) - ul * 0.5 * sech2);
    const double dFxf_ul = 0.5 * P.kd * dfd + 0.5 * P.kb * dfb;
    const double dFxr_ul = 0.5 * (1.0 - P.kd) * dfd + 0.5 * (1.0 - P.kb) * dfb;
    const double dax_ul = (dfd + dfb) * im;
    const double dax_vx = -P.cd * P.Af * vx * im;
    const double dFzf_vx = -c4 * dax_vx + 0.5 * P.clf * P.rho * P.Af * vx;
    const double dFzr_vx = c4 * dax_vx + 0.5 * P.clr * P.rho * P.Af * vx;
    const double dFzf_ul = -c4 * dax_ul, dFzr_ul = c4 * dax_ul;
    const double wf = 1.0 / (1.0 + rf * rf), wr = 1.0 / (1.0 + rr * rr);
    // slip angles
    const double daf_vx = wf * nf * ivxe * ivxe, daf_vy = -wf * ivxe, daf_om = -wf * lf * ivxe;  // daf_de = 1
    const double dar_vx = -wr * nr * ivxe * ivxe, dar_vy = -wr * ivxe, dar_om = wr * lr * ivxe;
    // lateral tyre forces  Fy = mu Fz sin(C atan(B a))
    const double kf = P.mu * Fzf * cqf * P.Cf * P.Bf / (1.0 + pf * pf);
    const double kr = P.mu * Fzr * cqr * P.Cr * P.Br / (1.0 + pr * pr);
    const double dFyf[5] = {P.mu * dFzf_vx * sqf + kf * daf_vx, kf * daf_vy, kf * daf_om, P.mu * dFzf_ul * sqf, kf};
    const double dFyr[5] = {P.mu * dFzr_vx * sqr + kr * dar_vx, kr * dar_vy, kr * dar_om, P.mu * dFzr_ul * sqr, 0.0};
    const double dFxf[5] = {0.0, 0.0, 0.0, dFxf_ul, 0.0};
    const double dFxr[5] = {0.0, 0.0, 0.0, dFxr_ul, 0.0};
    for (int c = 0; c < 5; c++) {
      const double dlat = 2.0 * dFyf[c] * cdl + 2.0 * dFxf[c] * sd;
      J[5][2 + c] = (-2.0 * dFyr[c] * lr + dlat * lf) * iJ;
      J[3][2 + c] = (2.0 * dFxr[c] + 2.0 * dFxf[c] * cdl - 2.0 * dFyf[c] * sd) * im;
      J[4][2 + c] = (2.0 * dFyr[c] + dlat) * im;
    }
    // explicit delta dependence through cos/sin(delta)
    const double dlat_de = -2.0 * Fyf * sd + 2.0 * Fxf * cdl;
    J[5][6] += dlat_de * lf * iJ;
    J[3][6] += (-2.0 * Fxf * sd - 2.0 * Fyf * cdl) * im;
    J[4][6] += dlat_de * im;
    // drag and the omega*v coupling terms
    J[3][2] += -P.cd * P.rho * P.Af * vx * im;
    J[3][3] += om;
    J[3][4] += vy;
    J[4][2] += -om;
    J[4][4] += -vx;
  }
}

// Tu: lmpc_trig_controls(u) -- u is held over an integrator step, so its tanh / sincos are evaluated once per step
template <bool JAC>
LMPC_HD void lmpc_f_u(const LmpcModel& P, const double* x, const double* u, double kappa, const LmpcTrig& Tu, double* xd, double (*J)[7]) {
  LmpcTrig T = Tu;
  lmpc_trig_front(P, x, u, T);
  lmpc_trig_rear(P, x, T);
  lmpc_trig_heading(x, T);
  lmpc_f_algebra<JAC>(P, x, u, kappa, T, xd, J);
}
template <bool JAC>
LMPC_HD void lmpc_f(const LmpcModel& P, const double* x, const double* u, double kappa, double* xd, double (*J)[7]) {
  LmpcTrig T;
  lmpc_trig_controls(u, T);
  lmpc_f_u<JAC>(P, x, u, kappa, T, xd, J);
}

// one integrator step (u, kappa held over the step)
LMPC_HD void lmpc_step(const LmpcModel& P, const double* x, const double* u, double kappa, double dt, double* xn) {
  double k1[6], k2[6], k3[6], k4[6], xt[6];
  LmpcTrig Tu;
  lmpc_trig_controls(u, Tu);
  lmpc_f_u<false>(P, x, u, kappa, Tu, k1, nullptr);
  if (P.integrator == 1) {
    for (int i = 0; i < 6; i++) xn[i] = x[i] + dt * k1[i];
    return;
  }
  for (int i = 0; i < 6; i++) xt[i] = x[i] + dt / 2.0 * k1[i];
  lmpc_f_u<false>(P, xt, u, kappa, Tu, k2, nullptr);
  for (int i = 0; i < 6; i++) xt[i] = x[i] + dt / 2.0 * k2[i];
  lmpc_f_u<false>(P, xt, u, kappa, Tu, k3, nullptr);
  for (int i = 0; i < 6; i++) xt[i] = x[i] + dt * k3[i];
  lmpc_f_u<false>(P, xt, u, kappa, Tu, k4, nullptr);
  for (int i = 0; i < 6; i++) xn[i] = x[i] + dt / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
}

// S <- Jn * (seed + c * Dprev) + Gn  for the 8 tangent columns (6 state seeds + 2 control seeds).
// D matrices are 6 x 8 row-major; column 0 (d/ds) stays zero except for the identity seed.
LMPC_HD void lmpc_tangent_stage(const double (*J)[7], double c, const double* Dprev, double* Dout) {
  for (int col = 0; col < 8; col++) {
    double v[6];  // seed column + c * Dprev[:, col]  (state part only)
    for (int r = 0; r < 6; r++) v[r] = (Dprev ? c * Dprev[r * 8 + col] : 0.0) + ((col < 6 && r == col) ? 1.0 : 0.0);
    for (int r = 0; r < 6; r++) {
      double a = 0.0;
      for (int k = 1; k < 6; k++) a += J[r][k - 1] * v[k];     // state columns e_y..omega (s column is zero)
      if (col >= 6) a += J[r][5 + (col - 6)];                  // direct control partial
      Dout[r * 8 + col] = a;
    }
  }
}

// x+ and A (6x6 col-major), B (6x2 col-major), g = x+ - A x - B u
// (single_track_planar_model.cpp:377-379).
LMPC_HD void lmpc_linearise(const LmpcModel& P, const double* x, const double* u, double kappa, double dt,
                            double* A, double* B, double* g, double* xnext) {
  double J[6][7];
  double k[6], xt[6], acc[6], D[48], Dn[48], Dacc[48];
  LmpcTrig Tu;
  lmpc_trig_controls(u, Tu);
  if (P.integrator == 1) {
    lmpc_f_u<true>(P, x, u, kappa, Tu, k, J);
    lmpc_tangent_stage(J, 0.0, nullptr, D);
    for (int i = 0; i < 6; i++) acc[i] = x[i] + dt * k[i];
    for (int i = 0; i < 48; i++) Dacc[i] = dt * D[i];
  } else {
    // the four stages as ONE loop body (not four inlined copies: the kernel is a single pass over its code, and 97 KB of
    // straight-line instructions cost it a quarter of its time in instruction fetch).  Stage n evaluates at
    // x + c_n k_{n-1} and adds w_n k_n; c = (0, dt/2, dt/2, dt), w = (1, 2, 2, 1): the same products and sums as the
    // formulas written out (0 + 1 k, acc + 2 k are exact): bit-identical where products and sums are separate operations
    // (the emulator); on the GPU the compiler's choice of fused multiply-adds moves a last place here and there.
    for (int i = 0; i < 6; i++) { acc[i] = 0.0; k[i] = 0.0; }
    for (int i = 0; i < 48; i++) { Dacc[i] = 0.0; D[i] = 0.0; }
    LMPC_NOUNROLL
    for (int n = 0; n < 4; n++) {
      const double cn = (n == 0) ? 0.0 : (n == 3 ? dt : dt / 2.0), wn = (n == 0 || n == 3) ? 1.0 : 2.0;
      for (int i = 0; i < 6; i++) xt[i] = (n == 0) ? x[i] : x[i] + cn * k[i];
      lmpc_f_u<true>(P, xt, u, kappa, Tu, k, J);
      lmpc_tangent_stage(J, cn, D, Dn);
      for (int i = 0; i < 6; i++) acc[i] += wn * k[i];
      for (int i = 0; i < 48; i++) { Dacc[i] += wn * Dn[i]; D[i] = Dn[i]; }
    }
    const double w6 = dt / 6.0;
    for (int i = 0; i < 6; i++) acc[i] = x[i] + w6 * acc[i];
    for (int i = 0; i < 48; i++) Dacc[i] = w6 * Dacc[i];
  }
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) A[r + 6 * c] = Dacc[r * 8 + c] + (r == c ? 1.0 : 0.0);
    for (int c = 0; c < 2; c++) B[r + 6 * c] = Dacc[r * 8 + 6 + c];
  }
  for (int r = 0; r < 6; r++) {
    double a = 0.0;
    for (int c = 0; c < 6; c++) a += A[r + 6 * c] * x[c];
    for (int c = 0; c < 2; c++) a += B[r + 6 * c] * u[c];
    g[r] = acc[r] - a;
    if (xnext) xnext[r] = acc[r];
  }
}

// The same linearisation with the accumulated tangents kept OUTSIDE the registers: `o` is the item's 54-value result
// [A 36 | B 12 | g 6] with element k at o[k * S] (the linearisation kernel's shared-memory staging column).  The sum over
// the stages of w_n dk_n/d(x,u) is accumulated where it will be stored -- dk[r][col] belongs to element r + 6 col of
// [A | B] -- and the tangents of the previous stage are overwritten column by column (column col of the new stage depends
// on column col of the old one only), so one 6 x 7 array lives in registers instead of three 6 x 8 arrays: no local
// memory (a third of the kernel's stall samples were waiting for it).  Column 0 (d/ds) is identically zero and skipped.
// The arithmetic per element is that of lmpc_linearise (tests/test_emulator.py compares the two bit for bit).
#if defined(__CUDACC__) && !defined(LMPC_EMULATE)
#define LMPC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define LMPC_HD_NOINLINE static __attribute__((noinline))
#endif
// explicit Euler: one stage, through the plain form (its own function: its arrays stay out of the RK4 path's frame)
LMPC_HD_NOINLINE void lmpc_linearise_staged_euler(const LmpcModel& P, const double* x, const double* u, double kappa, double dt, double* o, int S) {
  double A[36], B[12], g[6];
  lmpc_linearise(P, x, u, kappa, dt, A, B, g, nullptr);
  for (int k = 0; k < 36; k++) o[k * S] = A[k];
  for (int k = 0; k < 12; k++) o[(36 + k) * S] = B[k];
  for (int k = 0; k < 6; k++) o[(48 + k) * S] = g[k];
}
LMPC_HD void lmpc_linearise_staged(const LmpcModel& P, const double* x, const double* u, double kappa, double dt, double* o_, int S) {
  if (P.integrator == 1) { lmpc_linearise_staged_euler(P, x, u, kappa, dt, o_, S); return; }
  volatile double* o = o_;   // every access is a load or a store: the sums must not be promoted back into registers
  double J[6][7], k[6], xt[6], acc[6], D[6][7];   // D[r][col - 1], col = 1..7
  LmpcTrig Tu;
  lmpc_trig_controls(u, Tu);
  for (int i = 0; i < 6; i++) { acc[i] = 0.0; k[i] = 0.0; }
  for (int r = 0; r < 6; r++) for (int c = 0; c < 7; c++) D[r][c] = 0.0;
  for (int e = 0; e < 48; e++) o[e * S] = 0.0;
  LMPC_NOUNROLL
  for (int n = 0; n < 4; n++) {
    const double cn = (n == 0) ? 0.0 : (n == 3 ? dt : dt / 2.0), wn = (n == 0 || n == 3) ? 1.0 : 2.0;
    for (int i = 0; i < 6; i++) xt[i] = (n == 0) ? x[i] : x[i] + cn * k[i];
    lmpc_f_u<true>(P, xt, u, kappa, Tu, k, J);
    for (int i = 0; i < 6; i++) acc[i] += wn * k[i];
    LMPC_UNROLL
    for (int col = 1; col < 8; col++) {
      double v[6];
      LMPC_UNROLL
      for (int r = 0; r < 6; r++) v[r] = cn * D[r][col - 1] + ((col < 6 && r == col) ? 1.0 : 0.0);
      LMPC_UNROLL
      for (int r = 0; r < 6; r++) {
        double a = 0.0;
        LMPC_UNROLL
        for (int q = 1; q < 6; q++) a += J[r][q - 1] * v[q];
        if (col >= 6) a += J[r][5 + (col - 6)];
        D[r][col - 1] = a;
        o[(r + 6 * col) * S] += wn * a;
      }
    }
  }
  const double w6 = dt / 6.0;
  for (int i = 0; i < 6; i++) acc[i] = x[i] + w6 * acc[i];
  for (int r = 0; r < 6; r++) {
    double a = 0.0;
    for (int c = 0; c < 8; c++) {
      const double t = w6 * o[(r + 6 * c) * S];
      const double e = (c < 6) ? t + (r == c ? 1.0 : 0.0) : t;
      o[(r + 6 * c) * S] = e;
      a += e * (c < 6 ? x[c] : u[c - 6]);
    }
    o[(48 + r) * S] = acc[r] - a;
  }
}

// lmpc_utils align_abscissa (src/tools/lmpc_utils/include/lmpc_utils/utils.hpp:35-41)
LMPC_HD double lmpc_align_abscissa(double s1, double s2, double total) {
  const double k = fabs(s2 - s1) + total / 2.0;
  const double lq = k - fmod(fabs(s2 - s1) + total / 2.0, total);
  const double d = s2 - s1;
  const double sg = (double)((d > 0.0) - (d < 0.0));
  return s1 + lq * sg;
}
