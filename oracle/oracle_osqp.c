/* oracle_osqp.c -- CPU ORACLE (test infrastructure only; see lmpc_oracle.h): a restatement of what the REFERENCE's own
 * solver stack does with the tick's QP, to measure how far the reference's output sits from the exact optimum.
 *
 * The reference poses the QP in SCALED decision variables X_ = x / scale_x, U_ = u / scale_u, dU_ = du / scale_u with
 * scale_x = (2000, 10, 0.1, 80, 2, 2), scale_u = (10, 0.3) (racing_mpc.cpp:36-37,127-129,141,193,200,533) through
 * casadi::Opti("conic") and solves it with OSQP, options {polish: true} and OSQP's defaults otherwise
 * (racing_mpc.cpp:86-103).  Neither CasADi nor OSQP is part of /root/reference (CasADi "main", OSQP as
 * bundled); this file restates OSQP's PUBLISHED algorithm (Stellato et al., "OSQP: an operator splitting solver for
 * quadratic programs", Math. Prog. Comp. 2020; defaults of the 0.6 series):
 *   - rows as Opti states them: two-sided rows for the boxes (state, input, actuator, rate: racing_mpc.cpp:146-148,
 *     single_track_planar_model.cpp:114,120,146-151), one-sided rows where a bound depends on a variable (the soft track
 *     boundary, :534-537), equalities l = u; CasADi's OSQP plugin prepends the identity rows of the (here infinite)
 *     variable bounds -- `with_var_rows` (default 1) reproduces that
 *   - Ruiz equilibration of [P A'; A 0], 10 passes, scalings limited to [1e-4, 1e4], then the cost scaling c
 *   - ADMM with rho = 0.1 (x 1e3 on equality rows, 1e-6 on rows without finite bound), sigma = 1e-6, alpha = 1.6
 *   - termination on the UNSCALED residuals, eps_abs = eps_rel = 1e-3, checked every 25 iterations, max_iter 4000
 *     (eps_in / max_iter_in > 0 override them: the test that the restated problem IS the QP runs the ADMM to 1e-9)
 *   - adaptive rho: OSQP derives the update interval from wall-clock times (not reproducible); restated with a fixed
 *     interval `rho_interval` (default 100 = OSQP's ADAPTIVE_RHO_FIXED), update when the estimate moves by more than 5x
 *   - polish: active rows z_i - l_i < -y_i / u_i - z_i < y_i, KKT with delta = 1e-6 regularisation, 3 refinement steps,
 *     accepted only when it improves the residuals (else the ADMM point is what the caller gets)
 * "Parity unpinned" applies here as everywhere: this is the published algorithm, not the reference's binary. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

#define OSQP_INF 1e30
typedef struct { int n, m; double *P, *q, *A, *l, *u; int nvar_rows; } oqp;   /* dense row-major */

static int lu_factor_(double* M, int n, int* piv) {
  for (int k = 0; k < n; k++) {
    int pk = k; double mx = fabs(M[(size_t)k * n + k]);
    for (int i = k + 1; i < n; i++) { const double a = fabs(M[(size_t)i * n + k]); if (a > mx) { mx = a; pk = i; } }
    piv[k] = pk;
    if (mx == 0.0) return -1;
    if (pk != k) for (int j = 0; j < n; j++) { const double t = M[(size_t)k * n + j]; M[(size_t)k * n + j] = M[(size_t)pk * n + j]; M[(size_t)pk * n + j] = t; }
    const double inv = 1.0 / M[(size_t)k * n + k];
    for (int i = k + 1; i < n; i++) {
      double* ri = M + (size_t)i * n; const double* rk = M + (size_t)k * n;
      const double f = ri[k] * inv;
      if (f == 0.0) continue;
      ri[k] = f;
      for (int j = k + 1; j < n; j++) ri[j] -= f * rk[j];
    }
  }
  return 0;
}
static void lu_solve_(const double* M, int n, const int* piv, double* b) {
  for (int k = 0; k < n; k++) if (piv[k] != k) { const double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
  for (int i = 0; i < n; i++) { double s = b[i]; const double* ri = M + (size_t)i * n; for (int j = 0; j < i; j++) s -= ri[j] * b[j]; b[i] = s; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; const double* ri = M + (size_t)i * n; for (int j = i + 1; j < n; j++) s -= ri[j] * b[j]; b[i] = s / ri[i]; }
}
static int finite_b(double b) { return isfinite(b) && fabs(b) < 1e19; }

/* ---- the QP in the reference's scaled variables, rows in Opti's form.  Variable order: X_ (6N), U_ (2(N-1)), dU_ (2(N-1)),
 *      sigma_b, lambda (K), sigma_h (6). */
static void oqp_build(const orc_vehicle* vp, const orc_config* c, const orc_prob* p, int with_var_rows, oqp* Q, double* scale) {
  const int N = p->N, K = p->K, NS = N - 1;
  const double sx[6] = {2000.0, 10.0, 0.1, 80.0, 2.0, 2.0}, su[2] = {10.0, 0.3};   /* racing_mpc.cpp:36-37 */
  int n = 6 * N + 4 * NS;
  const int oU = 6 * N, oD = 6 * N + 2 * NS;
  int oSb = -1, oL = -1, oSh = -1;
  if (p->soft_boundary) { oSb = n; n += 1; }
  if (p->learning) { oL = n; n += K; if (p->hull_slack) { oSh = n; n += 6; } }
  for (int i = 0; i < n; i++) scale[i] = 1.0;
  for (int i = 0; i < N; i++) for (int k = 0; k < 6; k++) scale[6 * i + k] = sx[k];
  for (int i = 0; i < NS; i++) for (int k = 0; k < 2; k++) { scale[oU + 2 * i + k] = su[k]; scale[oD + 2 * i + k] = su[k]; }
  const int nvr = with_var_rows ? n : 0;
  const int mcap = nvr + NS * (2 + 2 + 6 + 2 + 6 + 2) + 6 + 2 * N + 1 + K + 1 + 6;
  Q->n = n; Q->nvar_rows = nvr;
  Q->P = (double*)calloc((size_t)n * n, sizeof(double)); Q->q = (double*)calloc((size_t)n, sizeof(double));
  Q->A = (double*)calloc((size_t)mcap * n, sizeof(double)); Q->l = (double*)calloc((size_t)mcap, sizeof(double)); Q->u = (double*)calloc((size_t)mcap, sizeof(double));
#define PP(i, j) Q->P[(size_t)(i) * n + (j)]
#define AA(r, j) Q->A[(size_t)(r) * n + (j)]
  /* cost in physical variables v = S v_: 1/2 v'Hv + g'v  ->  P = S H S, q = S g */
  if (p->soft_boundary) PP(oSb, oSb) += 2.0 * c->q_boundary;
  for (int i = 0; i < NS; i++)
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 2; b++) {
        PP(oU + 2 * i + a, oU + 2 * i + b) += (c->R[2 * a + b] + c->R[2 * b + a]) * su[a] * su[b];
        PP(oD + 2 * i + a, oD + 2 * i + b) += (c->R_d[2 * a + b] + c->R_d[2 * b + a]) * su[a] * su[b];
      }
  if (p->learning) {
    if (p->hull_slack) for (int k = 0; k < 6; k++) PP(oSh + k, oSh + k) += 2.0 * c->convex_hull_slack[k];
    for (int k = 0; k < K; k++) Q->q[oL + k] += p->ssc[k];
  } else {
    const double w[6] = {0.0, c->q_contour, c->q_heading, c->q_vel, c->q_vy, c->q_vyaw};
    for (int i = 0; i < N; i++) {
      const double sc = (i == N - 1) ? 10.0 : 1.0;
      for (int k = 1; k < 6; k++) { if (i == N - 1 && k >= 4) continue; PP(6 * i + k, 6 * i + k) += 2.0 * sc * w[k] * sx[k] * sx[k]; }
      Q->q[6 * i + 3] += -2.0 * sc * c->q_vel * p->vref[i] * sx[3];
    }
  }
  int r = 0;
  for (int i = 0; i < nvr; i++) { AA(r, i) = 1.0; Q->l[r] = -OSQP_INF; Q->u[r] = OSQP_INF; r++; }   /* CasADi's plugin: lbx <= x <= ubx as rows */
  const double m = p->margin;
  for (int i = 0; i < N; i++) {   /* track boundary (racing_mpc.cpp:529-541) */
    if (p->soft_boundary) {
      AA(r, 6 * i + 1) = sx[1]; AA(r, oSb) = -1.0; Q->l[r] = -OSQP_INF; Q->u[r] = p->bl[i] - m; r++;
      AA(r, 6 * i + 1) = sx[1]; AA(r, oSb) = 1.0; Q->l[r] = p->br[i] + m; Q->u[r] = OSQP_INF; r++;
    } else { AA(r, 6 * i + 1) = sx[1]; Q->l[r] = p->br[i] + m; Q->u[r] = p->bl[i] - m; r++; }
  }
  if (p->soft_boundary) { AA(r, oSb) = 1.0; Q->l[r] = 0.0; Q->u[r] = OSQP_INF; r++; }
  if (p->learning) {   /* :490-503 */
    for (int k = 0; k < K; k++) { AA(r, oL + k) = 1.0; Q->l[r] = 0.0; Q->u[r] = OSQP_INF; r++; }
    for (int k = 0; k < K; k++) AA(r, oL + k) = 1.0;
    Q->l[r] = Q->u[r] = 1.0; r++;
    for (int k = 0; k < 6; k++) {
      AA(r, 6 * (N - 1) + k) = sx[k];
      for (int j = 0; j < K; j++) AA(r, oL + j) = -p->ssx[6 * j + k];
      if (p->hull_slack) AA(r, oSh + k) = -1.0;
      Q->l[r] = Q->u[r] = 0.0; r++;
    }
  }
  const double alo[2] = {vp->Fb_max / 1000.0, -vp->max_steer}, ahi[2] = {vp->Fd_max / 1000.0, vp->max_steer};
  for (int i = 0; i < NS; i++) {
    for (int k = 0; k < 2; k++) { AA(r, oU + 2 * i + k) = su[k]; Q->l[r] = alo[k]; Q->u[r] = ahi[k]; r++; }             /* model rows on u */
    for (int k = 0; k < 2; k++) { AA(r, oD + 2 * i + k) = su[k]; Q->l[r] = p->dlo[k]; Q->u[r] = p->dhi[k]; r++; }       /* model rows on du */
    for (int k = 0; k < 6; k++) {                                                                                     /* :147 */
      AA(r, 6 * i + k) = sx[k];
      Q->l[r] = finite_b(c->x_min[k]) ? c->x_min[k] : -OSQP_INF; Q->u[r] = finite_b(c->x_max[k]) ? c->x_max[k] : OSQP_INF; r++;
    }
    for (int k = 0; k < 2; k++) {                                                                                     /* :148 */
      AA(r, oU + 2 * i + k) = su[k];
      Q->l[r] = finite_b(c->u_min[k]) ? c->u_min[k] : -OSQP_INF; Q->u[r] = finite_b(c->u_max[k]) ? c->u_max[k] : OSQP_INF; r++;
    }
    for (int k = 0; k < 6; k++) {                                                                                     /* :182 */
      AA(r, 6 * (i + 1) + k) = sx[k];
      for (int j = 0; j < 6; j++) AA(r, 6 * i + j) -= p->A[36 * i + k + 6 * j] * sx[j];
      for (int j = 0; j < 2; j++) AA(r, oU + 2 * i + j) -= p->B[12 * i + k + 6 * j] * su[j];
      Q->l[r] = Q->u[r] = p->g[6 * i + k]; r++;
    }
    for (int k = 0; k < 2; k++) {                                                                                     /* :195  uim1 + dui ti == ui */
      AA(r, oU + 2 * i + k) = -su[k]; AA(r, oD + 2 * i + k) = su[k] * p->T[i];
      if (i == 0) { Q->l[r] = Q->u[r] = -p->u_ic[k]; } else { AA(r, oU + 2 * (i - 1) + k) = su[k]; Q->l[r] = Q->u[r] = 0.0; }
      r++;
    }
  }
  for (int k = 0; k < 6; k++) { AA(r, k) = sx[k]; Q->l[r] = Q->u[r] = p->x_ic[k]; r++; }                               /* :199-201 */
  Q->m = r;
#undef PP
#undef AA
}
static void oqp_free(oqp* Q) { free(Q->P); free(Q->q); free(Q->A); free(Q->l); free(Q->u); }

static double vinf(const double* v, int n) { double m = 0.0; for (int i = 0; i < n; i++) if (fabs(v[i]) > m) m = fabs(v[i]); return m; }

/* out: X, U, dU (physical units), lambda; info[8] = {iterations, status (0 solved, 1 max_iter), polish accepted, ADMM
 * primal residual, ADMM dual residual, polished primal residual, polished dual residual, rho at the end} */
int orc_step_osqp(const orc_vehicle* vp, const orc_config* c, const orc_safe_set* ss, const orc_step_in* in, orc_step_out* out,
                  int with_var_rows, int rho_interval, int do_polish, int warm, double eps_in, int max_iter_in, double* info) {
  orc_prob* p = (orc_prob*)malloc(sizeof *p);
  int st = orc_build_prob(vp, c, ss, in, p);
  if (st != ORC_OK) { free(p); out->status = st; return st; }
  const int N = p->N, K = p->K, NS = N - 1;
  double* scl = (double*)malloc(sizeof(double) * (size_t)(10 * N + K + 16));
  oqp Q; oqp_build(vp, c, p, with_var_rows, &Q, scl);
  const int n = Q.n, m = Q.m, nk = n + m;
  /* ---- Ruiz equilibration (10 passes) + cost scaling */
  double* D = (double*)malloc(sizeof(double) * (size_t)n); double* E = (double*)malloc(sizeof(double) * (size_t)m);
  for (int i = 0; i < n; i++) D[i] = 1.0;
  for (int i = 0; i < m; i++) E[i] = 1.0;
  double cs = 1.0;
  double* dt = (double*)malloc(sizeof(double) * (size_t)n); double* et = (double*)malloc(sizeof(double) * (size_t)m);
  for (int pass = 0; pass < 10; pass++) {
    for (int j = 0; j < n; j++) {   /* column norms of [P; A] */
      double v = 0.0;
      for (int i = 0; i < n; i++) v = fmax(v, fabs(Q.P[(size_t)i * n + j]));
      for (int i = 0; i < m; i++) v = fmax(v, fabs(Q.A[(size_t)i * n + j]));
      v = v < 1e-4 ? 1.0 : (v > 1e4 ? 1e4 : v);
      dt[j] = 1.0 / sqrt(v);
    }
    for (int i = 0; i < m; i++) {   /* row norms of A */
      double v = 0.0;
      for (int j = 0; j < n; j++) v = fmax(v, fabs(Q.A[(size_t)i * n + j]));
      v = v < 1e-4 ? 1.0 : (v > 1e4 ? 1e4 : v);
      et[i] = 1.0 / sqrt(v);
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Q.P[(size_t)i * n + j] *= dt[i] * dt[j];
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) Q.A[(size_t)i * n + j] *= et[i] * dt[j];
    for (int j = 0; j < n; j++) { Q.q[j] *= dt[j]; D[j] *= dt[j]; }
    for (int i = 0; i < m; i++) E[i] *= et[i];
    double pn = 0.0;   /* cost scaling: 1 / max(mean column norm of P, |q|_inf) */
    for (int j = 0; j < n; j++) { double v = 0.0; for (int i = 0; i < n; i++) v = fmax(v, fabs(Q.P[(size_t)i * n + j])); pn += v; }
    pn /= n;
    double cq = fmax(pn, vinf(Q.q, n));
    cq = cq < 1e-4 ? 1.0 : (cq > 1e4 ? 1e4 : cq);
    const double ct = 1.0 / cq;
    for (size_t i = 0; i < (size_t)n * n; i++) Q.P[i] *= ct;
    for (int j = 0; j < n; j++) Q.q[j] *= ct;
    cs *= ct;
  }
  for (int i = 0; i < m; i++) { if (Q.l[i] > -OSQP_INF * 1e-4) Q.l[i] *= E[i]; else Q.l[i] = -OSQP_INF; if (Q.u[i] < OSQP_INF * 1e-4) Q.u[i] *= E[i]; else Q.u[i] = OSQP_INF; }
  /* ---- ADMM */
  const double sigma = 1e-6, alpha = 1.6, eps = eps_in > 0.0 ? eps_in : 1e-3;   /* OSQP default eps_abs = eps_rel = 1e-3 */
  const int max_it = max_iter_in > 0 ? max_iter_in : 4000;
  double rho = 0.1;
  double* rv = (double*)malloc(sizeof(double) * (size_t)m);
  double* KK = (double*)malloc(sizeof(double) * (size_t)nk * nk); int* piv = (int*)malloc(sizeof(int) * (size_t)nk);
  double* x = (double*)calloc((size_t)n, sizeof(double)); double* z = (double*)calloc((size_t)m, sizeof(double)); double* y = (double*)calloc((size_t)m, sizeof(double));
  double* rhs = (double*)malloc(sizeof(double) * (size_t)nk); double* xp = (double*)malloc(sizeof(double) * (size_t)n); double* zp = (double*)malloc(sizeof(double) * (size_t)m);
  double* Ax = (double*)malloc(sizeof(double) * (size_t)m); double* Px = (double*)malloc(sizeof(double) * (size_t)n); double* Aty = (double*)malloc(sizeof(double) * (size_t)n);
  /* warm start as RacingMPC::solve gives it (racing_mpc.cpp:293-340: set_initial of X, U, dU from the reference / previous
   * solution; no dual initial values are set): x0 from (X_ref, U_ref, dU of U_ref), z0 = A x0, y0 = 0 */
  if (warm) {
    for (int i = 0; i < N; i++) for (int k = 0; k < 6; k++) x[6 * i + k] = p->Xref[6 * i + k] / scl[6 * i + k] / D[6 * i + k];
    for (int i = 0; i < NS; i++) for (int k = 0; k < 2; k++) {
      const int ou = 6 * N + 2 * i + k, od = 6 * N + 2 * NS + 2 * i + k;
      const double u = in->U_ref[2 * i + k], up = i ? in->U_ref[2 * (i - 1) + k] : in->u_ic[k];
      x[ou] = u / scl[ou] / D[ou]; x[od] = (u - up) / p->T[i] / scl[od] / D[od];
    }
    if (p->learning) { const int oL = 6 * N + 4 * NS + (p->soft_boundary ? 1 : 0); for (int k = 0; k < K; k++) x[oL + k] = (1.0 / K) / D[oL + k]; }
    for (int i = 0; i < m; i++) { double sacc = 0.0; const double* ar = Q.A + (size_t)i * n; for (int j = 0; j < n; j++) sacc += ar[j] * x[j]; z[i] = sacc; }
  }
  int need_factor = 1, iters = 0, solved = 0;
  double pri = 0.0, dua = 0.0;
  for (int it = 1; it <= max_it; it++) {
    if (need_factor) {
      for (int i = 0; i < m; i++) {
        const int eq = Q.l[i] == Q.u[i], loose = Q.l[i] <= -OSQP_INF && Q.u[i] >= OSQP_INF;
        rv[i] = loose ? 1e-6 : (eq ? 1e3 * rho : rho);
      }
      memset(KK, 0, sizeof(double) * (size_t)nk * nk);
      for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) KK[(size_t)i * nk + j] = Q.P[(size_t)i * n + j]; KK[(size_t)i * nk + i] += sigma; }
      for (int i = 0; i < m; i++) { for (int j = 0; j < n; j++) { KK[(size_t)(n + i) * nk + j] = Q.A[(size_t)i * n + j]; KK[(size_t)j * nk + n + i] = Q.A[(size_t)i * n + j]; } KK[(size_t)(n + i) * nk + n + i] = -1.0 / rv[i]; }
      if (lu_factor_(KK, nk, piv)) { st = ORC_NUMERIC; break; }
      need_factor = 0;
    }
    memcpy(xp, x, sizeof(double) * (size_t)n); memcpy(zp, z, sizeof(double) * (size_t)m);
    for (int i = 0; i < n; i++) rhs[i] = sigma * xp[i] - Q.q[i];
    for (int i = 0; i < m; i++) rhs[n + i] = zp[i] - y[i] / rv[i];
    lu_solve_(KK, nk, piv, rhs);
    for (int i = 0; i < n; i++) x[i] = alpha * rhs[i] + (1.0 - alpha) * xp[i];
    for (int i = 0; i < m; i++) {
      const double zt = zp[i] + (rhs[n + i] - y[i]) / rv[i];
      const double zr = alpha * zt + (1.0 - alpha) * zp[i];
      double zn = zr + y[i] / rv[i];
      zn = zn < Q.l[i] ? Q.l[i] : (zn > Q.u[i] ? Q.u[i] : zn);
      y[i] += rv[i] * (zr - zn);
      z[i] = zn;
    }
    iters = it;
    if (it % 25 == 0 || it == max_it || (rho_interval > 0 && it % rho_interval == 0)) {
      for (int i = 0; i < m; i++) { double s = 0.0; const double* ar = Q.A + (size_t)i * n; for (int j = 0; j < n; j++) s += ar[j] * x[j]; Ax[i] = s; }
      for (int i = 0; i < n; i++) { double s = 0.0; const double* pr = Q.P + (size_t)i * n; for (int j = 0; j < n; j++) s += pr[j] * x[j]; Px[i] = s; }
      for (int j = 0; j < n; j++) Aty[j] = 0.0;
      for (int i = 0; i < m; i++) { const double* ar = Q.A + (size_t)i * n; for (int j = 0; j < n; j++) Aty[j] += ar[j] * y[i]; }
      /* unscaled norms (scaled_termination = 0) */
      double pr_ = 0.0, nAx = 0.0, nz = 0.0, du_ = 0.0, nPx = 0.0, nAty = 0.0, nq = 0.0;
      for (int i = 0; i < m; i++) { pr_ = fmax(pr_, fabs(Ax[i] - z[i]) / E[i]); nAx = fmax(nAx, fabs(Ax[i]) / E[i]); nz = fmax(nz, fabs(z[i]) / E[i]); }
      for (int j = 0; j < n; j++) { du_ = fmax(du_, fabs(Px[j] + Q.q[j] + Aty[j]) / D[j]); nPx = fmax(nPx, fabs(Px[j]) / D[j]); nAty = fmax(nAty, fabs(Aty[j]) / D[j]); nq = fmax(nq, fabs(Q.q[j]) / D[j]); }
      du_ /= cs; nPx /= cs; nAty /= cs; nq /= cs;
      pri = pr_; dua = du_;
      if (it % 25 == 0 || it == max_it) {
        if (pr_ <= eps + eps * fmax(nAx, nz) && du_ <= eps + eps * fmax(nPx, fmax(nAty, nq))) { solved = 1; break; }
      }
      if (rho_interval > 0 && it % rho_interval == 0) {   /* adaptive rho on the scaled, normalised residuals */
        double sp = 0.0, sAx = 0.0, sz = 0.0, sd = 0.0, sPx = 0.0, sAty = 0.0, sq = 0.0;
        for (int i = 0; i < m; i++) { sp = fmax(sp, fabs(Ax[i] - z[i])); sAx = fmax(sAx, fabs(Ax[i])); sz = fmax(sz, fabs(z[i])); }
        for (int j = 0; j < n; j++) { sd = fmax(sd, fabs(Px[j] + Q.q[j] + Aty[j])); sPx = fmax(sPx, fabs(Px[j])); sAty = fmax(sAty, fabs(Aty[j])); sq = fmax(sq, fabs(Q.q[j])); }
        const double np_ = sp / (fmax(sAx, sz) + 1e-10), nd_ = sd / (fmax(sPx, fmax(sAty, sq)) + 1e-10);
        double rn = rho * sqrt(np_ / (nd_ + 1e-10));
        rn = rn < 1e-6 ? 1e-6 : (rn > 1e6 ? 1e6 : rn);
        if (rn > 5.0 * rho || rn < rho / 5.0) { rho = rn; need_factor = 1; }
      }
    }
  }
  int polished = 0;
  double ppri = NAN, pdua = NAN;
  double* xs = (double*)malloc(sizeof(double) * (size_t)n);
  memcpy(xs, x, sizeof(double) * (size_t)n);
  if (do_polish && st == ORC_OK) {
    /* ---- polish: active rows from the ADMM iterate, regularised KKT, 3 refinement steps */
    int* act = (int*)malloc(sizeof(int) * (size_t)m); double* bnd = (double*)malloc(sizeof(double) * (size_t)m);
    int na = 0;
    for (int i = 0; i < m; i++) {
      if (z[i] - Q.l[i] < -y[i]) { act[na] = i; bnd[na] = Q.l[i]; na++; }
      else if (Q.u[i] - z[i] < y[i]) { act[na] = i; bnd[na] = Q.u[i]; na++; }
    }
    const int np2 = n + na; const double delta = 1e-6;
    double* K0 = (double*)calloc((size_t)np2 * np2, sizeof(double)); double* K1 = (double*)malloc(sizeof(double) * (size_t)np2 * np2);
    int* pv2 = (int*)malloc(sizeof(int) * (size_t)np2);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) K0[(size_t)i * np2 + j] = Q.P[(size_t)i * n + j];
    for (int a = 0; a < na; a++) for (int j = 0; j < n; j++) { K0[(size_t)(n + a) * np2 + j] = Q.A[(size_t)act[a] * n + j]; K0[(size_t)j * np2 + n + a] = Q.A[(size_t)act[a] * n + j]; }
    memcpy(K1, K0, sizeof(double) * (size_t)np2 * np2);
    for (int i = 0; i < n; i++) K1[(size_t)i * np2 + i] += delta;
    for (int a = 0; a < na; a++) K1[(size_t)(n + a) * np2 + n + a] -= delta;
    if (!lu_factor_(K1, np2, pv2)) {
      double* b2 = (double*)malloc(sizeof(double) * (size_t)np2); double* s2 = (double*)calloc((size_t)np2, sizeof(double)); double* r2 = (double*)malloc(sizeof(double) * (size_t)np2);
      for (int i = 0; i < n; i++) b2[i] = -Q.q[i];
      for (int a = 0; a < na; a++) b2[n + a] = bnd[a];
      memcpy(s2, b2, sizeof(double) * (size_t)np2);
      lu_solve_(K1, np2, pv2, s2);
      for (int ref = 0; ref < 3; ref++) {   /* (K + dK) ds = b - K s */
        for (int i = 0; i < np2; i++) { double s = b2[i]; const double* kr = K0 + (size_t)i * np2; for (int j = 0; j < np2; j++) s -= kr[j] * s2[j]; r2[i] = s; }
        lu_solve_(K1, np2, pv2, r2);
        for (int i = 0; i < np2; i++) s2[i] += r2[i];
      }
      /* residuals of the polished point (unscaled), accepted only if they improve on ADMM's */
      double* yp = (double*)calloc((size_t)m, sizeof(double));
      for (int a = 0; a < na; a++) yp[act[a]] = s2[n + a];
      double pr_ = 0.0, du_ = 0.0;
      for (int i = 0; i < m; i++) {
        double s = 0.0; const double* ar = Q.A + (size_t)i * n; for (int j = 0; j < n; j++) s += ar[j] * s2[j];
        const double zc = s < Q.l[i] ? Q.l[i] : (s > Q.u[i] ? Q.u[i] : s);
        pr_ = fmax(pr_, fabs(s - zc) / E[i]);
      }
      for (int j = 0; j < n; j++) {
        double s = Q.q[j]; for (int i = 0; i < n; i++) s += Q.P[(size_t)j * n + i] * s2[i];
        for (int i = 0; i < m; i++) s += Q.A[(size_t)i * n + j] * yp[i];
        du_ = fmax(du_, fabs(s) / D[j]);
      }
      du_ /= cs;
      ppri = pr_; pdua = du_;
      if ((pr_ < pri && du_ < dua) || (pr_ < pri && dua < 1e-10) || (du_ < dua && pri < 1e-10)) { polished = 1; memcpy(xs, s2, sizeof(double) * (size_t)n); }
      free(b2); free(s2); free(r2); free(yp);
    }
    free(act); free(bnd); free(K0); free(K1); free(pv2);
  }
  /* ---- back to physical units: v = S D x */
  const int oU = 6 * N, oD = 6 * N + 2 * NS;
  int o = 6 * N + 4 * NS;
  for (int i = 0; i < 6 * N; i++) out->X[i] = xs[i] * D[i] * scl[i];
  for (int i = 0; i < 2 * NS; i++) { out->U[i] = xs[oU + i] * D[oU + i] * scl[oU + i]; out->dU[i] = xs[oD + i] * D[oD + i] * scl[oD + i]; }
  if (p->soft_boundary) { out->sigma_b = xs[o] * D[o]; o++; }
  if (p->learning && out->lambda) for (int k = 0; k < K; k++) out->lambda[k] = xs[o + k] * D[o + k];
  if (info) { info[0] = iters; info[1] = solved ? 0 : 1; info[2] = polished; info[3] = pri; info[4] = dua; info[5] = ppri; info[6] = pdua; info[7] = rho; }
  out->iters = iters; out->polished = polished; out->status = (st == ORC_OK && !solved) ? ORC_MAX_ITER : st;
  st = out->status;
  free(D); free(E); free(dt); free(et); free(rv); free(KK); free(piv); free(x); free(z); free(y); free(rhs); free(xp); free(zp); free(Ax); free(Px); free(Aty); free(xs);
  oqp_free(&Q); free(scl); free(p);
  return st;
}
