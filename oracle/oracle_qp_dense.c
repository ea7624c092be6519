/*
 * oracle_qp_dense.c -- CPU ORACLE (test infrastructure only; see lmpc_oracle.h).
 *
 * Dense, literal statement of the QP that RacingMPC builds with CasADi Opti("conic")
 * (reference src/mpc/racing_mpc/src/racing_mpc.cpp:31-202 constraints, :442-477 tracking
 * cost, :479-522 LMPC cost, :524-543 boundary) and of the per-tick data flow of
 * RacingMPC::solve (:209-372).  The reference hands this QP to OSQP with polish=true; a
 * successful polish returns the exact optimum of the QP, so the oracle computes that
 * optimum with two independent methods and certifies it:
 *   1. dense Mehrotra primal-dual interior point on (H,q,Ae,be,G,h), LU with partial pivoting;
 *   2. active-set polish: equality-constrained KKT solve on the active set found by (1),
 *      tiny regularisation + iterative refinement (what OSQP's polish does);
 *   3. KKT certificate (stationarity, primal/dual feasibility, complementarity) from the
 *      dense data.
 * Deviations from a literal transcription, none of which changes the feasible set or the
 * optimum: variables are unscaled (scale_x_/scale_u_ only precondition OSQP,
 * racing_mpc.cpp:36-37); two-sided rows are split into finite one-sided rows; the u box
 * and the actuator box on the same variable are merged into one box; box rows on x_0 are
 * replaced by a feasibility check because x_0 == x_ic is an equality (racing_mpc.cpp:199-201).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

static int finite_bound(double b) { return isfinite(b) && fabs(b) < 1e19; }

int orc_build_prob(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss,
                   const orc_step_in* in, orc_prob* p) {
  const int N = c->N;
  memset(p, 0, sizeof *p);
  p->N = N; p->learning = c->learning; p->K = c->learning ? c->num_ss_pts : 0;
  p->soft_boundary = c->q_boundary > 0.0;
  double hs = 0.0; for (int i = 0; i < 6; i++) hs += c->convex_hull_slack[i] * c->convex_hull_slack[i];
  p->hull_slack = hs > 0.0;                                       /* racing_mpc.cpp:493 */
  memcpy(p->x_ic, in->x_ic, sizeof p->x_ic);
  memcpy(p->u_ic, in->u_ic, sizeof p->u_ic);
  memcpy(p->Xref, in->X_ref, sizeof(double) * 6 * (size_t)N);
  for (int i = 0; i < N; i++)                                      /* racing_mpc.cpp:219-223 */
    p->Xref[6 * i] = orc_align_abscissa(p->Xref[6 * i], in->x_ic[0], in->total_length);
  for (int i = 0; i < N; i++) { p->bl[i] = in->bound_left[i]; p->br[i] = in->bound_right[i]; p->vref[i] = in->vel_ref[i]; }
  for (int i = 0; i < N - 1; i++) {                                /* racing_mpc.cpp:169-176 */
    p->T[i] = in->T_ref[i];
    orc_linearise(v, p->Xref + 6 * i, in->U_ref + 2 * i, in->curvatures[i], in->T_ref[i],
                  p->A + 36 * i, p->B + 12 * i, p->g + 6 * i, NULL);
  }
  p->margin = c->margin + v->chassis_b / 2.0;                      /* racing_mpc.cpp:531 */
  const double alo[2] = {v->Fb_max / 1000.0, -v->max_steer};       /* single_track_planar_model.cpp:114,120 */
  const double ahi[2] = {v->Fd_max / 1000.0, v->max_steer};
  for (int k = 0; k < 2; k++) {
    p->ulo[k] = fmax(c->u_min[k], alo[k]);
    p->uhi[k] = fmin(c->u_max[k], ahi[k]);
  }
  p->dlo[0] = v->Fb_max / 1000.0 / v->Tb; p->dhi[0] = v->Fd_max / 1000.0 / v->Td;   /* :146-151 */
  p->dlo[1] = -v->max_steer_rate;         p->dhi[1] = v->max_steer_rate;
  for (int k = 0; k < 6; k++)
    if (in->x_ic[k] > c->x_max[k] || in->x_ic[k] < c->x_min[k]) return ORC_INFEASIBLE_IC;
  if (!p->soft_boundary) {
    if (in->x_ic[1] > p->bl[0] - p->margin || in->x_ic[1] < p->br[0] + p->margin) return ORC_INFEASIBLE_IC;
  }
  if (c->learning) {
    if (p->K > ORC_KMAX) return ORC_NUMERIC;
    /* query at X_ref(:, -1) after alignment (racing_mpc.cpp:249-255), pad/truncate, J - J0 */
    const double qs = in->ss_query_point ? in->ss_query_point[0] : p->Xref[6 * (N - 1)];
    const double qe = in->ss_query_point ? in->ss_query_point[1] : p->Xref[6 * (N - 1) + 1];
    p->ss_count = ss ? orc_ss_query_padded(ss, qs, qe, p->K,
                                           c->num_ss_pts_per_lap, p->ssx, p->ssc) : 0;
    if (p->ss_count == 0) return ORC_NO_SAFE_SET;
  }
  return ORC_OK;
}

double orc_eval_cost(const orc_config* c, const orc_prob* p, const double* X, const double* U,
                     const double* dU, double sigma_b, const double* lambda, const double* sigma_h) {
  const int N = p->N;
  double cost = 0.0;
  if (p->soft_boundary) cost += c->q_boundary * sigma_b * sigma_b;          /* :539 */
  for (int i = 0; i < N - 1; i++) {                                         /* :507-511 / :451-466 */
    const double* u = U + 2 * i; const double* d = dU + 2 * i;
    cost += u[0] * (c->R[0] * u[0] + c->R[1] * u[1]) + u[1] * (c->R[2] * u[0] + c->R[3] * u[1]);
    cost += d[0] * (c->R_d[0] * d[0] + c->R_d[1] * d[1]) + d[1] * (c->R_d[2] * d[0] + c->R_d[3] * d[1]);
  }
  if (c->learning) {
    for (int k = 0; k < 6; k++) if (p->hull_slack) cost += c->convex_hull_slack[k] * sigma_h[k] * sigma_h[k];
    for (int k = 0; k < p->K; k++) cost += p->ssc[k] * lambda[k];           /* :504 */
  } else {
    for (int i = 0; i < N - 1; i++) {                                       /* :448-463 */
      const double* x = X + 6 * i; const double dv = x[3] - p->vref[i];
      cost += c->q_contour * x[1] * x[1] + c->q_heading * x[2] * x[2] + c->q_vel * dv * dv +
              c->q_vy * x[4] * x[4] + c->q_vyaw * x[5] * x[5];
    }
    const double* x = X + 6 * (N - 1); const double dv = x[3] - p->vref[N - 1];   /* :469-476 */
    cost += 10.0 * (c->q_contour * x[1] * x[1] + c->q_heading * x[2] * x[2] + c->q_vel * dv * dv);
  }
  return cost;
}

/* ---------------------------------------------------------------------------------- */
typedef struct {
  int n, me, mi;
  double *H, *q, *Ae, *be;      /* dense, row-major */
  int* gidx; double* gval;      /* mi rows x 2 entries (idx < 0 => unused) */
  double* h;
  /* variable offsets */
  int oX, oU, oD, oSb, oL, oSh;
} dqp;

static void dqp_free(dqp* q) { free(q->H); free(q->q); free(q->Ae); free(q->be); free(q->gidx); free(q->gval); free(q->h); }

static void add_row1(dqp* q, int* r, int i0, double v0, double h) {
  q->gidx[2 * *r] = i0; q->gval[2 * *r] = v0; q->gidx[2 * *r + 1] = -1; q->gval[2 * *r + 1] = 0.0; q->h[*r] = h; (*r)++;
}
static void add_row2(dqp* q, int* r, int i0, double v0, int i1, double v1, double h) {
  q->gidx[2 * *r] = i0; q->gval[2 * *r] = v0; q->gidx[2 * *r + 1] = i1; q->gval[2 * *r + 1] = v1; q->h[*r] = h; (*r)++;
}

static void dqp_build(const orc_config* c, const orc_prob* p, dqp* q) {
  const int N = p->N, K = p->K;
  int n = 6 * N + 4 * (N - 1);
  q->oX = 0; q->oU = 6 * N; q->oD = 6 * N + 2 * (N - 1);
  q->oSb = -1; q->oL = -1; q->oSh = -1;
  if (p->soft_boundary) { q->oSb = n; n += 1; }
  if (p->learning) { q->oL = n; n += K; if (p->hull_slack) { q->oSh = n; n += 6; } }
  int me = 6 + 6 * (N - 1) + 2 * (N - 1) + (p->learning ? 7 : 0);
  int mi_cap = (N - 1) * (12 + 4 + 4) + 2 * N + 1 + K;
  q->n = n; q->me = me;
  q->H = (double*)calloc((size_t)n * n, sizeof(double));
  q->q = (double*)calloc((size_t)n, sizeof(double));
  q->Ae = (double*)calloc((size_t)me * n, sizeof(double));
  q->be = (double*)calloc((size_t)me, sizeof(double));
  q->gidx = (int*)calloc((size_t)mi_cap * 2, sizeof(int));
  q->gval = (double*)calloc((size_t)mi_cap * 2, sizeof(double));
  q->h = (double*)calloc((size_t)mi_cap, sizeof(double));
#define HH(i, j) q->H[(size_t)(i) * n + (j)]
#define AE(r, j) q->Ae[(size_t)(r) * n + (j)]
  /* ---- cost: 1/2 v'Hv + q'v ---- */
  if (p->soft_boundary) HH(q->oSb, q->oSb) += 2.0 * c->q_boundary;
  for (int i = 0; i < N - 1; i++)
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 2; b++) {
        HH(q->oU + 2 * i + a, q->oU + 2 * i + b) += 2.0 * 0.5 * (c->R[2 * a + b] + c->R[2 * b + a]);
        HH(q->oD + 2 * i + a, q->oD + 2 * i + b) += 2.0 * 0.5 * (c->R_d[2 * a + b] + c->R_d[2 * b + a]);
      }
  if (p->learning) {
    if (p->hull_slack) for (int k = 0; k < 6; k++) HH(q->oSh + k, q->oSh + k) += 2.0 * c->convex_hull_slack[k];
    for (int k = 0; k < K; k++) q->q[q->oL + k] += p->ssc[k];
  } else {
    const double w[6] = {0.0, c->q_contour, c->q_heading, c->q_vel, c->q_vy, c->q_vyaw};
    for (int i = 0; i < N; i++) {
      const double sc = (i == N - 1) ? 10.0 : 1.0;
      for (int k = 1; k < 6; k++) {
        if (i == N - 1 && k >= 4) continue;
        HH(q->oX + 6 * i + k, q->oX + 6 * i + k) += 2.0 * sc * w[k];
      }
      q->q[q->oX + 6 * i + 3] += -2.0 * sc * c->q_vel * p->vref[i];
    }
  }
  /* ---- equalities ---- */
  int r = 0;
  for (int k = 0; k < 6; k++) { AE(r, q->oX + k) = 1.0; q->be[r] = p->x_ic[k]; r++; }   /* :199-201 */
  for (int i = 0; i < N - 1; i++) {                                                    /* :182 */
    for (int k = 0; k < 6; k++) {
      AE(r, q->oX + 6 * (i + 1) + k) = 1.0;
      for (int j = 0; j < 6; j++) AE(r, q->oX + 6 * i + j) -= p->A[36 * i + k + 6 * j];
      for (int j = 0; j < 2; j++) AE(r, q->oU + 2 * i + j) -= p->B[12 * i + k + 6 * j];
      q->be[r] = p->g[6 * i + k];
      r++;
    }
  }
  for (int i = 0; i < N - 1; i++) {                                                    /* :189-196 */
    for (int k = 0; k < 2; k++) {
      AE(r, q->oU + 2 * i + k) = -1.0;
      AE(r, q->oD + 2 * i + k) = p->T[i];
      if (i == 0) q->be[r] = -p->u_ic[k]; else { AE(r, q->oU + 2 * (i - 1) + k) = 1.0; q->be[r] = 0.0; }
      r++;
    }
  }
  if (p->learning) {
    for (int k = 0; k < K; k++) AE(r, q->oL + k) = 1.0;                                 /* :491 */
    q->be[r] = 1.0; r++;
    for (int k = 0; k < 6; k++) {                                                      /* :493-502 */
      AE(r, q->oX + 6 * (N - 1) + k) = 1.0;
      for (int j = 0; j < K; j++) AE(r, q->oL + j) = -p->ssx[6 * j + k];
      if (p->hull_slack) AE(r, q->oSh + k) = -1.0;
      q->be[r] = 0.0; r++;
    }
  }
  /* ---- inequalities  G v <= h ---- */
  int m = 0;
  for (int i = 0; i < N - 1; i++) {
    if (i > 0)                                                                          /* :147 */
      for (int k = 0; k < 6; k++) {
        if (finite_bound(c->x_max[k])) add_row1(q, &m, q->oX + 6 * i + k, 1.0, c->x_max[k]);
        if (finite_bound(c->x_min[k])) add_row1(q, &m, q->oX + 6 * i + k, -1.0, -c->x_min[k]);
      }
    for (int k = 0; k < 2; k++) {                                                       /* :148 + model rows */
      if (finite_bound(p->uhi[k])) add_row1(q, &m, q->oU + 2 * i + k, 1.0, p->uhi[k]);
      if (finite_bound(p->ulo[k])) add_row1(q, &m, q->oU + 2 * i + k, -1.0, -p->ulo[k]);
      if (finite_bound(p->dhi[k])) add_row1(q, &m, q->oD + 2 * i + k, 1.0, p->dhi[k]);
      if (finite_bound(p->dlo[k])) add_row1(q, &m, q->oD + 2 * i + k, -1.0, -p->dlo[k]);
    }
  }
  for (int i = 0; i < N; i++) {                                                         /* :529-541 */
    if (p->soft_boundary) {
      add_row2(q, &m, q->oX + 6 * i + 1, 1.0, q->oSb, -1.0, p->bl[i] - p->margin);
      add_row2(q, &m, q->oX + 6 * i + 1, -1.0, q->oSb, -1.0, -(p->br[i] + p->margin));
    } else if (i > 0) {
      add_row1(q, &m, q->oX + 6 * i + 1, 1.0, p->bl[i] - p->margin);
      add_row1(q, &m, q->oX + 6 * i + 1, -1.0, -(p->br[i] + p->margin));
    }
  }
  if (p->soft_boundary) add_row1(q, &m, q->oSb, -1.0, 0.0);                             /* :538 */
  for (int k = 0; k < K; k++) add_row1(q, &m, q->oL + k, -1.0, 0.0);                    /* :490 */
  q->mi = m;
#undef HH
#undef AE
}

/* LU with partial pivoting, in place, row-major n x n; returns 0 on success */
static int lu_factor(double* M, int n, int* piv) {
  for (int k = 0; k < n; k++) {
    int pk = k; double mx = fabs(M[(size_t)k * n + k]);
    for (int i = k + 1; i < n; i++) { double a = fabs(M[(size_t)i * n + k]); if (a > mx) { mx = a; pk = i; } }
    piv[k] = pk;
    if (mx == 0.0) return -1;
    if (pk != k) for (int j = 0; j < n; j++) { double t = M[(size_t)k * n + j]; M[(size_t)k * n + j] = M[(size_t)pk * n + j]; M[(size_t)pk * n + j] = t; }
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
static void lu_solve(const double* M, int n, const int* piv, double* b) {
  for (int k = 0; k < n; k++) { if (piv[k] != k) { double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; } }
  for (int i = 0; i < n; i++) { double s = b[i]; const double* ri = M + (size_t)i * n; for (int j = 0; j < i; j++) s -= ri[j] * b[j]; b[i] = s; }
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; const double* ri = M + (size_t)i * n; for (int j = i + 1; j < n; j++) s -= ri[j] * b[j]; b[i] = s / ri[i]; }
}

static double row_dot(const dqp* q, int j, const double* v) {
  double s = q->gval[2 * j] * v[q->gidx[2 * j]];
  if (q->gidx[2 * j + 1] >= 0) s += q->gval[2 * j + 1] * v[q->gidx[2 * j + 1]];
  return s;
}

/* max KKT residual of (v, pi, y) for the dense QP */
static double kkt_residual(const dqp* q, const double* v, const double* pi, const double* y) {
  const int n = q->n, me = q->me, mi = q->mi;
  double res = 0.0;
  double* rs = (double*)malloc(sizeof(double) * (size_t)n);
  for (int i = 0; i < n; i++) { double s = q->q[i]; const double* hr = q->H + (size_t)i * n; for (int j = 0; j < n; j++) s += hr[j] * v[j]; rs[i] = s; }
  for (int r = 0; r < me; r++) { const double* ar = q->Ae + (size_t)r * n; double e = -q->be[r]; for (int j = 0; j < n; j++) { rs[j] += ar[j] * pi[r]; e += ar[j] * v[j]; } if (fabs(e) > res) res = fabs(e); }
  for (int j = 0; j < mi; j++) {
    rs[q->gidx[2 * j]] += q->gval[2 * j] * y[j];
    if (q->gidx[2 * j + 1] >= 0) rs[q->gidx[2 * j + 1]] += q->gval[2 * j + 1] * y[j];
    const double sl = q->h[j] - row_dot(q, j, v);
    if (-sl > res) res = -sl;
    if (-y[j] > res) res = -y[j];
    if (fabs(sl * y[j]) > res) res = fabs(sl * y[j]);
  }
  for (int i = 0; i < n; i++) if (fabs(rs[i]) > res) res = fabs(rs[i]);
  free(rs);
  return res;
}

/* dense Mehrotra predictor-corrector.  v (n), pi (me), s,y (mi) are outputs. */
static int dense_ipm(const dqp* q, double* v, double* pi, double* s, double* y, int max_iter, double tol, int* iters_out) {
  const int n = q->n, me = q->me, mi = q->mi, nk = n + me;
  double* M = (double*)malloc(sizeof(double) * (size_t)nk * nk);
  double* rhs = (double*)malloc(sizeof(double) * (size_t)nk);
  double* rd = (double*)malloc(sizeof(double) * (size_t)n);
  double* re = (double*)malloc(sizeof(double) * (size_t)me);
  double* rp = (double*)malloc(sizeof(double) * (size_t)mi);
  double* dv = (double*)malloc(sizeof(double) * (size_t)nk);
  double* ds = (double*)malloc(sizeof(double) * (size_t)mi);
  double* dy = (double*)malloc(sizeof(double) * (size_t)mi);
  double* corr = (double*)calloc((size_t)mi, sizeof(double));
  int* piv = (int*)malloc(sizeof(int) * (size_t)nk);
  int status = ORC_MAX_ITER, it;
  for (int j = 0; j < mi; j++) { double sl = q->h[j] - row_dot(q, j, v); s[j] = sl > 1e-2 ? sl : 1e-2; y[j] = 1.0 / s[j]; }
  for (int r = 0; r < me; r++) pi[r] = 0.0;
  for (it = 0; it < max_iter; it++) {
    double mu = 0.0, rpn = 0.0, rdn = 0.0, ren = 0.0;
    for (int i = 0; i < n; i++) { double a = q->q[i]; const double* hr = q->H + (size_t)i * n; for (int j = 0; j < n; j++) a += hr[j] * v[j]; rd[i] = a; }
    for (int r = 0; r < me; r++) { const double* ar = q->Ae + (size_t)r * n; double e = -q->be[r]; for (int j = 0; j < n; j++) { rd[j] += ar[j] * pi[r]; e += ar[j] * v[j]; } re[r] = e; if (fabs(e) > ren) ren = fabs(e); }
    for (int j = 0; j < mi; j++) {
      rd[q->gidx[2 * j]] += q->gval[2 * j] * y[j];
      if (q->gidx[2 * j + 1] >= 0) rd[q->gidx[2 * j + 1]] += q->gval[2 * j + 1] * y[j];
      rp[j] = row_dot(q, j, v) + s[j] - q->h[j];
      if (fabs(rp[j]) > rpn) rpn = fabs(rp[j]);
      mu += s[j] * y[j];
    }
    mu /= (double)mi;
    for (int i = 0; i < n; i++) if (fabs(rd[i]) > rdn) rdn = fabs(rd[i]);
    if (mu < tol && rpn < tol && rdn < tol && ren < tol) { status = ORC_OK; break; }
    /* KKT matrix */
    memset(M, 0, sizeof(double) * (size_t)nk * nk);
    for (int i = 0; i < n; i++) memcpy(M + (size_t)i * nk, q->H + (size_t)i * n, sizeof(double) * (size_t)n);
    for (int r = 0; r < me; r++) for (int j = 0; j < n; j++) { double a = q->Ae[(size_t)r * n + j]; M[(size_t)(n + r) * nk + j] = a; M[(size_t)j * nk + n + r] = a; }
    for (int j = 0; j < mi; j++) {
      const double d = y[j] / s[j];
      const int i0 = q->gidx[2 * j], i1 = q->gidx[2 * j + 1]; const double a0 = q->gval[2 * j], a1 = q->gval[2 * j + 1];
      M[(size_t)i0 * nk + i0] += d * a0 * a0;
      if (i1 >= 0) { M[(size_t)i1 * nk + i1] += d * a1 * a1; M[(size_t)i0 * nk + i1] += d * a0 * a1; M[(size_t)i1 * nk + i0] += d * a0 * a1; }
    }
    if (lu_factor(M, nk, piv)) { status = ORC_NUMERIC; break; }
    double sigma = 0.0, alpha = 1.0;
    for (int pass = 0; pass < 2; pass++) {
      /* rhs: -(rd) + G'((rc - y rp)/s),  rc = s y - sigma mu + corr */
      for (int i = 0; i < n; i++) rhs[i] = -rd[i];
      for (int r = 0; r < me; r++) rhs[n + r] = -re[r];
      for (int j = 0; j < mi; j++) {
        const double rc = s[j] * y[j] - sigma * mu + (pass ? corr[j] : 0.0);
        const double t = (rc - y[j] * rp[j]) / s[j];
        rhs[q->gidx[2 * j]] += q->gval[2 * j] * t;
        if (q->gidx[2 * j + 1] >= 0) rhs[q->gidx[2 * j + 1]] += q->gval[2 * j + 1] * t;
      }
      memcpy(dv, rhs, sizeof(double) * (size_t)nk);
      lu_solve(M, nk, piv, dv);
      double amax = 1e300;
      for (int j = 0; j < mi; j++) {
        const double rc = s[j] * y[j] - sigma * mu + (pass ? corr[j] : 0.0);
        ds[j] = -rp[j] - row_dot(q, j, dv);
        dy[j] = (-rc - y[j] * ds[j]) / s[j];
        if (ds[j] < 0.0 && -s[j] / ds[j] < amax) amax = -s[j] / ds[j];
        if (dy[j] < 0.0 && -y[j] / dy[j] < amax) amax = -y[j] / dy[j];
      }
      if (pass == 0) {
        const double aa = amax < 1.0 ? amax : 1.0;
        double mua = 0.0;
        for (int j = 0; j < mi; j++) { mua += (s[j] + aa * ds[j]) * (y[j] + aa * dy[j]); corr[j] = ds[j] * dy[j]; }
        mua /= (double)mi;
        sigma = pow(mua / mu, 3.0);
      } else {
        alpha = 0.995 * amax; if (alpha > 1.0) alpha = 1.0;
      }
    }
    for (int i = 0; i < n; i++) v[i] += alpha * dv[i];
    for (int r = 0; r < me; r++) pi[r] += alpha * dv[n + r];
    for (int j = 0; j < mi; j++) { s[j] += alpha * ds[j]; y[j] += alpha * dy[j]; }
  }
  *iters_out = it;
  free(M); free(rhs); free(rd); free(re); free(rp); free(dv); free(ds); free(dy); free(corr); free(piv);
  return status;
}

/* Active-set polish (OSQP-style): returns 1 if accepted (v, pi, y overwritten). */
static int polish(const dqp* q, double* v, double* pi, const double* s, double* y) {
  const int n = q->n, me = q->me, mi = q->mi;
  int* act = (int*)malloc(sizeof(int) * (size_t)mi); int na = 0;
  for (int j = 0; j < mi; j++) if (y[j] > s[j]) act[na++] = j;
  const int nk = n + me + na;
  const double delta = 1e-9;
  double* M = (double*)calloc((size_t)nk * nk, sizeof(double));
  double* M0 = (double*)malloc(sizeof(double) * (size_t)nk * nk);
  double* rhs = (double*)malloc(sizeof(double) * (size_t)nk);
  double* sol = (double*)calloc((size_t)nk, sizeof(double));
  double* res = (double*)malloc(sizeof(double) * (size_t)nk);
  int* piv = (int*)malloc(sizeof(int) * (size_t)nk);
  for (int i = 0; i < n; i++) memcpy(M + (size_t)i * nk, q->H + (size_t)i * n, sizeof(double) * (size_t)n);
  for (int r = 0; r < me; r++) for (int j = 0; j < n; j++) { double a = q->Ae[(size_t)r * n + j]; M[(size_t)(n + r) * nk + j] = a; M[(size_t)j * nk + n + r] = a; }
  for (int a = 0; a < na; a++) {
    const int j = act[a];
    for (int e = 0; e < 2; e++) { const int idx = q->gidx[2 * j + e]; if (idx < 0) continue; M[(size_t)(n + me + a) * nk + idx] = q->gval[2 * j + e]; M[(size_t)idx * nk + n + me + a] = q->gval[2 * j + e]; }
  }
  memcpy(M0, M, sizeof(double) * (size_t)nk * nk);
  for (int i = 0; i < n; i++) M[(size_t)i * nk + i] += delta;
  for (int i = n; i < nk; i++) M[(size_t)i * nk + i] -= delta;
  for (int i = 0; i < n; i++) rhs[i] = -q->q[i];
  for (int r = 0; r < me; r++) rhs[n + r] = q->be[r];
  for (int a = 0; a < na; a++) rhs[n + me + a] = q->h[act[a]];
  int ok = lu_factor(M, nk, piv) == 0;
  if (ok) {
    for (int ref = 0; ref < 6; ref++) {            /* iterative refinement on the unregularised system */
      for (int i = 0; i < nk; i++) { double a = rhs[i]; const double* r0 = M0 + (size_t)i * nk; for (int j = 0; j < nk; j++) a -= r0[j] * sol[j]; res[i] = a; }
      lu_solve(M, nk, piv, res);
      for (int i = 0; i < nk; i++) sol[i] += res[i];
    }
    /* acceptance: inactive rows feasible, active multipliers non-negative */
    for (int a = 0; a < na && ok; a++) if (sol[n + me + a] < -1e-9) ok = 0;
    for (int j = 0; j < mi && ok; j++) if (row_dot(q, j, sol) - q->h[j] > 1e-9) ok = 0;
    for (int i = 0; i < nk && ok; i++) if (!isfinite(sol[i])) ok = 0;
  }
  if (ok) {
    memcpy(v, sol, sizeof(double) * (size_t)n);
    memcpy(pi, sol + n, sizeof(double) * (size_t)me);
    for (int j = 0; j < mi; j++) y[j] = 0.0;
    for (int a = 0; a < na; a++) y[act[a]] = sol[n + me + a] > 0.0 ? sol[n + me + a] : 0.0;
  }
  free(act); free(M); free(M0); free(rhs); free(sol); free(res); free(piv);
  return ok;
}

static void unpack(const orc_prob* p, const dqp* q, const double* v, orc_step_out* out) {
  const int N = p->N;
  memcpy(out->X, v + q->oX, sizeof(double) * 6 * (size_t)N);
  memcpy(out->U, v + q->oU, sizeof(double) * 2 * (size_t)(N - 1));
  memcpy(out->dU, v + q->oD, sizeof(double) * 2 * (size_t)(N - 1));
  out->sigma_b = q->oSb >= 0 ? v[q->oSb] : 0.0;
  if (p->learning) {
    if (out->lambda) memcpy(out->lambda, v + q->oL, sizeof(double) * (size_t)p->K);
    for (int k = 0; k < 6; k++) {
      double acc = v[q->oX + 6 * (N - 1) + k];
      for (int j = 0; j < p->K; j++) acc -= p->ssx[6 * j + k] * v[q->oL + j];
      out->sigma_h[k] = acc;
    }
    if (out->ss_x) memcpy(out->ss_x, p->ssx, sizeof(double) * 6 * (size_t)p->K);
    if (out->ss_cost) memcpy(out->ss_cost, p->ssc, sizeof(double) * (size_t)p->K);
  } else {
    memset(out->sigma_h, 0, sizeof out->sigma_h);
  }
}

int orc_step_dense(const orc_vehicle* vp, const orc_config* c, const orc_safe_set* ss,
                   const orc_step_in* in, orc_step_out* out) {
  orc_prob* p = (orc_prob*)malloc(sizeof *p);
  out->iters = 0; out->polished = 0; out->kkt = NAN; out->cost = NAN;
  int st = orc_build_prob(vp, c, ss, in, p);
  if (st != ORC_OK) { out->status = st; free(p); return st; }
  dqp q; memset(&q, 0, sizeof q);
  dqp_build(c, p, &q);
  const int N = p->N;
  double* v = (double*)calloc((size_t)q.n, sizeof(double));
  double* pi = (double*)calloc((size_t)q.me, sizeof(double));
  double* s = (double*)calloc((size_t)q.mi, sizeof(double));
  double* y = (double*)calloc((size_t)q.mi, sizeof(double));
  /* start: the linearisation point, uniform convex weights */
  memcpy(v + q.oX, p->Xref, sizeof(double) * 6 * (size_t)N);
  memcpy(v + q.oX, p->x_ic, sizeof(double) * 6);
  memcpy(v + q.oU, in->U_ref, sizeof(double) * 2 * (size_t)(N - 1));
  if (q.oSb >= 0) v[q.oSb] = 0.1;
  if (p->learning) for (int k = 0; k < p->K; k++) v[q.oL + k] = 1.0 / p->K;
  st = dense_ipm(&q, v, pi, s, y, c->max_iter > 0 ? c->max_iter : 100, c->tol > 0 ? c->tol : 1e-11, &out->iters);
  if (st == ORC_OK || st == ORC_MAX_ITER) out->polished = polish(&q, v, pi, s, y);
  out->kkt = kkt_residual(&q, v, pi, y);
  unpack(p, &q, v, out);
  out->cost = orc_eval_cost(c, p, out->X, out->U, out->dU, out->sigma_b, p->learning ? v + q.oL : NULL, out->sigma_h);
  if (st == ORC_MAX_ITER && out->kkt < 1e-8) st = ORC_OK;
  out->status = st;
  free(v); free(pi); free(s); free(y); dqp_free(&q); free(p);
  return st;
}

double orc_check_candidate(const orc_vehicle* vp, const orc_config* c, const orc_safe_set* ss,
                           const orc_step_in* in, const double* X, const double* U, const double* dU,
                           const double* lambda, double* cost_out, double* prim_inf_out) {
  orc_prob* p = (orc_prob*)malloc(sizeof *p);
  int st = orc_build_prob(vp, c, ss, in, p);
  if (st != ORC_OK) { free(p); if (cost_out) *cost_out = NAN; if (prim_inf_out) *prim_inf_out = INFINITY; return INFINITY; }
  dqp q; memset(&q, 0, sizeof q);
  dqp_build(c, p, &q);
  const int N = p->N;
  double* v = (double*)calloc((size_t)q.n, sizeof(double));
  memcpy(v + q.oX, X, sizeof(double) * 6 * (size_t)N);
  memcpy(v + q.oU, U, sizeof(double) * 2 * (size_t)(N - 1));
  memcpy(v + q.oD, dU, sizeof(double) * 2 * (size_t)(N - 1));
  double sigma_h[6] = {0, 0, 0, 0, 0, 0};
  /* the slack variables are implied by the candidate: smallest feasible sigma_b, exact sigma_h */
  double sb = 0.0;
  if (q.oSb >= 0) {
    for (int i = 0; i < N; i++) {
      const double ey = X[6 * i + 1];
      sb = fmax(sb, ey - (p->bl[i] - p->margin));
      sb = fmax(sb, (p->br[i] + p->margin) - ey);
    }
    v[q.oSb] = sb;
  }
  if (p->learning) {
    memcpy(v + q.oL, lambda, sizeof(double) * (size_t)p->K);
    for (int k = 0; k < 6; k++) {
      double acc = X[6 * (N - 1) + k];
      for (int j = 0; j < p->K; j++) acc -= p->ssx[6 * j + k] * lambda[j];
      sigma_h[k] = acc;
      if (q.oSh >= 0) v[q.oSh + k] = acc;
    }
  }
  double inf = 0.0;
  for (int r = 0; r < q.me; r++) { double e = -q.be[r]; const double* ar = q.Ae + (size_t)r * q.n; for (int j = 0; j < q.n; j++) e += ar[j] * v[j]; if (fabs(e) > inf) inf = fabs(e); }
  for (int j = 0; j < q.mi; j++) { double e = row_dot(&q, j, v) - q.h[j]; if (e > inf) inf = e; }
  const double cost = orc_eval_cost(c, p, X, U, dU, sb, lambda, sigma_h);
  if (cost_out) *cost_out = cost;
  if (prim_inf_out) *prim_inf_out = inf;
  free(v); dqp_free(&q); free(p);
  return inf;
}
