/*
 * oracle_port.c -- CPU ORACLE, "port" arm (test infrastructure only; see lmpc_oracle.h).
 *
 * A scalar-C, structure-exploiting solver for exactly the QP of oracle_qp_dense.c (the
 * reference's RacingMPC QP, racing_mpc.cpp:31-202,442-543): Mehrotra primal-dual interior
 * point whose Newton systems are solved stage-wise by a Riccati recursion over
 * z_i = (x_i, u_{i-1}) with control u_i, the global boundary slack carried as a second
 * right-hand-side column, and the safe-set simplex (lambda, sigma_h) eliminated at the
 * terminal stage through a 6x6 Woodbury system.  It is the CPU counterpart of the CUDA
 * kernel's algorithm, is validated against the dense oracle by tests/, and is what
 * bench.py times as `cpu_baseline` ("kind": "port").  It is never linked into the product.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_internal.h"

#define NS ORC_NMAX
/* sigma_b >= 0 carried as a FREE variable (same optimum for q_b > 0; removes a complementarity pair that is degenerate
 * whenever no boundary row is active) -- the rule of csrc/lmpc_qp_core.cuh (LMPC_FREE_THETA), mirrored */
#ifndef ORC_FREE_THETA
#define ORC_FREE_THETA 1
#endif
#define QMAX 9              /* 1 + max explicit (basic) safe-set columns */
#define PMAX 6               /* active-set refinement rounds of the polish */
#define MAXROW 22            /* per-stage row slots: 12 x-box + 4 u-box + 4 du-box + 2 boundary */

typedef struct {
  /* static row structure */
  int nxb, xb_c[12]; double xb_sg[12], xb_h[12];
  int RS;                                   /* row stride per stage = nxb + 10 */
  int act[NS * MAXROW];                     /* row active mask */
  double rh[NS * MAXROW];                   /* row bound h */
  double s[NS * MAXROW], y[NS * MAXROW], rp[NS * MAXROW], corr[NS * MAXROW], gv[NS * MAXROW];
  double ds[NS * MAXROW], dy[NS * MAXROW];
  double th, yth, corr_th, dth, dyth;       /* sigma_b and its bound row (slack == variable) */
  double lam[ORC_KMAX], ylam[ORC_KMAX], corr_lam[ORC_KMAX], dlam[ORC_KMAX], dylam[ORC_KMAX];
  double St[6 * ORC_KMAX], cen[6];          /* centred safe-set columns */
  double x[6 * NS], u[2 * NS], dx[6 * NS], du[2 * NS];
  /* per-stage assembled data */
  double hx[6 * NS], czx[6 * NS], czu[2 * NS], czth[NS], cw[2 * NS], E[3 * NS], ev[2 * NS], Uq[3 * NS];
  /* factorisation */
  double Sinv[3 * NS], Kz[16 * NS], kff1[2 * NS], kffth[2 * NS], Cwth[2 * NS];
  double Yxu_s[12 * NS];
  /* terminal */
  int nh, hidx[6]; double Einv[6];
  double PT[36], pT[6], Phi[36], Phia[6], nu0, nvec[6], kap, W[36], avec[6], om1;
  double sig[6];
  double om[ORC_KMAX], gl[ORC_KMAX], Dk[ORC_KMAX];
  int pact[NS * MAXROW], pact_th, pnb[ORC_KMAX];
  double xsave[6 * NS], usave[2 * NS], lsave[ORC_KMAX], thsave, ysave[NS * MAXROW], ylsave[ORC_KMAX], ythsave;   /* polish: active rows / sigma_b bound / non-basic columns */ int isB[ORC_KMAX], Bidx[QMAX], mB, S2piv[QMAX];
  double PhiC[6 * QMAX], S2[QMAX * QMAX], Xq[QMAX * 6], q0[QMAX];
  int m_total;
} port_ws;

static inline int RID(const port_ws* w, int i, int slot) { return i * MAXROW + slot; (void)w; }

static void s2_solve(const double* LU, const int* piv, int n, double* b) {
  for (int k = 0; k < n; k++) if (piv[k] != k) { double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
  for (int i = 0; i < n; i++) { double a = b[i]; for (int k = 0; k < i; k++) a -= LU[i * QMAX + k] * b[k]; b[i] = a; }
  for (int i = n - 1; i >= 0; i--) { double a = b[i]; for (int k = i + 1; k < n; k++) a -= LU[i * QMAX + k] * b[k]; b[i] = a / LU[i * QMAX + i]; }
}

static void sym2_inv(const double q[3], double inv[3]) { /* [q0 q1; q1 q2] */
  const double det = q[0] * q[2] - q[1] * q[1];
  const double r = 1.0 / det;
  inv[0] = q[2] * r; inv[1] = -q[1] * r; inv[2] = q[0] * r;
}

/* dense Cholesky solve helpers for the <=6x6 terminal system */
static int chol6(double* M, int n) { /* in place lower, row-major n x n (stride 6) */
  for (int j = 0; j < n; j++) {
    double d = M[j * 6 + j];
    for (int k = 0; k < j; k++) d -= M[j * 6 + k] * M[j * 6 + k];
    if (!(d > 0.0)) return -1;
    d = sqrt(d); M[j * 6 + j] = d;
    for (int i = j + 1; i < n; i++) {
      double a = M[i * 6 + j];
      for (int k = 0; k < j; k++) a -= M[i * 6 + k] * M[j * 6 + k];
      M[i * 6 + j] = a / d;
    }
  }
  return 0;
}
static void chol6_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; i++) { double a = b[i]; for (int k = 0; k < i; k++) a -= L[i * 6 + k] * b[k]; b[i] = a / L[i * 6 + i]; }
  for (int i = n - 1; i >= 0; i--) { double a = b[i]; for (int k = i + 1; k < n; k++) a -= L[k * 6 + i] * b[k]; b[i] = a / L[i * 6 + i]; }
}

int orc_step_port(const orc_vehicle* vp, const orc_config* c, const orc_safe_set* ss,
                  const orc_step_in* in, orc_step_out* out) {
  orc_prob* p = (orc_prob*)malloc(sizeof *p);
  out->iters = 0; out->polished = 0; out->kkt = NAN; out->cost = NAN;
  int st = orc_build_prob(vp, c, ss, in, p);
  if (st != ORC_OK) { out->status = st; free(p); return st; }
  port_ws* w = (port_ws*)calloc(1, sizeof *w);
  const int N = p->N, K = p->K, soft = p->soft_boundary, learn = p->learning;
  const double step_tol = c->tol > 0 ? c->tol : 1e-7;   /* interior-point stage tolerance (the polish follows) */
  const double tol = 0.1 * step_tol;                   /* complementarity level at which the polish takes over */
  const int max_iter = c->max_iter > 0 ? c->max_iter : 60;
  const double Rm[3] = {c->R[0], 0.5 * (c->R[1] + c->R[2]), c->R[3]};
  const double Rd[3] = {c->R_d[0], 0.5 * (c->R_d[1] + c->R_d[2]), c->R_d[3]};
  const double qb = c->q_boundary;
  const int MB = getenv("ORC_MB") ? atoi(getenv("ORC_MB")) : 4;   /* explicit (basic-candidate) safe-set columns */

  /* ---- static row structure ---- */
  w->nxb = 0;
  for (int k = 0; k < 6; k++) {
    if (isfinite(c->x_max[k]) && fabs(c->x_max[k]) < 1e19) { w->xb_c[w->nxb] = k; w->xb_sg[w->nxb] = 1.0; w->xb_h[w->nxb] = c->x_max[k]; w->nxb++; }
    if (isfinite(c->x_min[k]) && fabs(c->x_min[k]) < 1e19) { w->xb_c[w->nxb] = k; w->xb_sg[w->nxb] = -1.0; w->xb_h[w->nxb] = -c->x_min[k]; w->nxb++; }
  }
  const int oUB = 12, oDB = 16, oBD = 20;
  int m_total = 0;
  for (int i = 0; i < N; i++) {
    if (i >= 1 && i <= N - 2)
      for (int r = 0; r < w->nxb; r++) { w->act[RID(w, i, r)] = 1; w->rh[RID(w, i, r)] = w->xb_h[r]; }
    if (i <= N - 2)
      for (int k = 0; k < 2; k++) {
        const double hb[4] = {p->uhi[k], -p->ulo[k], p->dhi[k], -p->dlo[k]};
        const int sl[4] = {oUB + 2 * k, oUB + 2 * k + 1, oDB + 2 * k, oDB + 2 * k + 1};
        for (int q = 0; q < 4; q++) if (isfinite(hb[q]) && fabs(hb[q]) < 1e19) { w->act[RID(w, i, sl[q])] = 1; w->rh[RID(w, i, sl[q])] = hb[q]; }
      }
    if (soft || i >= 1) {
      w->act[RID(w, i, oBD)] = 1;     w->rh[RID(w, i, oBD)] = p->bl[i] - p->margin;
      w->act[RID(w, i, oBD + 1)] = 1; w->rh[RID(w, i, oBD + 1)] = -(p->br[i] + p->margin);
    }
  }
  for (int j = 0; j < N * MAXROW; j++) m_total += w->act[j];
  if (soft && !ORC_FREE_THETA) m_total += 1;
  m_total += K;
  w->m_total = m_total;

  /* ---- terminal (safe-set) constants ---- */
  w->nh = 0;
  if (learn) {
    for (int k = 0; k < 6; k++) {
      if (p->hull_slack && c->convex_hull_slack[k] == 0.0) continue;   /* free slack component: row vacuous */
      w->hidx[w->nh] = k; w->Einv[w->nh] = p->hull_slack ? 1.0 / (2.0 * c->convex_hull_slack[k]) : 0.0; w->nh++;
    }
    for (int k = 0; k < 6; k++) w->cen[k] = p->Xref[6 * (N - 1) + k];  /* centre = the query point */
    for (int j = 0; j < K; j++) for (int k = 0; k < 6; k++) w->St[6 * j + k] = p->ssx[6 * j + k] - w->cen[k];
  }
  const int nh = w->nh;

  /* ---- initial point: u = U_ref (clipped into the box), x = linear rollout from x_ic ---- */
  memcpy(w->x, p->x_ic, sizeof(double) * 6);
  for (int i = 0; i < N - 1; i++) {
    for (int k = 0; k < 2; k++) {
      double uu = in->U_ref[2 * i + k];
      if (uu > p->uhi[k]) uu = p->uhi[k];
      if (uu < p->ulo[k]) uu = p->ulo[k];
      w->u[2 * i + k] = uu;
    }
    for (int r = 0; r < 6; r++) {
      double a = p->g[6 * i + r];
      for (int k = 0; k < 6; k++) a += p->A[36 * i + r + 6 * k] * w->x[6 * i + k];
      for (int k = 0; k < 2; k++) a += p->B[12 * i + r + 6 * k] * w->u[2 * i + k];
      w->x[6 * (i + 1) + r] = a;
    }
  }
  /* A linear rollout that leaves the neighbourhood of the linearisation is useless as a start and would set the channel
   * scales of the stopping test (an explicit-Euler discretisation of the stiff lateral dynamics is unstable at
   * dt = 0.025: |x| grows to 1e5 over 19 stages).  Any channel beyond 100 x max(1, |x_ic|, |X_ref[N-1]|) => roll out
   * again with controls chosen stage by stage (2x2 least squares, clipped into the box) so that the velocity states
   * (v_x, v_y, omega) follow the chord x_ic -> X_ref[N-1]; the iterate stays dynamically feasible. */
  {
    int diverged = 0;
    for (int k = 0; k < 6; k++) {
      const double lim = 100.0 * fmax(1.0, fmax(fabs(p->x_ic[k]), fabs(p->Xref[6 * (N - 1) + k])));
      for (int i = 1; i < N; i++) if (!(fabs(w->x[6 * i + k]) <= lim)) diverged = 1;
    }
    if (diverged) {
      for (int i = 0; i < N - 1; i++) {
        const double* A = p->A + 36 * i; const double* B = p->B + 12 * i; const double* g = p->g + 6 * i;
        double r[3], M00 = 1e-12, M01 = 0.0, M11 = 1e-12, b0 = 0.0, b1 = 0.0;
        for (int q = 0; q < 3; q++) {
          const int c = 3 + q;
          double a = g[c] - (p->x_ic[c] + (p->Xref[6 * (N - 1) + c] - p->x_ic[c]) * ((double)(i + 1) / (double)(N - 1)));
          for (int k = 0; k < 6; k++) a += A[c + 6 * k] * w->x[6 * i + k];
          r[q] = a;
          M00 += B[c] * B[c]; M01 += B[c] * B[c + 6]; M11 += B[c + 6] * B[c + 6];
          b0 -= B[c] * a; b1 -= B[c + 6] * a;
        }
        (void)r;
        const double det = M00 * M11 - M01 * M01;
        double uu[2] = {(M11 * b0 - M01 * b1) / det, (M00 * b1 - M01 * b0) / det};
        for (int k = 0; k < 2; k++) {
          if (!(uu[k] <= p->uhi[k])) uu[k] = p->uhi[k];
          if (!(uu[k] >= p->ulo[k])) uu[k] = p->ulo[k];
          w->u[2 * i + k] = uu[k];
        }
        for (int rr = 0; rr < 6; rr++) {
          double a = g[rr];
          for (int k = 0; k < 6; k++) a += A[rr + 6 * k] * w->x[6 * i + k];
          for (int k = 0; k < 2; k++) a += B[rr + 6 * k] * w->u[2 * i + k];
          w->x[6 * (i + 1) + rr] = a;
        }
      }
    }
  }
  /* channel scales max(1, |channel|) of the parity metric, from the initial iterate */
  double chx[6] = {1, 1, 1, 1, 1, 1}, chu[2] = {1, 1}, chd[2] = {1, 1};
  for (int i = 0; i < N; i++) for (int k = 0; k < 6; k++) chx[k] = fmax(chx[k], fabs(w->x[6 * i + k]));
  for (int i = 0; i < N - 1; i++) for (int k = 0; k < 2; k++) {
    const double up = i ? w->u[2 * (i - 1) + k] : p->u_ic[k];
    chu[k] = fmax(chu[k], fabs(w->u[2 * i + k])); chd[k] = fmax(chd[k], fabs(w->u[2 * i + k] - up) / p->T[i]);
  }
  const double sfloor = getenv("ORC_SFLOOR") ? atof(getenv("ORC_SFLOOR")) : 1e-2;
  const double mu0 = getenv("ORC_MU0") ? atof(getenv("ORC_MU0")) : 0.1;
  const double th0 = getenv("ORC_TH0") ? atof(getenv("ORC_TH0")) : 0.01;
  w->th = th0;
  if (soft) {
    /* a start whose rollout leaves the track: the boundary slack sigma_b absorbs the violation from the beginning
     * (boundary rows strictly feasible at the start) instead of being dragged there by an infeasible-start crawl */
    const double thoff = getenv("ORC_THOFF") ? atof(getenv("ORC_THOFF")) : 0.1;
    double viol = -1e300;
    for (int i = 0; i < N; i++)
      for (int sl = 20; sl < MAXROW; sl++) {
        const int j = RID(w, i, sl);
        if (!w->act[j]) continue;
        const double gv = ((sl & 1) ? -1.0 : 1.0) * w->x[6 * i + 1] - w->rh[j];
        if (gv > viol) viol = gv;
      }
    if (thoff >= 0.0 && viol + thoff > w->th) w->th = viol + thoff;
  }
  w->yth = ORC_FREE_THETA ? 0.0 : mu0 / w->th;
  for (int j = 0; j < K; j++) { w->lam[j] = 1.0 / K; w->ylam[j] = mu0 * K; }
  double R0 = 1.0; /* size of the initial dual residual (pi0 = 0): bounds the tracked reduction */
  {
    /* rows: s = max(slack, 1e-2), y = 1/s */
    for (int i = 0; i < N; i++) {
      for (int sl = 0; sl < MAXROW; sl++) {
        const int j = RID(w, i, sl);
        if (!w->act[j]) continue;
        double gv;
        if (sl < 12) gv = w->xb_sg[sl] * w->x[6 * i + w->xb_c[sl]];
        else if (sl < 16) { const int k = (sl - 12) >> 1; gv = ((sl & 1) ? -1.0 : 1.0) * w->u[2 * i + k]; }
        else if (sl < 20) { const int k = (sl - 16) >> 1; const double up = i ? w->u[2 * (i - 1) + k] : p->u_ic[k]; gv = ((sl & 1) ? -1.0 : 1.0) * (w->u[2 * i + k] - up) / p->T[i]; }
        else gv = ((sl & 1) ? -1.0 : 1.0) * w->x[6 * i + 1] - (soft ? w->th : 0.0);
        const double slack = w->rh[j] - gv;
        w->s[j] = slack > sfloor ? slack : sfloor;
        w->y[j] = mu0 / w->s[j];
        if (w->y[j] > R0) R0 = w->y[j];
      }
    }
    for (int j = 0; j < K; j++) if (fabs(p->ssc[j]) > R0) R0 = fabs(p->ssc[j]);
    if (2.0 * qb * w->th > R0) R0 = 2.0 * qb * w->th;
  }
  double rho_d = 1.0, prev_stepn = 0.0, mu_m1 = 1e300, mu_m2 = 1e300, mu_m3 = 1e300;
  const double stall_frac = getenv("ORC_STALL") ? atof(getenv("ORC_STALL")) : 0.5;
  int it, status = ORC_MAX_ITER;

  /* Active-set polish (what OSQP's polish=true does for the reference, racing_mpc.cpp:90-95), as one
   * augmented-Lagrangian Newton step on the active set the interior point identified: active rows get the
   * weight rho and the gradient y + rho (g'v - h), inactive rows are dropped, basic safe-set columns are free. */
  const int do_polish = getenv("ORC_POLISH") ? atoi(getenv("ORC_POLISH")) : 1;
  const double prho = getenv("ORC_PRHO") ? atof(getenv("ORC_PRHO")) : 1e7;
  int polishing = 0, polish_tries = 0, numfail_polish = 0, prev_changed = 0;
  double step_tol2 = step_tol, tol2 = tol;
  for (it = 0; it < max_iter || polishing; it++) {
    /* ---------- residuals, mu ---------- */
    double mu = 0.0, rpn = 0.0;
    for (int i = 0; i < N; i++) {
      for (int sl = 0; sl < MAXROW; sl++) {
        const int j = RID(w, i, sl);
        if (!w->act[j]) continue;
        double gv;
        if (sl < 12) gv = w->xb_sg[sl] * w->x[6 * i + w->xb_c[sl]];
        else if (sl < 16) { const int k = (sl - 12) >> 1; gv = ((sl & 1) ? -1.0 : 1.0) * w->u[2 * i + k]; }
        else if (sl < 20) { const int k = (sl - 16) >> 1; const double up = i ? w->u[2 * (i - 1) + k] : p->u_ic[k]; gv = ((sl & 1) ? -1.0 : 1.0) * (w->u[2 * i + k] - up) / p->T[i]; }
        else gv = ((sl & 1) ? -1.0 : 1.0) * w->x[6 * i + 1] - (soft ? w->th : 0.0);
        w->gv[j] = gv;
        w->rp[j] = gv + w->s[j] - w->rh[j];
        if (fabs(w->rp[j]) > rpn) rpn = fabs(w->rp[j]);
        mu += w->s[j] * w->y[j];
      }
    }
    if (soft && !ORC_FREE_THETA) mu += w->th * w->yth;
    double rnu = -1.0;
    for (int j = 0; j < K; j++) { mu += w->lam[j] * w->ylam[j]; rnu += w->lam[j]; }
    if (!learn) rnu = 0.0;
    mu /= (double)m_total;
    if (!polishing && ((mu < tol2 && rpn < tol2 && rho_d * R0 < tol2 && fabs(rnu) < tol2) ||
                       (mu < 1e-13 && rpn < 1e-9 && rho_d * R0 < 1e-9 && fabs(rnu) < 1e-9))) {   /* or the floor of double precision */
      if (!do_polish) { status = ORC_OK; break; }
      polishing = 1;
    }
    /* stalled interior point (Mehrotra limit cycle at a badly centred iterate: mu has not halved over
     * three iterations although the iterate is primal feasible): let the active-set polish decide from here */
    if (!polishing && do_polish && polish_tries < 2 && it >= 3 && mu < 1e-5 && rpn < 1e-8 && fabs(rnu) < 1e-8 &&
        mu > stall_frac * mu_m3) polishing = 1;
    if (!polishing) { mu_m3 = mu_m2; mu_m2 = mu_m1; mu_m1 = mu; }
    if (polishing == 1) {   /* first polish round: classify from the interior-point iterate */
      memcpy(w->xsave, w->x, sizeof w->xsave); memcpy(w->usave, w->u, sizeof w->usave); memcpy(w->lsave, w->lam, sizeof w->lsave); w->thsave = w->th;
      memcpy(w->ysave, w->y, sizeof w->y); memcpy(w->ylsave, w->ylam, sizeof w->ylam); w->ythsave = w->yth; polish_tries++; prev_changed = 0;
      for (int j = 0; j < N * MAXROW; j++) w->pact[j] = w->act[j] && w->y[j] > w->s[j];
      w->pact_th = !ORC_FREE_THETA && soft && w->yth > w->th;
      for (int j = 0; j < K; j++) w->pnb[j] = !(w->lam[j] >= w->ylam[j]);
      for (int j = 0; j < N * MAXROW; j++) if (w->act[j] && !w->pact[j]) w->y[j] = 0.0;
      if (soft && !w->pact_th) w->yth = 0.0;
      for (int j = 0; j < K; j++) if (!w->pnb[j]) w->ylam[j] = 0.0;
    }
    if (polishing && learn) {   /* the explicit-column system holds at most MB free columns */
      int nb = 0;
      for (int j = 0; j < K; j++) if (!w->pnb[j]) nb++;
      if (nb > MB || nb < 1) goto polish_failed;
    }
    if (0) {
polish_failed:
      /* no consistent active set: restore the interior-point iterate; the first time, keep iterating with a
       * 100x tighter tolerance and try once more, the second time return the interior-point solution */
      memcpy(w->x, w->xsave, sizeof(double) * 6 * (size_t)N); memcpy(w->u, w->usave, sizeof(double) * 2 * (size_t)N);
      memcpy(w->lam, w->lsave, sizeof(double) * (size_t)(K > 0 ? K : 1)); w->th = w->thsave;
      memcpy(w->y, w->ysave, sizeof w->y); memcpy(w->ylam, w->ylsave, sizeof w->ylam); w->yth = w->ythsave;
      polishing = 0;
      if (numfail_polish) { status = ORC_NUMERIC; break; }
      if (polish_tries >= 2 || it >= max_iter) { status = ORC_INACCURATE; break; }
      step_tol2 *= 1e-2; tol2 *= 1e-2;
      continue;
    }

    /* hull residual sigma = x_{N-1} - c - St lam (the slack sigma_h by definition) */
    if (learn) {
      for (int a = 0; a < nh; a++) {
        const int k = w->hidx[a];
        double r = w->x[6 * (N - 1) + k] - w->cen[k];
        for (int j = 0; j < K; j++) r -= w->St[6 * j + k] * w->lam[j];
        w->sig[a] = r;
      }
    }

    double sigma = 0.0, alpha = 1.0, Pithth_keep = 0.0;
    for (int pass = 0; pass < (polishing ? 1 : 2); pass++) {
      const double smu = sigma * mu;
      /* ---------- assemble stage data ---------- */
      double Dthth = 0.0, cth = 0.0;
      if (soft && ORC_FREE_THETA) { Dthth = 2.0 * qb; cth = 2.0 * qb * w->th; }
      else if (soft) { Dthth = 2.0 * qb + w->yth / w->th; cth = 2.0 * qb * w->th - (smu - (pass ? w->corr_th : 0.0)) / w->th; }
      if (soft && polishing) {
        Dthth = 2.0 * qb; cth = 2.0 * qb * w->th;
        if (w->pact_th) { Dthth += prho; cth += -w->yth + prho * w->th; }
      }
      for (int i = 0; i < N; i++) {
        double* hx = w->hx + 6 * i; double* czx = w->czx + 6 * i;
        for (int k = 0; k < 6; k++) { hx[k] = 0.0; czx[k] = 0.0; }
        w->czth[i] = 0.0;
        if (!learn) { /* tracking cost (racing_mpc.cpp:448-476) */
          const double sc = (i == N - 1) ? 10.0 : 1.0;
          const double wq[6] = {0.0, c->q_contour, c->q_heading, c->q_vel, (i == N - 1) ? 0.0 : c->q_vy, (i == N - 1) ? 0.0 : c->q_vyaw};
          for (int k = 1; k < 6; k++) { hx[k] += 2.0 * sc * wq[k]; czx[k] += 2.0 * sc * wq[k] * (w->x[6 * i + k] - (k == 3 ? p->vref[i] : 0.0)); }
        }
        double dd[2] = {0, 0}, td[2] = {0, 0}, dub[2] = {0, 0}, tub[2] = {0, 0};
        for (int sl = 0; sl < MAXROW; sl++) {
          const int j = RID(w, i, sl);
          if (!w->act[j]) continue;
          double d = w->y[j] / w->s[j];
          double t = (smu - (pass ? w->corr[j] : 0.0)) / w->s[j] + d * w->rp[j];
          if (polishing) {
            if (w->pact[j]) { d = prho; t = w->y[j] + prho * (w->rp[j] - w->s[j]); } else { d = 0.0; t = 0.0; }
          }
          const double sg = (sl < 12) ? w->xb_sg[sl] : ((sl & 1) ? -1.0 : 1.0);
          if (sl < 12) { hx[w->xb_c[sl]] += d; czx[w->xb_c[sl]] += sg * t; }
          else if (sl < 16) { dub[(sl - 12) >> 1] += d; tub[(sl - 12) >> 1] += sg * t; }
          else if (sl < 20) { dd[(sl - 16) >> 1] += d; td[(sl - 16) >> 1] += sg * t; }
          else {
            hx[1] += d; czx[1] += sg * t;
            if (soft) { w->czth[i] += -sg * d; Dthth += d; cth += -t; }
          }
        }
        if (i <= N - 2) {
          const double T = p->T[i], iT = 1.0 / T;
          const double* u = w->u + 2 * i;
          const double dcur[2] = {(u[0] - (i ? w->u[2 * (i - 1)] : p->u_ic[0])) * iT, (u[1] - (i ? w->u[2 * (i - 1) + 1] : p->u_ic[1])) * iT};
          double* E = w->E + 3 * i; double* ev = w->ev + 2 * i; double* Uq = w->Uq + 3 * i; double* cw = w->cw + 2 * i;
          E[0] = (2.0 * Rd[0] + dd[0]) * iT * iT; E[1] = 2.0 * Rd[1] * iT * iT; E[2] = (2.0 * Rd[2] + dd[1]) * iT * iT;
          ev[0] = (2.0 * (Rd[0] * dcur[0] + Rd[1] * dcur[1]) + td[0]) * iT;
          ev[1] = (2.0 * (Rd[1] * dcur[0] + Rd[2] * dcur[1]) + td[1]) * iT;
          Uq[0] = 2.0 * Rm[0] + dub[0]; Uq[1] = 2.0 * Rm[1]; Uq[2] = 2.0 * Rm[2] + dub[1];
          cw[0] = 2.0 * (Rm[0] * u[0] + Rm[1] * u[1]) + tub[0] + ev[0];
          cw[1] = 2.0 * (Rm[1] * u[0] + Rm[2] * u[1]) + tub[1] + ev[1];
          w->czu[2 * i] = -ev[0]; w->czu[2 * i + 1] = -ev[1];
        } else { w->czu[2 * i] = 0.0; w->czu[2 * i + 1] = 0.0; }
      }

      /* ---------- terminal block ---------- */
      double Pm[64]; /* P_{i+1}, 8x8 row-major */
      double l1[8], lth[8];
      memset(Pm, 0, sizeof Pm);
      for (int k = 0; k < 6; k++) { Pm[8 * k + k] = w->hx[6 * (N - 1) + k]; l1[k] = w->czx[6 * (N - 1) + k]; lth[k] = 0.0; }
      l1[6] = l1[7] = lth[6] = lth[7] = 0.0;
      lth[1] = w->czth[N - 1];
      if (learn) {
        /* Safe-set simplex block.  lambda_k with small Omega_k = lambda_k/y_k (non-basic) are
         * eliminated through the 6x6 Woodbury system; the MB columns with the largest Omega
         * (the basic candidates, whose Omega -> inf as mu -> 0) stay explicit unknowns together
         * with the simplex multiplier nu and are solved by a pivoted (1+MB)x(1+MB) LU, which is
         * what keeps the elimination accurate down to mu ~ 1e-12. */
        double bvec[6] = {0, 0, 0, 0, 0, 0}, omg = 0.0;
        if (pass == 0) {
          memset(w->W, 0, sizeof w->W); memset(w->avec, 0, sizeof w->avec); w->om1 = 0.0;
          for (int j = 0; j < K; j++) {
            w->om[j] = w->lam[j] / w->ylam[j]; w->isB[j] = 0; w->Dk[j] = w->ylam[j] / w->lam[j];
            if (polishing) { const int basic = !w->pnb[j]; w->Dk[j] = basic ? 0.0 : prho; w->om[j] = basic ? 1e300 : 1.0 / prho; }
          }
          w->mB = MB < K ? MB : K;
          for (int q = 0; q < w->mB; q++) {            /* top-MB by Omega, ties -> lowest index */
            int best = -1; double bv = -1.0;
            for (int j = 0; j < K; j++) if (!w->isB[j] && w->om[j] > bv) { bv = w->om[j]; best = j; }
            if (best < 0) best = 0;                    /* NaN weights (an iterate that blew up; it ends as NUMERIC): no winner, no address */
            w->Bidx[q] = best; w->isB[best] = 1;
          }
        }
        const int mB = w->mB, nq = 1 + mB;
        for (int j = 0; j < K; j++) {
          double gl = p->ssc[j] - (smu - (pass ? w->corr_lam[j] : 0.0)) / w->lam[j];
          if (polishing) gl = (w->Dk[j] > 0.0) ? p->ssc[j] - w->ylam[j] + prho * w->lam[j] : p->ssc[j];
          w->gl[j] = gl;
          if (w->isB[j]) continue;
          const double om = w->om[j];
          omg += om * gl;
          for (int a = 0; a < nh; a++) {
            const double sa = w->St[6 * j + w->hidx[a]];
            bvec[a] += sa * om * gl;
            if (pass == 0) { w->avec[a] += sa * om; for (int b = 0; b <= a; b++) w->W[6 * a + b] += om * sa * w->St[6 * j + w->hidx[b]]; }
          }
          if (pass == 0) w->om1 += om;
        }
        if (pass == 0) {
          double Lc[36];
          for (int a = 0; a < nh; a++) for (int b = 0; b <= a; b++) { Lc[6 * a + b] = w->W[6 * a + b] + (a == b ? w->Einv[a] : 0.0); }
          if (chol6(Lc, nh)) { if (polishing) goto polish_failed; status = ORC_NUMERIC; goto done; }
          for (int col = 0; col < nh; col++) { double e[6] = {0, 0, 0, 0, 0, 0}; e[col] = 1.0; chol6_solve(Lc, nh, e); for (int a = 0; a < nh; a++) w->Phi[6 * a + col] = e[a]; }
          /* C = [-a_N, St_B]  (nh x nq);  PhiC = Phi C */
          double Cm[6 * QMAX];
          for (int a = 0; a < nh; a++) { Cm[a * QMAX] = -w->avec[a]; for (int q = 0; q < mB; q++) Cm[a * QMAX + 1 + q] = w->St[6 * w->Bidx[q] + w->hidx[a]]; }
          for (int a = 0; a < nh; a++) for (int q = 0; q < nq; q++) { double s2 = 0.0; for (int b = 0; b < nh; b++) s2 += w->Phi[6 * a + b] * Cm[b * QMAX + q]; w->PhiC[a * QMAX + q] = s2; }
          /* S2 = Z - C' Phi C */
          for (int q = 0; q < nq; q++) for (int r = 0; r < nq; r++) {
            double z = 0.0;
            if (q == 0 && r == 0) z = w->om1; else if (q == 0 || r == 0) z = -1.0; else if (q == r) z = -w->Dk[w->Bidx[q - 1]];
            double s2 = 0.0; for (int a = 0; a < nh; a++) s2 += Cm[a * QMAX + q] * w->PhiC[a * QMAX + r];
            w->S2[q * QMAX + r] = z - s2;
          }
          /* LU with partial pivoting of S2 (nq x nq) */
          for (int k = 0; k < nq; k++) {
            int pk = k; double mx = fabs(w->S2[k * QMAX + k]);
            for (int i2 = k + 1; i2 < nq; i2++) if (fabs(w->S2[i2 * QMAX + k]) > mx) { mx = fabs(w->S2[i2 * QMAX + k]); pk = i2; }
            w->S2piv[k] = pk;
            if (mx == 0.0) { if (polishing) goto polish_failed; status = ORC_NUMERIC; goto done; }
            if (pk != k) for (int c2 = 0; c2 < nq; c2++) { double t = w->S2[k * QMAX + c2]; w->S2[k * QMAX + c2] = w->S2[pk * QMAX + c2]; w->S2[pk * QMAX + c2] = t; }
            for (int i2 = k + 1; i2 < nq; i2++) { const double f = w->S2[i2 * QMAX + k] / w->S2[k * QMAX + k]; w->S2[i2 * QMAX + k] = f; for (int c2 = k + 1; c2 < nq; c2++) w->S2[i2 * QMAX + c2] -= f * w->S2[k * QMAX + c2]; }
          }
          /* X = S2^{-1} PhiC'  (nq x nh);  PT = Phi + PhiC X */
          for (int a = 0; a < nh; a++) {
            double col[QMAX]; for (int q = 0; q < nq; q++) col[q] = w->PhiC[a * QMAX + q];
            s2_solve(w->S2, w->S2piv, nq, col);
            for (int q = 0; q < nq; q++) w->Xq[q * 6 + a] = col[q];
          }
          for (int a = 0; a < nh; a++) for (int b = 0; b < nh; b++) { double s2 = w->Phi[6 * a + b]; for (int q = 0; q < nq; q++) s2 += w->PhiC[a * QMAX + q] * w->Xq[q * 6 + b]; w->PT[6 * a + b] = s2; }
        }
        /* right-hand side: r1 = sigma + b_N, r2 = (rnu - omg_N, g_B);  q0 = S2^{-1}(r2 - PhiC' r1) */
        double r1[6], q0[QMAX];
        for (int a = 0; a < nh; a++) r1[a] = w->sig[a] + bvec[a];
        for (int q = 0; q < nq; q++) {
          double r2 = (q == 0) ? (rnu - omg) : w->gl[w->Bidx[q - 1]];
          for (int a = 0; a < nh; a++) r2 -= w->PhiC[a * QMAX + q] * r1[a];
          q0[q] = r2;
        }
        s2_solve(w->S2, w->S2piv, nq, q0);
        for (int q = 0; q < nq; q++) w->q0[q] = q0[q];
        for (int a = 0; a < nh; a++) {
          double s2 = 0.0;
          for (int b = 0; b < nh; b++) s2 += w->Phi[6 * a + b] * r1[b];
          for (int q = 0; q < nq; q++) s2 -= w->PhiC[a * QMAX + q] * q0[q];
          w->pT[a] = s2;
        }
        for (int a = 0; a < nh; a++) {
          l1[w->hidx[a]] += w->pT[a];
          for (int b = 0; b < nh; b++) Pm[8 * w->hidx[a] + w->hidx[b]] += w->PT[6 * a + b];
        }
      }

      /* ---------- backward Riccati sweep ---------- */
      double Pi1th = cth, Pithth = Dthth;
      for (int i = N - 2; i >= 0; i--) {
        const double* A = p->A + 36 * i; const double* B = p->B + 12 * i;
        const double* E = w->E + 3 * i;
        double* Kz = w->Kz + 16 * i; double* Sinv = w->Sinv + 3 * i; double* Yxu = w->Yxu_s + 12 * i;
        if (pass == 0) {
          double MA[36], MB[12], Yxx[36], Yuu[3];
          for (int r = 0; r < 6; r++) {
            for (int cc = 0; cc < 6; cc++) { double a = 0.0; for (int k = 0; k < 6; k++) a += Pm[8 * r + k] * A[k + 6 * cc]; MA[6 * r + cc] = a; }
            for (int cc = 0; cc < 2; cc++) { double a = Pm[8 * r + 6 + cc]; for (int k = 0; k < 6; k++) a += Pm[8 * r + k] * B[k + 6 * cc]; MB[2 * r + cc] = a; }
          }
          for (int r = 0; r < 6; r++) {
            for (int cc = 0; cc < 6; cc++) { double a = 0.0; for (int k = 0; k < 6; k++) a += A[k + 6 * r] * MA[6 * k + cc]; Yxx[6 * r + cc] = a; }
            for (int cc = 0; cc < 2; cc++) { double a = 0.0; for (int k = 0; k < 6; k++) a += A[k + 6 * r] * MB[2 * k + cc]; Yxu[2 * r + cc] = a; }
          }
          for (int r = 0; r < 2; r++)
            for (int cc = r; cc < 2; cc++) {
              double a = Pm[8 * (6 + r) + 6 + cc];
              for (int k = 0; k < 6; k++) a += B[k + 6 * r] * MB[2 * k + cc] + Pm[8 * k + 6 + r] * B[k + 6 * cc];
              Yuu[r + cc] = a;
            }
          const double* Uq = w->Uq + 3 * i;
          const double Qww[3] = {Yuu[0] + E[0] + Uq[0], Yuu[1] + E[1] + Uq[1], Yuu[2] + E[2] + Uq[2]};
          if (!(Qww[0] > 0.0) || !(Qww[0] * Qww[2] - Qww[1] * Qww[1] > 0.0)) {
            /* numerical floor of the barrier-weighted recursion: accept the iterate if already converged enough */
            if (polishing) goto polish_failed;
            if (do_polish && !numfail_polish && mu < 1e-5) { numfail_polish = 1; polishing = 1; polish_tries = 1; it--; goto next_trip; }
            status = ORC_NUMERIC; goto done;
          }
          sym2_inv(Qww, Sinv);
          /* Qzw = [Yxu; -E],  Kz = Sinv Qzw' (2x8) */
          double Qzw[16];
          for (int r = 0; r < 6; r++) { Qzw[2 * r] = Yxu[2 * r]; Qzw[2 * r + 1] = Yxu[2 * r + 1]; }
          Qzw[12] = -E[0]; Qzw[13] = -E[1]; Qzw[14] = -E[1]; Qzw[15] = -E[2];
          for (int r = 0; r < 8; r++) { Kz[r] = Sinv[0] * Qzw[2 * r] + Sinv[1] * Qzw[2 * r + 1]; Kz[8 + r] = Sinv[1] * Qzw[2 * r] + Sinv[2] * Qzw[2 * r + 1]; }
          /* RHS columns first (they need P_{i+1} only through l) -- computed below; now P_i */
          double Pn[64];
          for (int r = 0; r < 8; r++)
            for (int cc = 0; cc < 8; cc++) {
              double qzz = 0.0;
              if (r < 6 && cc < 6) qzz = Yxx[6 * r + cc] + (r == cc ? w->hx[6 * i + r] : 0.0);
              else if (r >= 6 && cc >= 6) qzz = E[(r - 6) + (cc - 6)];
              Pn[8 * r + cc] = qzz - (Qzw[2 * r] * Kz[cc] + Qzw[2 * r + 1] * Kz[8 + cc]);
            }
          { /* P_uu = E - E Sinv E has catastrophic cancellation once the rate-bound barrier weights in E
             * dominate; use the product form  P_uu = Q Sinv E  with Q = Qww - E  (no subtraction), symmetrised. */
            const double Q[3] = {Yuu[0] + Uq[0], Yuu[1] + Uq[1], Yuu[2] + Uq[2]};
            const double SE[4] = {Sinv[0] * E[0] + Sinv[1] * E[1], Sinv[0] * E[1] + Sinv[1] * E[2],
                                  Sinv[1] * E[0] + Sinv[2] * E[1], Sinv[1] * E[1] + Sinv[2] * E[2]};
            const double p00 = Q[0] * SE[0] + Q[1] * SE[2], p01 = Q[0] * SE[1] + Q[1] * SE[3];
            const double p10 = Q[1] * SE[0] + Q[2] * SE[2], p11 = Q[1] * SE[1] + Q[2] * SE[3];
            Pn[8 * 6 + 6] = p00; Pn[8 * 7 + 7] = p11; Pn[8 * 6 + 7] = Pn[8 * 7 + 6] = 0.5 * (p01 + p10);
          }
          /* RHS theta column */
          double ax[6], bw[2];
          for (int r = 0; r < 6; r++) { double a = 0.0; for (int k = 0; k < 6; k++) a += A[k + 6 * r] * lth[k]; ax[r] = a; }
          for (int r = 0; r < 2; r++) { double a = lth[6 + r]; for (int k = 0; k < 6; k++) a += B[k + 6 * r] * lth[k]; bw[r] = a; }
          w->Cwth[2 * i] = bw[0]; w->Cwth[2 * i + 1] = bw[1];
          double* kf = w->kffth + 2 * i;
          kf[0] = Sinv[0] * bw[0] + Sinv[1] * bw[1]; kf[1] = Sinv[1] * bw[0] + Sinv[2] * bw[1];
          for (int r = 0; r < 6; r++) lth[r] = (r == 1 ? w->czth[i] : 0.0) + ax[r] - (Yxu[2 * r] * kf[0] + Yxu[2 * r + 1] * kf[1]);
          lth[6] = E[0] * kf[0] + E[1] * kf[1]; lth[7] = E[1] * kf[0] + E[2] * kf[1];
          Pithth -= bw[0] * kf[0] + bw[1] * kf[1];
          /* (l1 update below uses the old Pm only through l1, so Pm may be replaced now) */
          memcpy(Pm, Pn, sizeof Pm);
        }
        /* RHS "1" column (both passes) */
        {
          double ax[6], Cw[2];
          for (int r = 0; r < 6; r++) { double a = 0.0; for (int k = 0; k < 6; k++) a += A[k + 6 * r] * l1[k]; ax[r] = a; }
          for (int r = 0; r < 2; r++) { double a = l1[6 + r]; for (int k = 0; k < 6; k++) a += B[k + 6 * r] * l1[k]; Cw[r] = a + w->cw[2 * i + r]; }
          double* kf = w->kff1 + 2 * i;
          kf[0] = Sinv[0] * Cw[0] + Sinv[1] * Cw[1]; kf[1] = Sinv[1] * Cw[0] + Sinv[2] * Cw[1];
          for (int r = 0; r < 6; r++) l1[r] = w->czx[6 * i + r] + ax[r] - (Yxu[2 * r] * kf[0] + Yxu[2 * r + 1] * kf[1]);
          l1[6] = w->czu[2 * i] + E[0] * kf[0] + E[1] * kf[1];
          l1[7] = w->czu[2 * i + 1] + E[1] * kf[0] + E[2] * kf[1];
          Pi1th -= w->Cwth[2 * i] * kf[0] + w->Cwth[2 * i + 1] * kf[1];
        }
      }
      /* the theta-theta curvature only changes with the factorisation (pass 0) */
      if (pass == 0) Pithth_keep = Pithth; else Pithth = Pithth_keep;

      /* ---------- scalar theta, forward sweep ---------- */
      w->dth = soft ? -Pi1th / Pithth : 0.0;
      for (int k = 0; k < 6; k++) w->dx[k] = 0.0;
      for (int i = 0; i < N - 1; i++) {
        const double* A = p->A + 36 * i; const double* B = p->B + 12 * i; const double* Kz = w->Kz + 16 * i;
        double dz[8];
        for (int k = 0; k < 6; k++) dz[k] = w->dx[6 * i + k];
        dz[6] = i ? w->du[2 * (i - 1)] : 0.0; dz[7] = i ? w->du[2 * (i - 1) + 1] : 0.0;
        for (int r = 0; r < 2; r++) {
          double a = -w->kff1[2 * i + r] - w->kffth[2 * i + r] * w->dth;
          for (int k = 0; k < 8; k++) a -= Kz[8 * r + k] * dz[k];
          w->du[2 * i + r] = a;
        }
        for (int r = 0; r < 6; r++) {
          double a = 0.0;
          for (int k = 0; k < 6; k++) a += A[r + 6 * k] * w->dx[6 * i + k];
          for (int k = 0; k < 2; k++) a += B[r + 6 * k] * w->du[2 * i + k];
          w->dx[6 * (i + 1) + r] = a;
        }
      }
      /* ---------- terminal directions ---------- */
      if (learn) {
        double e[6], qv[QMAX];
        const int mB = w->mB, nq = 1 + mB;
        for (int a = 0; a < nh; a++) {
          double s2 = w->pT[a];
          for (int b = 0; b < nh; b++) s2 += w->PT[6 * a + b] * w->dx[6 * (N - 1) + w->hidx[b]];
          e[a] = s2;
        }
        for (int q = 0; q < nq; q++) { double s2 = w->q0[q]; for (int a = 0; a < nh; a++) s2 -= w->Xq[q * 6 + a] * w->dx[6 * (N - 1) + w->hidx[a]]; qv[q] = s2; }
        const double nu = qv[0];
        for (int j = 0; j < K; j++) {
          if (!w->isB[j]) {
            double se = 0.0; for (int a = 0; a < nh; a++) se += w->St[6 * j + w->hidx[a]] * e[a];
            w->dlam[j] = w->om[j] * (se - w->gl[j] - nu);
          }
        }
        for (int q = 0; q < mB; q++) w->dlam[w->Bidx[q]] = qv[1 + q];
        for (int j = 0; j < K; j++) {
          const double tl = (smu - (pass ? w->corr_lam[j] : 0.0)) / w->lam[j];
          w->dylam[j] = tl - w->ylam[j] - w->dlam[j] * (w->ylam[j] / w->lam[j]);
        }
      }
      if (soft && !ORC_FREE_THETA) { const double tt = (smu - (pass ? w->corr_th : 0.0)) / w->th; w->dyth = tt - w->yth - (w->yth / w->th) * w->dth; }
      else w->dyth = 0.0;
      /* ---------- row directions and step length ---------- */
      double amax = 1e300;
      for (int i = 0; i < N; i++) {
        for (int sl = 0; sl < MAXROW; sl++) {
          const int j = RID(w, i, sl);
          if (!w->act[j]) continue;
          double dg;
          if (sl < 12) dg = w->xb_sg[sl] * w->dx[6 * i + w->xb_c[sl]];
          else if (sl < 16) { const int k = (sl - 12) >> 1; dg = ((sl & 1) ? -1.0 : 1.0) * w->du[2 * i + k]; }
          else if (sl < 20) { const int k = (sl - 16) >> 1; const double up = i ? w->du[2 * (i - 1) + k] : 0.0; dg = ((sl & 1) ? -1.0 : 1.0) * (w->du[2 * i + k] - up) / p->T[i]; }
          else dg = ((sl & 1) ? -1.0 : 1.0) * w->dx[6 * i + 1] - (soft ? w->dth : 0.0);
          const double rc = w->s[j] * w->y[j] - smu + (pass ? w->corr[j] : 0.0);
          w->ds[j] = -w->rp[j] - dg;
          w->dy[j] = (-rc - w->y[j] * w->ds[j]) / w->s[j];
          if (w->ds[j] < 0.0 && -w->s[j] / w->ds[j] < amax) amax = -w->s[j] / w->ds[j];
          if (w->dy[j] < 0.0 && -w->y[j] / w->dy[j] < amax) amax = -w->y[j] / w->dy[j];
        }
      }
      if (soft && !ORC_FREE_THETA) {
        if (w->dth < 0.0 && -w->th / w->dth < amax) amax = -w->th / w->dth;
        if (w->dyth < 0.0 && -w->yth / w->dyth < amax) amax = -w->yth / w->dyth;
      }
      for (int j = 0; j < K; j++) {
        if (w->dlam[j] < 0.0 && -w->lam[j] / w->dlam[j] < amax) amax = -w->lam[j] / w->dlam[j];
        if (w->dylam[j] < 0.0 && -w->ylam[j] / w->dylam[j] < amax) amax = -w->ylam[j] / w->dylam[j];
      }
      if (pass == 0) {
        const double aa = amax < 1.0 ? amax : 1.0;
        double mua = 0.0;
        for (int j = 0; j < N * MAXROW; j++) if (w->act[j]) { mua += (w->s[j] + aa * w->ds[j]) * (w->y[j] + aa * w->dy[j]); w->corr[j] = w->ds[j] * w->dy[j]; }
        if (soft && !ORC_FREE_THETA) { mua += (w->th + aa * w->dth) * (w->yth + aa * w->dyth); w->corr_th = w->dth * w->dyth; }
        for (int j = 0; j < K; j++) { mua += (w->lam[j] + aa * w->dlam[j]) * (w->ylam[j] + aa * w->dylam[j]); w->corr_lam[j] = w->dlam[j] * w->dylam[j]; }
        mua /= (double)m_total;
        sigma = pow(mua / mu, 3.0);
        { /* safeguarded second-order term (Mehrotra's corrector is harmful when the affine step is short) */
          const int mode = getenv("ORC_CORR") ? atoi(getenv("ORC_CORR")) : 5;
          const double sc = mode == 1 ? aa : (mode == 2 ? aa * aa : (mode == 3 ? (aa < 0.5 ? aa : 1.0) : (mode == 4 ? fmin(1.0, aa / 0.7) : (mode == 5 ? (aa < 0.2 ? aa : 1.0) : 1.0))));
          if (sc != 1.0) {
            for (int j = 0; j < N * MAXROW; j++) if (w->act[j]) w->corr[j] *= sc;
            w->corr_th *= sc;
            for (int j = 0; j < K; j++) w->corr_lam[j] *= sc;
          }
        }
      } else {
        { const double eta = getenv("ORC_ETA") ? atof(getenv("ORC_ETA")) : 1.0; double tau = 1.0 - fmin(0.005, eta * mu); alpha = tau * amax; if (alpha > 1.0) alpha = 1.0; }
      }
    }
    if (polishing) {
      /* Full Newton step of the augmented-Lagrangian model on the primal, multiplier update y += rho r on the
       * active rows, then one active-set refinement: active rows whose multiplier turned negative are released,
       * violated inactive rows are activated, and the step is repeated (at most PMAX rounds). */
      const double ftol = 1e-10, dtol = 1e-9;
      int changed = 0, viol = 0;
      double dymax = 0.0;   /* largest relative multiplier update: a second AL step removes the remaining bias */
      for (int i = 0; i < N - 1; i++) for (int k = 0; k < 2; k++) w->u[2 * i + k] += w->du[2 * i + k];
      for (int i = 1; i < N; i++) for (int k = 0; k < 6; k++) w->x[6 * i + k] += w->dx[6 * i + k];
      if (soft) {
        w->th += w->dth;
        if (w->pact_th) { dymax = fmax(dymax, fabs(prho * w->th) / (1.0 + fabs(w->yth))); w->yth += prho * (-w->th); if (w->yth < -dtol) { w->pact_th = 0; w->yth = 0.0; changed++; } }
        else if (!ORC_FREE_THETA && w->th < -ftol) { w->pact_th = 1; changed++; }
      }
      for (int j = 0; j < K; j++) {
        w->lam[j] += w->dlam[j];
        if (w->pnb[j]) { dymax = fmax(dymax, fabs(prho * w->lam[j]) / (1.0 + fabs(w->ylam[j]))); w->ylam[j] += prho * (-w->lam[j]); if (w->ylam[j] < -dtol) { w->pnb[j] = 0; w->ylam[j] = 0.0; changed++; } }
        else if (w->lam[j] < -ftol) { w->pnb[j] = 1; changed++; }
      }
      for (int i = 0; i < N; i++)
        for (int sl = 0; sl < MAXROW; sl++) {
          const int j = RID(w, i, sl);
          if (!w->act[j]) continue;
          double gv;
          if (sl < 12) gv = w->xb_sg[sl] * w->x[6 * i + w->xb_c[sl]];
          else if (sl < 16) { const int k = (sl - 12) >> 1; gv = ((sl & 1) ? -1.0 : 1.0) * w->u[2 * i + k]; }
          else if (sl < 20) { const int k = (sl - 16) >> 1; const double up = i ? w->u[2 * (i - 1) + k] : p->u_ic[k]; gv = ((sl & 1) ? -1.0 : 1.0) * (w->u[2 * i + k] - up) / p->T[i]; }
          else gv = ((sl & 1) ? -1.0 : 1.0) * w->x[6 * i + 1] - (soft ? w->th : 0.0);
          const double r = gv - w->rh[j];
          if (w->pact[j]) { dymax = fmax(dymax, fabs(prho * r) / (1.0 + fabs(w->y[j]))); w->y[j] += prho * r; if (w->y[j] < -dtol * fmax(1.0, fabs(w->y[j]))) { w->pact[j] = 0; w->y[j] = 0.0; changed++; } }
          else if (r > ftol) { w->pact[j] = 1; changed++; }
          if (!w->pact[j] && r > 1e-9) viol++;
        }
      /* same rule as the kernel (lmpc_qp_core.cuh): a polish whose active set is coming apart is abandoned at once */
      const int diverging = polishing >= 2 && changed > 8 && changed > 2 * prev_changed;
      prev_changed = changed;
      if (!diverging && (changed || dymax > (polishing >= 2 ? 1e-4 : 1e-6)) && polishing < PMAX) { polishing++; continue; }   /* LMPC_PDY / LMPC_PDY2 of the kernel */
      out->polished = (!diverging && changed == 0 && viol == 0) ? polishing : 0;
      if (out->polished) { status = ORC_OK; break; }
      goto polish_failed;
    }
    /* ---------- update ---------- */
    for (int i = 0; i < N - 1; i++) for (int k = 0; k < 2; k++) w->u[2 * i + k] += alpha * w->du[2 * i + k];
    for (int i = 1; i < N; i++) for (int k = 0; k < 6; k++) w->x[6 * i + k] += alpha * w->dx[6 * i + k];
    for (int j = 0; j < N * MAXROW; j++) if (w->act[j]) { w->s[j] += alpha * w->ds[j]; w->y[j] += alpha * w->dy[j]; }
    if (soft) { w->th += alpha * w->dth; w->yth += alpha * w->dyth; }
    for (int j = 0; j < K; j++) { w->lam[j] += alpha * w->dlam[j]; w->ylam[j] += alpha * w->dylam[j]; }
    rho_d *= (1.0 - alpha);
    { /* step-based acceptance: the primal step, per channel relative to max(1, |channel|), bounds the
       * remaining error once the iteration is in its fast final phase */
      double sx[6] = {0}, su[2] = {0}, sd[2] = {0};
      for (int i = 0; i < N; i++) for (int k = 0; k < 6; k++) sx[k] = fmax(sx[k], fabs(w->dx[6 * i + k]));
      for (int i = 0; i < N - 1; i++) for (int k = 0; k < 2; k++) {
        const double dup = i ? w->du[2 * (i - 1) + k] : 0.0;
        su[k] = fmax(su[k], fabs(w->du[2 * i + k])); sd[k] = fmax(sd[k], fabs(w->du[2 * i + k] - dup) / p->T[i]);
      }
      double stepn = 0.0;
      for (int k = 0; k < 6; k++) stepn = fmax(stepn, sx[k] / chx[k]);
      for (int k = 0; k < 2; k++) stepn = fmax(stepn, fmax(su[k] / chu[k], sd[k] / chd[k]));
      stepn *= alpha;
      if (getenv("ORC_PORT_DEBUG")) fprintf(stderr, "   stepn %.3e\n", stepn);
      /* geometric-tail estimate of the remaining error from two consecutive steps */
      const double ratio = (prev_stepn > 0.0) ? stepn / prev_stepn : 1.0;
      const double est = (ratio < 0.9) ? stepn * ratio / (1.0 - ratio) : 1e300;
      prev_stepn = stepn;
      if (!getenv("ORC_NOSTEP") && stepn < step_tol2 && est < step_tol2 && alpha > 0.5 && mu < 1e-6 && rpn < 1e-9 && fabs(rnu) < 1e-9) {
        if (!do_polish) { status = ORC_OK; it++; break; }
        polishing = 1;
      }
    }
    if (getenv("ORC_PORT_DEBUG")) fprintf(stderr, "it %2d mu %.3e rp %.3e sigma %.3e alpha %.4f th %.3e rnu %.2e\n", it, mu, rpn, sigma, alpha, w->th, rnu);
    if (0) { next_trip: ; }
  }
done:
  out->iters = it;
  /* final consistent rollout of the linear dynamics, outputs */
  for (int i = 0; i < N - 1; i++)
    for (int r = 0; r < 6; r++) {
      double a = p->g[6 * i + r];
      for (int k = 0; k < 6; k++) a += p->A[36 * i + r + 6 * k] * w->x[6 * i + k];
      for (int k = 0; k < 2; k++) a += p->B[12 * i + r + 6 * k] * w->u[2 * i + k];
      w->x[6 * (i + 1) + r] = a;
    }
  memcpy(out->X, w->x, sizeof(double) * 6 * (size_t)N);
  memcpy(out->U, w->u, sizeof(double) * 2 * (size_t)(N - 1));
  for (int i = 0; i < N - 1; i++) for (int k = 0; k < 2; k++) out->dU[2 * i + k] = (w->u[2 * i + k] - (i ? w->u[2 * (i - 1) + k] : p->u_ic[k])) / p->T[i];
  out->sigma_b = soft ? w->th : 0.0;
  memset(out->sigma_h, 0, sizeof out->sigma_h);
  if (learn) {
    if (out->lambda) for (int j = 0; j < K; j++) out->lambda[j] = w->lam[j] > 0.0 ? w->lam[j] : 0.0;
    for (int k = 0; k < 6; k++) { double a = w->x[6 * (N - 1) + k]; for (int j = 0; j < K; j++) a -= p->ssx[6 * j + k] * w->lam[j]; out->sigma_h[k] = a; }
    if (out->ss_x) memcpy(out->ss_x, p->ssx, sizeof(double) * 6 * (size_t)K);
    if (out->ss_cost) memcpy(out->ss_cost, p->ssc, sizeof(double) * (size_t)K);
  }
  out->cost = orc_eval_cost(c, p, out->X, out->U, out->dU, out->sigma_b, w->lam, out->sigma_h);
  out->status = status;
  free(w); free(p);
  return status;
}

/* ---------------------------------------------------------------------------------- */
typedef struct {
  const orc_vehicle* v; const orc_config* c; const orc_safe_set* ss; int B, impl, tid, nth;
  const double *x_ic, *u_ic, *X_ref, *U_ref, *T_ref, *bl, *br, *kap, *vref, *L;
  double *X, *U, *dU, *lambda, *cost, *kkt; int *status, *iters; int nfail;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  const int N = j->c->N, K = j->c->num_ss_pts > 0 ? j->c->num_ss_pts : 1;
  for (int b = j->tid; b < j->B; b += j->nth) {
    orc_step_in in = {j->x_ic + 6 * (size_t)b, j->u_ic + 2 * (size_t)b, j->X_ref + 6 * (size_t)N * b,
                      j->U_ref + 2 * (size_t)(N - 1) * b, j->T_ref + (size_t)(N - 1) * b, j->bl + (size_t)N * b,
                      j->br + (size_t)N * b, j->kap + (size_t)N * b, j->vref + (size_t)N * b, j->L[b], NULL};
    orc_step_out out; memset(&out, 0, sizeof out);
    out.X = j->X + 6 * (size_t)N * b; out.U = j->U + 2 * (size_t)(N - 1) * b; out.dU = j->dU + 2 * (size_t)(N - 1) * b;
    out.lambda = j->lambda ? j->lambda + (size_t)K * b : NULL;
    const int st = j->impl ? orc_step_dense(j->v, j->c, j->ss, &in, &out) : orc_step_port(j->v, j->c, j->ss, &in, &out);
    if (j->cost) j->cost[b] = out.cost;
    if (j->kkt) j->kkt[b] = j->impl ? out.kkt : (double)out.polished;   /* port: 1 = polish accepted */
    if (j->status) j->status[b] = st;
    if (j->iters) j->iters[b] = out.iters;
    if (st != ORC_OK) j->nfail++;
  }
  return NULL;
}

int orc_step_batch(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss, int B,
                   const double* x_ic, const double* u_ic, const double* X_ref, const double* U_ref,
                   const double* T_ref, const double* bl, const double* br, const double* kap,
                   const double* vref, const double* total_length, double* X, double* U, double* dU,
                   double* lambda, double* cost, int* status, int* iters, double* kkt, int impl, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  batch_job jobs[256]; pthread_t th[256];
  for (int t = 0; t < nthreads; t++) {
    batch_job j = {v, c, ss, B, impl, t, nthreads, x_ic, u_ic, X_ref, U_ref, T_ref, bl, br, kap, vref, total_length,
                   X, U, dU, lambda, cost, kkt, status, iters, 0};
    jobs[t] = j;
    if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  batch_worker(&jobs[0]);
  int nfail = jobs[0].nfail;
  for (int t = 1; t < nthreads; t++) { pthread_join(th[t], NULL); nfail += jobs[t].nfail; }
  return nfail;
}

/* ---------------------------------------------------------------------------------- SQP to convergence */
/* Damped fixed-point iteration on "linearise at (Xk, Uk) -> solve the QP".  The relaxation factor alpha follows the
 * angle between successive QP displacements d_k = QP(Xk, Uk) - (Xk, Uk): cos < -0.25 (oscillating: the Gauss-Newton
 * model has no constraint curvature and the LMPC terminal cost is linear) halves alpha (>= 1/8), cos > 0.25 doubles it
 * (<= 1).  Converged when the undamped displacement max |d| / max(1, |QP|) < tol; the QP solution is returned. */
int orc_step_sqp(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss, const orc_step_in* in,
                 orc_step_out* out, int max_sqp_iter, double tol, int impl, int* sqp_iters, double* defect) {
  const int N = c->N, nst = N - 1, nd = 6 * N + 2 * nst;
  double* Xk = (double*)malloc(sizeof(double) * 6 * (size_t)N);
  double* Uk = (double*)malloc(sizeof(double) * 2 * (size_t)nst);
  double* dprev = (double*)calloc((size_t)nd, sizeof(double));
  for (int i = 0; i < N; i++) {
    Xk[6 * i] = orc_align_abscissa(in->X_ref[6 * i], in->x_ic[0], in->total_length);
    for (int k = 1; k < 6; k++) Xk[6 * i + k] = in->X_ref[6 * i + k];
  }
  memcpy(Uk, in->U_ref, sizeof(double) * 2 * (size_t)nst);
  const double qp[2] = {Xk[6 * (N - 1)], Xk[6 * (N - 1) + 1]};
  int st = ORC_MAX_ITER, its = 0, converged = 0;
  double alpha = 1.0;
  for (int k = 0; k < max_sqp_iter; k++) {
    orc_step_in ik = *in;
    ik.X_ref = Xk; ik.U_ref = Uk; ik.ss_query_point = in->ss_query_point ? in->ss_query_point : qp;
    st = impl ? orc_step_dense(v, c, ss, &ik, out) : orc_step_port(v, c, ss, &ik, out);
    its++;
    if (st != ORC_OK && st != ORC_INACCURATE) break;
    double step = 0.0, dd = 0.0, dp = 0.0, pp = 0.0;
    for (int q = 0; q < nd; q++) {
      const double nv = q < 6 * N ? out->X[q] : out->U[q - 6 * N], ov = q < 6 * N ? Xk[q] : Uk[q - 6 * N];
      const double d = nv - ov;
      step = fmax(step, fabs(d) / fmax(1.0, fabs(nv)));
      dd += d * d; dp += d * dprev[q]; pp += dprev[q] * dprev[q];
      dprev[q] = d;
    }
    if (step < tol) { converged = 1; break; }
    if (k > 0) {
      /* successive displacements (anti)parallel: one mode d_k = (1 - alpha (1 + rho)) d_{k-1} dominates; the secant step
       * alpha / (1 - d_k.d_{k-1} / |d_{k-1}|^2) = 1 / (1 + rho) cancels it.  Otherwise halve on oscillation, double on progress. */
      const double cs = dp / sqrt(dd * pp + 1e-300);
      const double sr = dp / (pp + 1e-300);
      if (fabs(cs) > 0.9) alpha = (1.0 - sr > 0.1) ? fmin(fmax(alpha / (1.0 - sr), 0.125), 1.0) : 1.0;
      else if (cs < -0.25) alpha = fmax(0.5 * alpha, 0.125);
      else if (cs > 0.25) alpha = fmin(2.0 * alpha, 1.0);
    }
    for (int q = 0; q < 6 * N; q++) Xk[q] += alpha * dprev[q];
    for (int q = 0; q < 2 * nst; q++) Uk[q] += alpha * dprev[6 * N + q];
  }
  /* a run whose step test never passed is not a solution of the nonlinear problem (IPOPT: Maximum_Iterations_Exceeded) */
  if ((st == ORC_OK || st == ORC_INACCURATE) && !converged) st = ORC_SQP_MAX_ITER;
  out->status = st;
  if (sqp_iters) *sqp_iters = its;
  if (defect) {
    double dmax = 0.0;
    for (int i = 0; i < nst; i++) {
      double xn[6];
      orc_discrete_dynamics(v, out->X + 6 * i, out->U + 2 * i, in->curvatures[i], in->T_ref[i], xn);
      for (int k = 0; k < 6; k++) dmax = fmax(dmax, fabs(xn[k] - out->X[6 * (i + 1) + k]));
    }
    *defect = dmax;
  }
  free(Xk); free(Uk); free(dprev);
  return st;
}
