// lmpc_qp_core.cuh -- one warp solves one MPC instance: Mehrotra primal-dual interior point on the
// reference's QP (racing_mpc.cpp:31-202,442-543) with the Newton systems solved by a block
// Riccati recursion held in shared memory.
//
//   stage state  z_i = (x_i[6], u_{i-1}[2]),  stage control  u_i,  i = 0..N-2
//   z_{i+1} = [A_i 0; 0 0] z_i + [B_i; I] u_i + [g_i; 0]          (racing_mpc.cpp:182,195)
//   rate du_i = (u_i - u_{i-1}) / T_i enters cost and rows through (z_i, u_i)
//   global boundary slack sigma_b ("theta")    : second right-hand-side column + scalar solve
//   safe-set simplex (lambda, sigma_h)          : eliminated at the terminal stage; non-basic
//       columns through a 6x6 Woodbury system, the MB largest-Omega columns and the simplex
//       multiplier through a pivoted (1+MB)^2 LU (keeps the elimination accurate as mu -> 0)
//
// Inequality rows (all simple): state box, merged input box, input-rate box, soft track
// boundary, sigma_b >= 0, lambda >= 0.  Rows live in shared memory as [slot][stage].
//
// Written with the lane DSL of lmpc_warp.cuh: LANES_BEGIN/END phases, LaneVar registers, warp_*
// collectives; plain `double` variables outside phases are warp-uniform.
#pragma once
#include "lmpc_warp.cuh"
#include "../../include/lmpc_b200.h"

#define LMPC_MB 4               // explicit (basic-candidate) safe-set columns
#define LMPC_NQ (1 + LMPC_MB)   // + simplex multiplier
#define LMPC_KPL_MAX 4          // safe-set columns per lane (K <= 128)

struct LmpcQpParams {
  int N, NS, K, learning, soft, hull_slack;
  int nh, hidx[6];
  double Einv[6], chs[6];
  int nxb, xb_c[12];
  double xb_sg[12], xb_h[12];
  int RS;                        // row slots per stage = nxb + 10
  double ulo[2], uhi[2], dlo[2], dhi[2];
  int ub_act[4], db_act[4];      // finite flags in slot order (c0 hi, c0 lo, c1 hi, c1 lo)
  double margin, qb;
  double qx[6], qxN[6];          // tracking weights, stage / terminal (x10)
  double Rm[3], Rd[3];
  int max_iter;
  double tol;
  int NSd;                       // odd stage stride of the [.][stage] arrays
  // shared-memory offsets (doubles)
  int oABG, oS, oY, oCR, oX, oU, oDX, oDU, oHX, oCZX, oCZU, oCZTH, oCW, oEE, oUQ, oFAC, oKFF, oBL, oBR,
      oVREF, oT, oPM, oL1, oLTH, oMAB, oAXBW, oYY, oTERM, total;
};

// terminal-block scratch layout inside oTERM (doubles)
#define TB_PHI 0                       // 6x6
#define TB_PHIC (TB_PHI + 36)          // 6 x NQ
#define TB_S2 (TB_PHIC + 6 * LMPC_NQ)  // NQ x NQ (LU)
#define TB_XQ (TB_S2 + LMPC_NQ * LMPC_NQ)  // NQ x 6
#define TB_Q0 (TB_XQ + 6 * LMPC_NQ)    // NQ
#define TB_PT (TB_Q0 + LMPC_NQ)        // 6x6
#define TB_PTV (TB_PT + 36)            // 6
#define TB_BCOL (TB_PTV + 6)           // MB x 6  centred basic columns
#define TB_BD (TB_BCOL + 6 * LMPC_MB)  // MB      y/lambda of basic columns
#define TB_BG (TB_BD + LMPC_MB)        // MB      g_lambda of basic columns
#define TB_PHIR (TB_BG + LMPC_MB)      // 6       Phi r1
#define TB_PIV (TB_PHIR + 6)           // NQ (as doubles)
#define TB_SIZE (TB_PIV + LMPC_NQ)

struct LmpcQpIn {
  const double* x_ic;    // 6
  const double* u_ic;    // 2
  const double* U0;      // 2 x NS initial controls
  const double* T;       // NS
  const double* bl;      // N
  const double* br;      // N
  const double* vref;    // N
  const double* ABg;     // NS x 54  (A 36 col-major, B 12, g 6)
  const double* ssx;     // K x 6 safe-set columns (padded), learning only
  const double* ssc;     // K   J - J0
  const double* cen;     // 6   centre for the columns (the query point X_ref[:, N-1])
  int ss_count;          // 0 => no safe set
};

struct LmpcQpOut {
  double* X;       // N x 6
  double* U;       // NS x 2
  double* dU;      // NS x 2
  double* lam;     // K (may be null)
  double* cost;    // 1 (may be null)
  int* status;     // 1
  int* iters;      // 1
};

struct Arr4 { double a[LMPC_KPL_MAX]; };
struct Arr4x6 { double a[LMPC_KPL_MAX][6]; };
struct Arr4i { int a[LMPC_KPL_MAX]; };

#if defined(LMPC_EMULATE)
#define LANE0_ONLY(stmt) { stmt; }
#define WARP_SYNC()
#else
#define LANE0_ONLY(stmt) { if ((threadIdx.x & 31u) == 0u) { stmt; } }
#define WARP_SYNC() __syncwarp()
#endif

// row geometry helpers ---------------------------------------------------------------------
// slot classes: [0,nxb) x-box, [nxb,nxb+4) u-box, [nxb+4,nxb+8) du-box, nxb+8 / nxb+9 boundary L / R
LMPC_DEV bool row_active(const LmpcQpParams& P, int sl, int i) {
  if (sl < P.nxb) return i >= 1 && i <= P.N - 2;
  if (sl < P.nxb + 4) return i <= P.N - 2 && P.ub_act[sl - P.nxb];
  if (sl < P.nxb + 8) return i <= P.N - 2 && P.db_act[sl - P.nxb - 4];
  return P.soft || i >= 1;
}
// G v of a row from the vectors xs/us (either the iterate or the direction); th is sigma_b or its step;
// uprev0 is u_ic (iterate) or 0 (direction)
LMPC_DEV double row_gv(const LmpcQpParams& P, int sl, int i, const double* xs, const double* us, const double* Ts,
                       double th, const double* uprev0) {
  const int d = P.NSd;
  if (sl < P.nxb) return P.xb_sg[sl] * xs[P.xb_c[sl] * d + i];
  if (sl < P.nxb + 4) { const int q = sl - P.nxb; const double sg = (q & 1) ? -1.0 : 1.0; return sg * us[(q >> 1) * d + i]; }
  if (sl < P.nxb + 8) {
    const int q = sl - P.nxb - 4, c = q >> 1; const double sg = (q & 1) ? -1.0 : 1.0;
    const double up = i ? us[c * d + i - 1] : uprev0[c];
    return sg * (us[c * d + i] - up) / Ts[i];
  }
  const double sg = (sl == P.nxb + 8) ? 1.0 : -1.0;
  return sg * xs[1 * d + i] - (P.soft ? th : 0.0);
}
LMPC_DEV double row_h(const LmpcQpParams& P, int sl, int i, const double* bl, const double* br) {
  if (sl < P.nxb) return P.xb_h[sl];
  if (sl < P.nxb + 4) { const int q = sl - P.nxb; return (q & 1) ? -P.ulo[q >> 1] : P.uhi[q >> 1]; }
  if (sl < P.nxb + 8) { const int q = sl - P.nxb - 4; return (q & 1) ? -P.dlo[q >> 1] : P.dhi[q >> 1]; }
  return (sl == P.nxb + 8) ? (bl[i] - P.margin) : -(br[i] + P.margin);
}

// -------------------------------------------------------------------------------------------
// KPL = ceil(K/32) columns per lane (compile-time so that the per-column state stays in registers)
template <int KPL>
LMPC_DEV void lmpc_qp_solve_warp(const LmpcQpParams& P, const LmpcQpIn& in, double* sm, const LmpcQpOut& out) {
  const int N = P.N, NS = P.NS, d = P.NSd, RS = P.RS;
  const bool learn = P.learning != 0, soft = P.soft != 0;
  // Columns beyond the number actually found are copies of the last one (racing_mpc.cpp:263-272); they
  // are dropped here (their lambda stays 0): the optimum in X, U, dU and SS*lambda is the same, and
  // the explicit-column system stays non-singular.
  const int K = learn ? ((in.ss_count > 0 && in.ss_count < P.K) ? in.ss_count : P.K) : 0;
  double* ABG = sm + P.oABG;
  double* RSs = sm + P.oS; double* RSy = sm + P.oY; double* RScr = sm + P.oCR;
  double* X = sm + P.oX; double* U = sm + P.oU; double* DX = sm + P.oDX; double* DU = sm + P.oDU;
  double* HX = sm + P.oHX; double* CZX = sm + P.oCZX; double* CZU = sm + P.oCZU; double* CZTH = sm + P.oCZTH;
  double* CW = sm + P.oCW; double* EE = sm + P.oEE; double* UQ = sm + P.oUQ;
  double* FAC = sm + P.oFAC;   // per stage: Kz[16] (2x8 row-major), Sinv[3], pad
  double* KFF = sm + P.oKFF;   // per stage: kff1[2], kffth[2], Cwth[2]
  double* BL = sm + P.oBL; double* BR = sm + P.oBR; double* VREF = sm + P.oVREF; double* TT = sm + P.oT;
  double* PM = sm + P.oPM; double* L1 = sm + P.oL1; double* LTH = sm + P.oLTH;
  double* MAB = sm + P.oMAB; double* AXBW = sm + P.oAXBW; double* YY = sm + P.oYY;
  double* TB = sm + P.oTERM;
  const double sfloor = 1e-2, mu0 = 0.1, th0 = 0.01;

  // ---------------------------------------------------------------- load
  LANES_BEGIN
    for (int idx = lane; idx < 54 * NS; idx += 32) ABG[idx] = in.ABg[idx];
    for (int i = lane; i < N; i += 32) {
      BL[i] = in.bl[i]; BR[i] = in.br[i]; VREF[i] = in.vref[i];
      if (i < NS) {
        TT[i] = in.T[i];
        for (int c = 0; c < 2; c++) {
          double uu = in.U0[2 * i + c];
          uu = fmin(fmax(uu, P.ulo[c]), P.uhi[c]);
          U[c * d + i] = uu;
        }
      }
    }
    if (lane < 6) X[lane * d] = in.x_ic[lane];
  LANES_END
  double uic[2] = {in.u_ic[0], in.u_ic[1]};
  double zero2[2] = {0.0, 0.0};

  // status pre-checks (uniform)
  int status = LMPC_MAX_ITER;
  {
    bool bad = false;
    for (int sl = 0; sl < P.nxb; sl++) if (P.xb_sg[sl] * in.x_ic[P.xb_c[sl]] > P.xb_h[sl]) bad = true;
    if (!soft && (in.x_ic[1] > in.bl[0] - P.margin || in.x_ic[1] < in.br[0] + P.margin)) bad = true;
    if (bad) status = LMPC_INFEASIBLE_IC;
    else if (learn && in.ss_count <= 0) status = LMPC_NO_SAFE_SET;
  }
  int it = 0;

  // ---------------------------------------------------------------- linear rollout from x_ic
  for (int i = 0; i < NS; i++) {
    LANES_BEGIN
      if (lane < 6) {
        const double* A = ABG + 54 * i; const double* B = A + 36; const double* g = A + 48;
        double a = g[lane];
        for (int k = 0; k < 6; k++) a += A[lane + 6 * k] * X[k * d + i];
        for (int k = 0; k < 2; k++) a += B[lane + 6 * k] * U[k * d + i];
        X[lane * d + i + 1] = a;
      }
    LANES_END
  }

  // ---------------------------------------------------------------- initial slacks / multipliers
  double th = th0, yth = mu0 / th0, corr_th = 0.0, dth = 0.0, dyth = 0.0;
  LaneVar<double> red0, red1, red2;
  LaneVar<Arr4> lam, ylam, corl, dlam, glam, omg_;
  LaneVar<Arr4x6> St;
  LaneVar<Arr4> sscv;
  LaneVar<Arr4i> isB;
  LANES_BEGIN
    double r0 = 1.0;
    for (int i = lane; i < N; i += 32)
      for (int sl = 0; sl < RS; sl++) {
        if (!row_active(P, sl, i)) { RSs[sl * d + i] = 1.0; RSy[sl * d + i] = 0.0; RScr[sl * d + i] = 0.0; continue; }
        const double slack = row_h(P, sl, i, BL, BR) - row_gv(P, sl, i, X, U, TT, th, uic);
        const double s = slack > sfloor ? slack : sfloor;
        RSs[sl * d + i] = s; RSy[sl * d + i] = mu0 / s; RScr[sl * d + i] = 0.0;
        r0 = fmax(r0, mu0 / s);
      }
    for (int p = 0; p < KPL; p++) {
      const int k = lane + 32 * p;
      const bool on = learn && k < K;
      lam(lane).a[p] = on ? 1.0 / K : 0.0; ylam(lane).a[p] = on ? mu0 * K : 0.0;
      corl(lane).a[p] = 0.0; dlam(lane).a[p] = 0.0; glam(lane).a[p] = 0.0; omg_(lane).a[p] = 0.0; isB(lane).a[p] = 0;
      sscv(lane).a[p] = on ? in.ssc[k] : 0.0;
      for (int c = 0; c < 6; c++) St(lane).a[p][c] = on ? in.ssx[6 * k + c] - in.cen[c] : 0.0;
      if (on) r0 = fmax(r0, fabs(in.ssc[k]));
    }
    red0(lane) = r0;
  LANES_END
  warp_max(red0);
  double R0 = red0(0);
  if (soft) R0 = fmax(R0, 2.0 * P.qb * th);
  double rho_d = 1.0, prev_stepn = 0.0;
  // channel scales max(1, |channel|) of the parity metric (x, u, du), from the initial iterate
  double chs_[10];
  {
    LaneVar<double> rc_[10];
    LANES_BEGIN
      double m[10] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
      for (int i = lane; i < N; i += 32) {
        for (int c = 0; c < 6; c++) m[c] = fmax(m[c], fabs(X[c * d + i]));
        if (i < NS) for (int c = 0; c < 2; c++) {
          const double up = i ? U[c * d + i - 1] : uic[c];
          m[6 + c] = fmax(m[6 + c], fabs(U[c * d + i]));
          m[8 + c] = fmax(m[8 + c], fabs(U[c * d + i] - up) / TT[i]);
        }
      }
      for (int q = 0; q < 10; q++) rc_[q](lane) = m[q];
    LANES_END
    for (int q = 0; q < 10; q++) { warp_max(rc_[q]); chs_[q] = 1.0 / rc_[q](0); }
  }
  const double mu_floor = 1e-4 * P.tol;
  int m_total = 0;
  {
    for (int sl = 0; sl < RS; sl++) for (int i = 0; i < N; i++) m_total += row_active(P, sl, i) ? 1 : 0;
    if (soft) m_total += 1;
    if (learn) m_total += K;
  }
  const double inv_m = 1.0 / (double)m_total;

  // ================================================================ interior-point iterations
  for (; status == LMPC_MAX_ITER && it < P.max_iter; it++) {
    // ---------- residuals, mu, hull residual
    double sig[6] = {0, 0, 0, 0, 0, 0};
    LaneVar<double> rsig[6];
    LANES_BEGIN
      double msum = 0.0, rpm = 0.0;
      for (int i = lane; i < N; i += 32)
        for (int sl = 0; sl < RS; sl++) {
          if (!row_active(P, sl, i)) continue;
          const double s = RSs[sl * d + i], y = RSy[sl * d + i];
          const double rp = row_gv(P, sl, i, X, U, TT, th, uic) + s - row_h(P, sl, i, BL, BR);
          rpm = fmax(rpm, fabs(rp));
          msum += s * y;
        }
      double lsum = 0.0;
      double sg[6] = {0, 0, 0, 0, 0, 0};
      for (int p = 0; p < KPL; p++) {
        const double l = lam(lane).a[p];
        msum += l * ylam(lane).a[p]; lsum += l;
        for (int c = 0; c < 6; c++) sg[c] += St(lane).a[p][c] * l;
      }
      red0(lane) = msum; red1(lane) = rpm; red2(lane) = lsum;
      for (int c = 0; c < 6; c++) rsig[c](lane) = sg[c];
    LANES_END
    warp_sum(red0); warp_max(red1); warp_sum(red2);
    double mu = red0(0) + (soft ? th * yth : 0.0);
    mu *= inv_m;
    const double rpn = red1(0);
    const double rnu = learn ? red2(0) - 1.0 : 0.0;
    if (learn) {
      for (int c = 0; c < 6; c++) warp_sum(rsig[c]);
      for (int a = 0; a < P.nh; a++) { const int c = P.hidx[a]; sig[a] = X[c * d + N - 1] - in.cen[c] - rsig[c](0); }
    }
    if (mu < mu_floor && rpn < mu_floor && rho_d * R0 < mu_floor && fabs(rnu) < mu_floor) { status = LMPC_SOLVED; break; }

    double sigma = 0.0, alpha = 1.0, Pithth_keep = 0.0, csc = 1.0;   // csc: safeguard scale of the second-order term
    bool fail = false;
    for (int pass = 0; pass < 2 && !fail; pass++) {
      const double smu = sigma * mu;
      // ---------- assemble per-stage data (lane = stage)
      LANES_BEGIN
        double dth_acc = 0.0, cth_acc = 0.0;
        for (int i = lane; i < N; i += 32) {
          for (int c = 0; c < 6; c++) {
            double h = 0.0, g = 0.0;
            if (!learn) { const double w = (i == N - 1) ? P.qxN[c] : P.qx[c]; h = 2.0 * w; g = 2.0 * w * (X[c * d + i] - (c == 3 ? VREF[i] : 0.0)); }
            HX[c * d + i] = h; CZX[c * d + i] = g;
          }
          double cz_th = 0.0;
          double dd[2] = {0, 0}, td[2] = {0, 0}, dub[2] = {0, 0}, tub[2] = {0, 0};
          for (int sl = 0; sl < RS; sl++) {
            if (!row_active(P, sl, i)) continue;
            const double s = RSs[sl * d + i], y = RSy[sl * d + i];
            const double is = 1.0 / s;
            const double dj = y * is;
            const double rp = row_gv(P, sl, i, X, U, TT, th, uic) + s - row_h(P, sl, i, BL, BR);
            const double t = (smu - (pass ? csc * RScr[sl * d + i] : 0.0)) * is + dj * rp;
            if (sl < P.nxb) { const int c = P.xb_c[sl]; HX[c * d + i] += dj; CZX[c * d + i] += P.xb_sg[sl] * t; }
            else if (sl < P.nxb + 4) { const int q = sl - P.nxb; dub[q >> 1] += dj; tub[q >> 1] += ((q & 1) ? -t : t); }
            else if (sl < P.nxb + 8) { const int q = sl - P.nxb - 4; dd[q >> 1] += dj; td[q >> 1] += ((q & 1) ? -t : t); }
            else {
              const double sg = (sl == P.nxb + 8) ? 1.0 : -1.0;
              HX[1 * d + i] += dj; CZX[1 * d + i] += sg * t;
              if (soft) { cz_th += -sg * dj; dth_acc += dj; cth_acc += -t; }
            }
          }
          CZTH[i] = cz_th;
          if (i <= N - 2) {
            const double iT = 1.0 / TT[i];
            const double u0 = U[i], u1 = U[d + i];
            const double dc0 = (u0 - (i ? U[i - 1] : uic[0])) * iT, dc1 = (u1 - (i ? U[d + i - 1] : uic[1])) * iT;
            const double e0 = (2.0 * P.Rd[0] + dd[0]) * iT * iT, e1 = 2.0 * P.Rd[1] * iT * iT, e2 = (2.0 * P.Rd[2] + dd[1]) * iT * iT;
            const double ev0 = (2.0 * (P.Rd[0] * dc0 + P.Rd[1] * dc1) + td[0]) * iT;
            const double ev1 = (2.0 * (P.Rd[1] * dc0 + P.Rd[2] * dc1) + td[1]) * iT;
            EE[i] = e0; EE[d + i] = e1; EE[2 * d + i] = e2;
            UQ[i] = 2.0 * P.Rm[0] + dub[0]; UQ[d + i] = 2.0 * P.Rm[1]; UQ[2 * d + i] = 2.0 * P.Rm[2] + dub[1];
            CW[i] = 2.0 * (P.Rm[0] * u0 + P.Rm[1] * u1) + tub[0] + ev0;
            CW[d + i] = 2.0 * (P.Rm[1] * u0 + P.Rm[2] * u1) + tub[1] + ev1;
            CZU[i] = -ev0; CZU[d + i] = -ev1;
          } else { CZU[i] = 0.0; CZU[d + i] = 0.0; }
        }
        red0(lane) = dth_acc; red1(lane) = cth_acc;
      LANES_END
      warp_sum(red0); warp_sum(red1);
      double Dthth = red0(0), cth = red1(0);
      if (soft) { Dthth += 2.0 * P.qb + yth / th; cth += 2.0 * P.qb * th - (smu - (pass ? csc * corr_th : 0.0)) / th; }

      // ---------- terminal value  P_{N-1}, l_{N-1}
      LANES_BEGIN
        for (int o = lane; o < 64; o += 32) { const int r = o >> 3, c = o & 7; PM[o] = (r == c && r < 6) ? HX[r * d + N - 1] : 0.0; }
        if (lane < 8) { L1[lane] = lane < 6 ? CZX[lane * d + N - 1] : 0.0; LTH[lane] = (lane == 1) ? CZTH[N - 1] : 0.0; }
      LANES_END
      if (learn) {
        const int nh = P.nh;
        // ---- per-column weights; pass 0: pick the MB largest Omega as explicit columns
        if (pass == 0) {
          LANES_BEGIN
            for (int p = 0; p < KPL; p++) { const int k = lane + 32 * p; omg_(lane).a[p] = (k < K) ? lam(lane).a[p] / ylam(lane).a[p] : -1.0; isB(lane).a[p] = 0; }
          LANES_END
          for (int q = 0; q < LMPC_MB; q++) {
            LaneVar<double> bv; LaneVar<int> bi;
            LANES_BEGIN
              double v = -2.0; int ix = 1 << 30;
              for (int p = 0; p < KPL; p++) { const int k = lane + 32 * p; if (k < K && !isB(lane).a[p] && omg_(lane).a[p] > v) { v = omg_(lane).a[p]; ix = k; } }
              bv(lane) = v; bi(lane) = ix;
            LANES_END
            warp_argmax(bv, bi);
            const int kb = bi(0);   // uniform; K >= MB is required by the host
            LANES_BEGIN
              for (int p = 0; p < KPL; p++) if (lane + 32 * p == kb) {
                isB(lane).a[p] = 1 + q;
                for (int c = 0; c < 6; c++) TB[TB_BCOL + 6 * q + c] = St(lane).a[p][c];
                TB[TB_BD + q] = ylam(lane).a[p] / lam(lane).a[p];
              }
            LANES_END
          }
        }
        // ---- sums over the non-basic columns
        LaneVar<double> rW[21], ra[6], rb[6], rom, rog;
        LANES_BEGIN
          double W[21], a[6], b[6], om1 = 0.0, og = 0.0;
          for (int q = 0; q < 21; q++) W[q] = 0.0;
          for (int c = 0; c < 6; c++) { a[c] = 0.0; b[c] = 0.0; }
          for (int p = 0; p < KPL; p++) {
            const int k = lane + 32 * p;
            if (k >= K) continue;
            const double gl = sscv(lane).a[p] - (smu - (pass ? csc * corl(lane).a[p] : 0.0)) / lam(lane).a[p];
            glam(lane).a[p] = gl;
            if (isB(lane).a[p]) { TB[TB_BG + isB(lane).a[p] - 1] = gl; continue; }
            const double om = omg_(lane).a[p];
            og += om * gl; om1 += om;
            int q = 0;
            for (int aa = 0; aa < 6; aa++) {
              const double sa = (aa < nh) ? St(lane).a[p][P.hidx[aa]] : 0.0;
              b[aa] += sa * om * gl; a[aa] += sa * om;
              for (int bb = 0; bb <= aa; bb++, q++) W[q] += om * sa * ((bb < nh) ? St(lane).a[p][P.hidx[bb]] : 0.0);
            }
          }
          for (int q = 0; q < 21; q++) rW[q](lane) = W[q];
          for (int c = 0; c < 6; c++) { ra[c](lane) = a[c]; rb[c](lane) = b[c]; }
          rom(lane) = om1; rog(lane) = og;
        LANES_END
        if (pass == 0) { for (int q = 0; q < 21; q++) warp_sum(rW[q]); for (int c = 0; c < 6; c++) warp_sum(ra[c]); warp_sum(rom); }
        for (int c = 0; c < 6; c++) warp_sum(rb[c]);
        warp_sum(rog);
        // uniform: r1 = sigma + b_N ; (pass 0) Cholesky of Einv + W_N
        double r1[6];
        for (int a = 0; a < 6; a++) r1[a] = (a < nh) ? sig[a] + rb[a](0) : 0.0;
        if (pass == 0) {
          double Lc[21];   // lower triangle, row-major packed: (a,b) -> a(a+1)/2 + b
          for (int a = 0, q = 0; a < 6; a++) for (int b = 0; b <= a; b++, q++) Lc[q] = (a < nh && b < nh) ? rW[q](0) + (a == b ? P.Einv[a] : 0.0) : (a == b ? 1.0 : 0.0);
          bool ok = true;
          for (int j = 0; j < 6; j++) {
            double dg = Lc[j * (j + 1) / 2 + j];
            for (int k = 0; k < j; k++) dg -= Lc[j * (j + 1) / 2 + k] * Lc[j * (j + 1) / 2 + k];
            if (!(dg > 0.0)) { ok = false; dg = 1.0; }
            const double ld = sqrt(dg), il = 1.0 / ld;
            Lc[j * (j + 1) / 2 + j] = il;   // store the reciprocal of the diagonal
            for (int i2 = j + 1; i2 < 6; i2++) {
              double a2 = Lc[i2 * (i2 + 1) / 2 + j];
              for (int k = 0; k < j; k++) a2 -= Lc[i2 * (i2 + 1) / 2 + k] * Lc[j * (j + 1) / 2 + k];
              Lc[i2 * (i2 + 1) / 2 + j] = a2 * il;
            }
          }
          if (!ok) { fail = true; }
          // lanes solve in parallel: 0..5 identity columns (Phi), 6..6+NQ-1 columns of C, 11 unused
          LANES_BEGIN
            if (lane < 6 + LMPC_NQ) {
              double v[6];
              for (int a = 0; a < 6; a++) {
                if (lane < 6) v[a] = (a == lane) ? 1.0 : 0.0;
                else if (lane == 6) v[a] = (a < nh) ? -ra[a](lane) : 0.0;
                else v[a] = (a < nh) ? TB[TB_BCOL + 6 * (lane - 7) + P.hidx[a]] : 0.0;
              }
              for (int i2 = 0; i2 < 6; i2++) { double a2 = v[i2]; for (int k = 0; k < i2; k++) a2 -= Lc[i2 * (i2 + 1) / 2 + k] * v[k]; v[i2] = a2 * Lc[i2 * (i2 + 1) / 2 + i2]; }
              for (int i2 = 5; i2 >= 0; i2--) { double a2 = v[i2]; for (int k = i2 + 1; k < 6; k++) a2 -= Lc[k * (k + 1) / 2 + i2] * v[k]; v[i2] = a2 * Lc[i2 * (i2 + 1) / 2 + i2]; }
              if (lane < 6) { for (int a = 0; a < 6; a++) TB[TB_PHI + 6 * a + lane] = v[a]; }
              else { for (int a = 0; a < 6; a++) TB[TB_PHIC + a * LMPC_NQ + (lane - 6)] = v[a]; }
            }
          LANES_END
          // S2 = Z - C' Phi C   (NQ x NQ)
          LANES_BEGIN
            if (lane < LMPC_NQ * LMPC_NQ) {
              const int q = lane / LMPC_NQ, r = lane % LMPC_NQ;
              double z = 0.0;
              if (q == 0 && r == 0) z = rom(lane); else if (q == 0 || r == 0) z = -1.0; else if (q == r) z = -TB[TB_BD + q - 1];
              double s2 = 0.0;
              for (int a = 0; a < nh; a++) {
                const double cq = (q == 0) ? -ra[a](lane) : TB[TB_BCOL + 6 * (q - 1) + P.hidx[a]];
                s2 += cq * TB[TB_PHIC + a * LMPC_NQ + r];
              }
              TB[TB_S2 + q * LMPC_NQ + r] = z - s2;
            }
          LANES_END
          // pivoted LU of S2 (uniform, every lane redundantly; lane 0 stores)
          {
            double M[LMPC_NQ][LMPC_NQ]; int piv[LMPC_NQ];
            for (int q = 0; q < LMPC_NQ; q++) for (int r = 0; r < LMPC_NQ; r++) M[q][r] = TB[TB_S2 + q * LMPC_NQ + r];
            WARP_SYNC();
            for (int k = 0; k < LMPC_NQ; k++) {
              int pk = k; double mx = fabs(M[k][k]);
              for (int i2 = k + 1; i2 < LMPC_NQ; i2++) if (fabs(M[i2][k]) > mx) { mx = fabs(M[i2][k]); pk = i2; }
              piv[k] = pk;
              if (!(mx > 0.0)) { fail = true; mx = 1.0; M[pk][k] = 1.0; }
              for (int i2 = k + 1; i2 < LMPC_NQ; i2++) if (pk == i2) for (int c2 = 0; c2 < LMPC_NQ; c2++) { const double t = M[k][c2]; M[k][c2] = M[i2][c2]; M[i2][c2] = t; }
              const double ip = 1.0 / M[k][k];
              for (int i2 = k + 1; i2 < LMPC_NQ; i2++) { const double f = M[i2][k] * ip; M[i2][k] = f; for (int c2 = k + 1; c2 < LMPC_NQ; c2++) M[i2][c2] -= f * M[k][c2]; }
            }
            LANE0_ONLY(for (int q = 0; q < LMPC_NQ; q++) { for (int r = 0; r < LMPC_NQ; r++) TB[TB_S2 + q * LMPC_NQ + r] = M[q][r]; TB[TB_PIV + q] = (double)piv[q]; })
            WARP_SYNC();
          }
        }
        // ---- solves with the LU: lanes 0..5 -> Xq columns (pass 0), lane 6 -> q0
        LANES_BEGIN
          const bool doX = (pass == 0) && lane < 6 && lane < nh;
          const bool doQ = lane == 6;
          if (doX || doQ) {
            double v[LMPC_NQ];
            for (int q = 0; q < LMPC_NQ; q++) {
              if (doX) v[q] = TB[TB_PHIC + lane * LMPC_NQ + q];
              else {
                double r2 = (q == 0) ? (rnu - rog(lane)) : TB[TB_BG + q - 1];
                for (int a = 0; a < nh; a++) r2 -= TB[TB_PHIC + a * LMPC_NQ + q] * r1[a];
                v[q] = r2;
              }
            }
            for (int k = 0; k < LMPC_NQ; k++) { const int pk = (int)TB[TB_PIV + k]; for (int i2 = k + 1; i2 < LMPC_NQ; i2++) if (pk == i2) { const double t = v[k]; v[k] = v[i2]; v[i2] = t; } }
            for (int i2 = 0; i2 < LMPC_NQ; i2++) { double a2 = v[i2]; for (int k = 0; k < i2; k++) a2 -= TB[TB_S2 + i2 * LMPC_NQ + k] * v[k]; v[i2] = a2; }
            for (int i2 = LMPC_NQ - 1; i2 >= 0; i2--) { double a2 = v[i2]; for (int k = i2 + 1; k < LMPC_NQ; k++) a2 -= TB[TB_S2 + i2 * LMPC_NQ + k] * v[k]; v[i2] = a2 / TB[TB_S2 + i2 * LMPC_NQ + i2]; }
            if (doX) { for (int q = 0; q < LMPC_NQ; q++) TB[TB_XQ + q * 6 + lane] = v[q]; }
            else { for (int q = 0; q < LMPC_NQ; q++) TB[TB_Q0 + q] = v[q]; }
          }
        LANES_END
        // ---- PT = Phi + PhiC Xq (pass 0),  pT = Phi r1 - PhiC q0 ; add into P_{N-1}, l_{N-1}
        LANES_BEGIN
          if (pass == 0) {
            for (int o = lane; o < 36; o += 32) {
              const int a = o / 6, b = o % 6;
              double s2 = 0.0;
              if (a < nh && b < nh) { s2 = TB[TB_PHI + 6 * a + b]; for (int q = 0; q < LMPC_NQ; q++) s2 += TB[TB_PHIC + a * LMPC_NQ + q] * TB[TB_XQ + q * 6 + b]; }
              TB[TB_PT + o] = s2;
            }
          }
          if (lane < 6) {
            double s2 = 0.0;
            if (lane < nh) {
              for (int b = 0; b < nh; b++) s2 += TB[TB_PHI + 6 * lane + b] * r1[b];
              for (int q = 0; q < LMPC_NQ; q++) s2 -= TB[TB_PHIC + lane * LMPC_NQ + q] * TB[TB_Q0 + q];
            }
            TB[TB_PTV + lane] = s2;
          }
        LANES_END
        LANES_BEGIN
          for (int o = lane; o < 36; o += 32) { const int a = o / 6, b = o % 6; if (a < nh && b < nh) PM[8 * P.hidx[a] + P.hidx[b]] += TB[TB_PT + o]; }
          if (lane < nh) L1[P.hidx[lane]] += TB[TB_PTV + lane];
        LANES_END
      }
      if (fail) break;

      // ---------- backward Riccati sweep
      double Pi1th = cth, Pithth = Dthth;
      for (int i = NS - 1; i >= 0; i--) {
        const double* A = ABG + 54 * i; const double* B = A + 36;
        double* fac = FAC + 20 * i; double* kf = KFF + 6 * i;
        if (pass == 0) {
          // phase a: [M_xx A | M_xx B + M_xu] (6x8) and A' l_x, B' l_x + l_u for both rhs columns
          LANES_BEGIN
            for (int o = lane; o < 64; o += 32) {
              if (o < 48) {
                const int r = o >> 3, c = o & 7;
                double a = (c < 6) ? 0.0 : PM[8 * r + c];
                const double* col = (c < 6) ? (A + 6 * c) : (B + 6 * (c - 6));
                for (int k = 0; k < 6; k++) a += PM[8 * r + k] * col[k];
                MAB[o] = a;
              } else {
                const int q = o - 48, e = q & 7; const double* l = (q < 8) ? L1 : LTH;
                double a = (e < 6) ? 0.0 : l[e];
                const double* col = (e < 6) ? (A + 6 * e) : (B + 6 * (e - 6));
                for (int k = 0; k < 6; k++) a += col[k] * l[k];
                AXBW[q] = a;
              }
            }
          LANES_END
          // phase b: Yxx = A' MA, Yxu = A' MB, Yuu = B' MB + M_ux B + M_uu
          LANES_BEGIN
            for (int o = lane; o < 64; o += 32) {
              const int r = o >> 3, c = o & 7;
              if (r < 6) {
                double a = 0.0;
                for (int k = 0; k < 6; k++) a += A[k + 6 * r] * MAB[8 * k + c];
                YY[o] = a;
              } else if (c >= 6) {
                double a = PM[8 * r + c];
                for (int k = 0; k < 6; k++) a += B[k + 6 * (r - 6)] * MAB[8 * k + c] + PM[8 * k + r] * B[k + 6 * (c - 6)];
                YY[o] = a;
              }
            }
          LANES_END
          // phase c (uniform part): S = Yuu + E + Uq, its inverse
          const double e0 = EE[i], e1 = EE[d + i], e2 = EE[2 * d + i];
          const double q0_ = YY[8 * 6 + 6] + UQ[i], q1_ = 0.5 * (YY[8 * 6 + 7] + YY[8 * 7 + 6]) + UQ[d + i], q2_ = YY[8 * 7 + 7] + UQ[2 * d + i];
          const double s0 = q0_ + e0, s1 = q1_ + e1, s2_ = q2_ + e2;
          const double det = s0 * s2_ - s1 * s1;
          if (!(s0 > 0.0) || !(det > 0.0)) { fail = true; break; }
          const double idet = 1.0 / det;
          const double i0 = s2_ * idet, i1 = -s1 * idet, i2_ = s0 * idet;
          // Cw for both columns, kff = Sinv Cw
          const double cw1_0 = CW[i] + AXBW[6], cw1_1 = CW[d + i] + AXBW[7];
          const double cwt_0 = AXBW[8 + 6], cwt_1 = AXBW[8 + 7];
          const double k1_0 = i0 * cw1_0 + i1 * cw1_1, k1_1 = i1 * cw1_0 + i2_ * cw1_1;
          const double kt_0 = i0 * cwt_0 + i1 * cwt_1, kt_1 = i1 * cwt_0 + i2_ * cwt_1;
          Pithth -= cwt_0 * kt_0 + cwt_1 * kt_1;
          Pi1th -= cwt_0 * k1_0 + cwt_1 * k1_1;
          LANES_BEGIN
            // Qzw rows: r<6 -> Yxu[r][:], r>=6 -> -E[r-6][:]
            for (int o = lane; o < 64; o += 32) {
              const int r = o >> 3, c = o & 7;
              const double qr0 = (r < 6) ? YY[8 * r + 6] : ((r == 6) ? -e0 : -e1);
              const double qr1 = (r < 6) ? YY[8 * r + 7] : ((r == 6) ? -e1 : -e2);
              const double qc0 = (c < 6) ? YY[8 * c + 6] : ((c == 6) ? -e0 : -e1);
              const double qc1 = (c < 6) ? YY[8 * c + 7] : ((c == 6) ? -e1 : -e2);
              double pv;
              if (r >= 6 && c >= 6) {
                // P_uu = Q Sinv E  (product form, no cancellation), symmetrised
                const double se00 = i0 * e0 + i1 * e1, se01 = i0 * e1 + i1 * e2, se10 = i1 * e0 + i2_ * e1, se11 = i1 * e1 + i2_ * e2;
                const double p00 = q0_ * se00 + q1_ * se10, p01 = q0_ * se01 + q1_ * se11;
                const double p10 = q1_ * se00 + q2_ * se10, p11 = q1_ * se01 + q2_ * se11;
                pv = (r == 6 && c == 6) ? p00 : ((r == 7 && c == 7) ? p11 : 0.5 * (p01 + p10));
              } else {
                double qzz = 0.0;
                if (r < 6 && c < 6) qzz = YY[8 * r + c] + (r == c ? HX[r * d + i] : 0.0);
                pv = qzz - (qr0 * (i0 * qc0 + i1 * qc1) + qr1 * (i1 * qc0 + i2_ * qc1));
              }
              PM[o] = pv;
            }
            if (lane < 16) {   // Kz (2x8): Kz[j][c] = Sinv[j][:] . Qzw[c][:]
              const int j = lane >> 3, c = lane & 7;
              const double qc0 = (c < 6) ? YY[8 * c + 6] : ((c == 6) ? -e0 : -e1);
              const double qc1 = (c < 6) ? YY[8 * c + 7] : ((c == 6) ? -e1 : -e2);
              fac[lane] = (j == 0) ? (i0 * qc0 + i1 * qc1) : (i1 * qc0 + i2_ * qc1);
            } else if (lane < 32) {   // l1 (lanes 16..23) and lth (24..31):  l = Cz' - Qzw kff
              const int q = lane - 16, r = q & 7; const bool isth = q >= 8;
              const double qr0 = (r < 6) ? YY[8 * r + 6] : ((r == 6) ? -e0 : -e1);
              const double qr1 = (r < 6) ? YY[8 * r + 7] : ((r == 6) ? -e1 : -e2);
              const double kk0 = isth ? kt_0 : k1_0, kk1 = isth ? kt_1 : k1_1;
              double cz;
              if (isth) cz = (r == 1) ? CZTH[i] : 0.0;
              else cz = (r < 6) ? CZX[r * d + i] : CZU[(r - 6) * d + i];
              const double ax = (r < 6) ? AXBW[(isth ? 8 : 0) + r] : 0.0;
              (isth ? LTH : L1)[r] = cz + ax - (qr0 * kk0 + qr1 * kk1);
            }
            if (lane == 0) { fac[16] = i0; fac[17] = i1; fac[18] = i2_; kf[0] = k1_0; kf[1] = k1_1; kf[2] = kt_0; kf[3] = kt_1; kf[4] = cwt_0; kf[5] = cwt_1; }
          LANES_END
        } else {
          // pass 1: right-hand side "1" column only (factors unchanged)
          const double i0 = fac[16], i1 = fac[17], i2_ = fac[18];
          double bw0 = L1[6], bw1 = L1[7];
          for (int k = 0; k < 6; k++) { bw0 += B[k] * L1[k]; bw1 += B[6 + k] * L1[k]; }
          const double cw0 = CW[i] + bw0, cw1 = CW[d + i] + bw1;
          const double k0 = i0 * cw0 + i1 * cw1, k1 = i1 * cw0 + i2_ * cw1;
          Pi1th -= kf[4] * k0 + kf[5] * k1;
          LaneVar<double> newl;
          LANES_BEGIN
            if (lane < 8) {
              const int r = lane;
              double a = (r < 6) ? CZX[r * d + i] : CZU[(r - 6) * d + i];
              if (r < 6) for (int k = 0; k < 6; k++) a += A[k + 6 * r] * L1[k];
              a -= fac[r] * cw0 + fac[8 + r] * cw1;
              newl(lane) = a;
            }
          LANES_END
          LANES_BEGIN
            if (lane < 8) L1[lane] = newl(lane);
            if (lane == 8) { kf[0] = k0; kf[1] = k1; }
          LANES_END
        }
      }
      if (fail) break;
      if (pass == 0) Pithth_keep = Pithth; else Pithth = Pithth_keep;

      // ---------- sigma_b step, forward sweep
      dth = soft ? -Pi1th / Pithth : 0.0;
      LANES_BEGIN
        if (lane < 6) DX[lane * d] = 0.0;
      LANES_END
      for (int i = 0; i < NS; i++) {
        const double* A = ABG + 54 * i; const double* B = A + 36;
        const double* fac = FAC + 20 * i; const double* kf = KFF + 6 * i;
        double dz[8];
        for (int k = 0; k < 6; k++) dz[k] = DX[k * d + i];
        dz[6] = i ? DU[i - 1] : 0.0; dz[7] = i ? DU[d + i - 1] : 0.0;
        double du0 = -kf[0] - kf[2] * dth, du1 = -kf[1] - kf[3] * dth;
        for (int k = 0; k < 8; k++) { du0 -= fac[k] * dz[k]; du1 -= fac[8 + k] * dz[k]; }
        LANES_BEGIN
          if (lane < 6) {
            double a = B[lane] * du0 + B[6 + lane] * du1;
            for (int k = 0; k < 6; k++) a += A[lane + 6 * k] * dz[k];
            DX[lane * d + i + 1] = a;
          } else if (lane == 6) { DU[i] = du0; DU[d + i] = du1; }
        LANES_END
      }

      // ---------- terminal directions (lambda), sigma_b dual
      if (learn) {
        const int nh = P.nh;
        double e[6], qv[LMPC_NQ];
        for (int a = 0; a < 6; a++) {
          double s2 = 0.0;
          if (a < nh) { s2 = TB[TB_PTV + a]; for (int b = 0; b < nh; b++) s2 += TB[TB_PT + 6 * a + b] * DX[P.hidx[b] * d + N - 1]; }
          e[a] = s2;
        }
        for (int q = 0; q < LMPC_NQ; q++) { double s2 = TB[TB_Q0 + q]; for (int a = 0; a < nh; a++) s2 -= TB[TB_XQ + q * 6 + a] * DX[P.hidx[a] * d + N - 1]; qv[q] = s2; }
        const double nu = qv[0];
        LANES_BEGIN
          for (int p = 0; p < KPL; p++) {
            const int k = lane + 32 * p;
            if (k >= K) continue;
            const double l = lam(lane).a[p], y = ylam(lane).a[p];
            const double tl = (smu - (pass ? csc * corl(lane).a[p] : 0.0)) / l;
            double dl;
            if (isB(lane).a[p]) dl = qv[isB(lane).a[p]];
            else {
              double se = 0.0;
              for (int a = 0; a < nh; a++) se += St(lane).a[p][P.hidx[a]] * e[a];
              dl = omg_(lane).a[p] * (se - glam(lane).a[p] - nu);
            }
            dlam(lane).a[p] = dl;
            // dy written into omg_? no: keep separate -- reuse glam for dy after its last use
            glam(lane).a[p] = tl - y - dl * (y / l);
          }
        LANES_END
      }
      if (soft) { const double tt = (smu - (pass ? csc * corr_th : 0.0)) / th; dyth = tt - yth - (yth / th) * dth; }

      // ---------- row directions, step length (lane = stage)
      LANES_BEGIN
        double amax = 1e300, cross = 0.0;
        for (int i = lane; i < N; i += 32)
          for (int sl = 0; sl < RS; sl++) {
            if (!row_active(P, sl, i)) continue;
            const double s = RSs[sl * d + i], y = RSy[sl * d + i];
            const double rp = row_gv(P, sl, i, X, U, TT, th, uic) + s - row_h(P, sl, i, BL, BR);
            const double dg = row_gv(P, sl, i, DX, DU, TT, dth, zero2);
            const double rc = s * y - smu + (pass ? csc * RScr[sl * d + i] : 0.0);
            const double ds = -rp - dg;
            const double dy = (-rc - y * ds) / s;
            if (ds < 0.0) amax = fmin(amax, -s / ds);
            if (dy < 0.0) amax = fmin(amax, -y / dy);
            if (pass == 0) { RScr[sl * d + i] = ds * dy; cross += ds * dy; }
          }
        for (int p = 0; p < KPL; p++) {
          const int k = lane + 32 * p;
          if (k >= K) continue;
          const double dl = dlam(lane).a[p], dy = glam(lane).a[p];
          if (dl < 0.0) amax = fmin(amax, -lam(lane).a[p] / dl);
          if (dy < 0.0) amax = fmin(amax, -ylam(lane).a[p] / dy);
          if (pass == 0) { corl(lane).a[p] = dl * dy; cross += dl * dy; }
        }
        red0(lane) = amax; red1(lane) = cross;
      LANES_END
      warp_min(red0); warp_sum(red1);
      double amax = red0(0), cross = red1(0);
      if (soft) {
        if (dth < 0.0) amax = fmin(amax, -th / dth);
        if (dyth < 0.0) amax = fmin(amax, -yth / dyth);
        if (pass == 0) { corr_th = dth * dyth; cross += corr_th; }
      }
      if (pass == 0) {
        // mu_aff: sum (s + a ds)(y + a dy) = (1 - a) sum s y + a^2 sum ds dy   (affine step: s dy + y ds = -s y)
        const double aa = amax < 1.0 ? amax : 1.0;
        const double mua = (1.0 - aa) * mu + aa * aa * cross * inv_m;
        const double rt = mua / mu;
        sigma = rt * rt * rt;
        // Mehrotra's second-order term is harmful when the affine step is short: damp it then
        csc = (aa < 0.2) ? aa : 1.0;
      } else {
        const double tau = 1.0 - fmin(0.005, mu);
        alpha = tau * amax; if (alpha > 1.0) alpha = 1.0;
      }
    }  // pass
    if (fail) {
      status = (mu < 1e-9 && rpn < 1e-9) ? LMPC_SOLVED : LMPC_NUMERIC;   // numerical floor of the recursion
      break;
    }
    // ---------- update the iterate
    {
      const double smu = sigma * mu;
      LANES_BEGIN
        for (int i = lane; i < N; i += 32)
          for (int sl = 0; sl < RS; sl++) {
            if (!row_active(P, sl, i)) continue;
            const double s = RSs[sl * d + i], y = RSy[sl * d + i];
            const double rp = row_gv(P, sl, i, X, U, TT, th, uic) + s - row_h(P, sl, i, BL, BR);
            const double dg = row_gv(P, sl, i, DX, DU, TT, dth, zero2);
            const double rc = s * y - smu + csc * RScr[sl * d + i];
            const double ds = -rp - dg;
            const double dy = (-rc - y * ds) / s;
            RSs[sl * d + i] = s + alpha * ds; RSy[sl * d + i] = y + alpha * dy;
          }
        for (int p = 0; p < KPL; p++) {
          const int k = lane + 32 * p;
          if (k >= K) continue;
          lam(lane).a[p] += alpha * dlam(lane).a[p]; ylam(lane).a[p] += alpha * glam(lane).a[p];
        }
      LANES_END
      // the rows read X/U of neighbouring stages, so the primal update is a separate phase
      LANES_BEGIN
        for (int i = lane; i < N; i += 32) {
          if (i >= 1) for (int c = 0; c < 6; c++) X[c * d + i] += alpha * DX[c * d + i];
          if (i < NS) for (int c = 0; c < 2; c++) U[c * d + i] += alpha * DU[c * d + i];
        }
      LANES_END
      if (soft) { th += alpha * dth; yth += alpha * dyth; }
      rho_d *= (1.0 - alpha);
      // step-based acceptance: the primal step per channel, relative to max(1, |channel|), with a
      // geometric-tail estimate of what is still to come
      LANES_BEGIN
        double m = 0.0;
        for (int i = lane; i < N; i += 32) {
          for (int c = 0; c < 6; c++) m = fmax(m, fabs(DX[c * d + i]) * chs_[c]);
          if (i < NS) for (int c = 0; c < 2; c++) {
            const double dup = i ? DU[c * d + i - 1] : 0.0;
            m = fmax(m, fabs(DU[c * d + i]) * chs_[6 + c]);
            m = fmax(m, fabs(DU[c * d + i] - dup) / TT[i] * chs_[8 + c]);
          }
        }
        red0(lane) = m;
      LANES_END
      warp_max(red0);
      const double stepn = alpha * red0(0);
      const double ratio = (prev_stepn > 0.0) ? stepn / prev_stepn : 1.0;
      const double est = (ratio < 0.9) ? stepn * ratio / (1.0 - ratio) : 1e300;
      prev_stepn = stepn;
      if (stepn < P.tol && est < P.tol && alpha > 0.5 && mu < 1e-6 && rpn < 1e-9 && fabs(rnu) < 1e-9) { status = LMPC_SOLVED; it++; break; }
    }
  }  // iterations

  // ---------------------------------------------------------------- outputs
  for (int i = 0; i < NS; i++) {   // consistent rollout of the linear dynamics
    LANES_BEGIN
      if (lane < 6) {
        const double* A = ABG + 54 * i; const double* B = A + 36; const double* g = A + 48;
        double a = g[lane];
        for (int k = 0; k < 6; k++) a += A[lane + 6 * k] * X[k * d + i];
        for (int k = 0; k < 2; k++) a += B[lane + 6 * k] * U[k * d + i];
        X[lane * d + i + 1] = a;
      }
    LANES_END
  }
  LaneVar<double> rhs_[6];
  LANES_BEGIN
    double cst = 0.0;
    for (int i = lane; i < N; i += 32) {
      for (int c = 0; c < 6; c++) out.X[6 * i + c] = X[c * d + i];
      if (i < NS) {
        const double u0 = U[i], u1 = U[d + i];
        const double iT = 1.0 / TT[i];
        const double d0 = (u0 - (i ? U[i - 1] : uic[0])) * iT, d1 = (u1 - (i ? U[d + i - 1] : uic[1])) * iT;
        out.U[2 * i] = u0; out.U[2 * i + 1] = u1; out.dU[2 * i] = d0; out.dU[2 * i + 1] = d1;
        cst += u0 * (P.Rm[0] * u0 + P.Rm[1] * u1) + u1 * (P.Rm[1] * u0 + P.Rm[2] * u1);
        cst += d0 * (P.Rd[0] * d0 + P.Rd[1] * d1) + d1 * (P.Rd[1] * d0 + P.Rd[2] * d1);
      }
      if (!learn) {
        for (int c = 1; c < 6; c++) { const double w = (i == N - 1) ? P.qxN[c] : P.qx[c]; const double v = X[c * d + i] - (c == 3 ? VREF[i] : 0.0); cst += w * v * v; }
      }
    }
    double sg[6] = {0, 0, 0, 0, 0, 0};
    for (int p = 0; p < KPL; p++) {
      const int k = lane + 32 * p;
      if (k >= K) continue;
      const double l = lam(lane).a[p];
      cst += sscv(lane).a[p] * l;
      for (int c = 0; c < 6; c++) sg[c] += St(lane).a[p][c] * l;
    }
    if (out.lam) for (int p = 0; p < KPL; p++) { const int k = lane + 32 * p; if (k < P.K) out.lam[k] = (k < K) ? lam(lane).a[p] : 0.0; }
    red0(lane) = cst;
    for (int c = 0; c < 6; c++) rhs_[c](lane) = sg[c];
  LANES_END
  warp_sum(red0);
  double cost = red0(0);
  if (soft) cost += P.qb * th * th;
  if (learn && P.hull_slack) {
    for (int c = 0; c < 6; c++) warp_sum(rhs_[c]);
    for (int c = 0; c < 6; c++) { const double sh = X[c * d + N - 1] - in.cen[c] - rhs_[c](0); cost += P.chs[c] * sh * sh; }
  }
  LANE0_ONLY(if (out.cost) *out.cost = cost; *out.status = status; *out.iters = it;)
}
