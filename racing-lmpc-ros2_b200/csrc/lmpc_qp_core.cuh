// lmpc_qp_core.cuh -- one warp group (NW warps, one CTA) solves one MPC instance: Mehrotra primal-dual
// interior point on the reference's QP (racing_mpc.cpp:31-202,442-543) with the Newton systems solved
// by a block Riccati recursion held in shared memory.
//
//   stage state  z_i = (x_i[6], u_{i-1}[2]),  stage control  u_i,  i = 0..N-2
//   z_{i+1} = [A_i 0; 0 0] z_i + [B_i; I] u_i + [g_i; 0]          (racing_mpc.cpp:182,195)
//   rate du_i = (u_i - u_{i-1}) / T_i enters cost and rows through (z_i, u_i)
//   global boundary slack sigma_b ("theta")    : second right-hand-side column + scalar solve
//   safe-set simplex (lambda, sigma_h)          : eliminated at the terminal stage; non-basic
//       columns through a 6x6 Woodbury system, the MB largest-Omega columns and the simplex
//       multiplier through a pivoted (1+MB)^2 LU (keeps the elimination accurate as mu -> 0)
//
// Inequality rows (all simple): state box, merged input box, input-rate box, soft track boundary,
// sigma_b >= 0, lambda >= 0.  Rows live in shared memory as [slot][stage] with (s, y, 1/(s y)).
// Row work is organised in 10 *groups* per stage -- GX(c) c=0..5 (box rows of state c, plus the two
// boundary rows when c = e_y), GU(c) (input box), GD(c) (rate box) -- one lane per (group, stage).
//
// Written with the lane DSL of lmpc_warp.cuh; plain variables outside phases are group-uniform.
#pragma once
#include "lmpc_warp.cuh"
#include "../../include/lmpc_b200.h"
#ifdef LMPC_DEBUG_TRACE
#include <stdio.h>
#endif

#define LMPC_MB 4               // explicit (basic-candidate) safe-set columns
#define LMPC_NQ (1 + LMPC_MB)   // + simplex multiplier
#define LMPC_KPL_MAX 4          // safe-set columns per lane
#define LMPC_NRED 36            // widest multi-value reduction
#define LMPC_MAX_PEERS 8        // GPUs of one NVSwitch domain that exchange results through peer memory
#ifndef LMPC_PMAX
#define LMPC_PMAX 6             // active-set refinement rounds of the polish
#endif
#ifndef LMPC_PDY
#define LMPC_PDY 1e-6           // relative multiplier change below which the polish stops after its first round (with the boundary
                                // slack free most instances do, and this is what bounds their error: change / rho)
#endif
#ifndef LMPC_PDY2
#define LMPC_PDY2 1e-4          // ... and after a later round (an ill-conditioned instance does not get below 1e-6 at all)
#endif
#ifndef LMPC_PRHO
#define LMPC_PRHO 1e7           // augmented-Lagrangian weight of the polish
#endif
// Qzw row r lives in YY: rows 0..5 are Yxu = YY[8r+6..7]; rows 6,7 (= -E) are parked in YY[8r+0..1]
#define QZ0(r) (((r) < 6) ? (8 * (r) + 6) : (8 * (r)))
// 8x8 block phases: LMPC_L8 rows of 8 outputs per round, LMPC_R8 rounds (NT = 32: 4 x 2; 64: 8 x 1; 128: 8 x 1, half idle)
#define LMPC_L8 ((NT) >= 64 ? 8 : (NT) / 8)
#define LMPC_R8 (8 / LMPC_L8)

// Shared-memory layout (offsets in doubles) as a function of (N, rows per stage, warps per instance).
// constexpr so that the kernels instantiated for the named horizons fold every address into an immediate.
struct LmpcLayout {
  int oABG, oS, oY, oISY, oX, oU, oDXA, oDUA, oDXF, oDUF, oCZX, oCZTH, oGUD, oFAC, oKFF, oBL, oBR, oVREF, oIT,
      oPM, oL1, oLTH, oMAB, oAXBW, oYY, oRED, oTERM, oROWS, oMBAR, oCHS, total;
};
#define LMPC_MAX_ROWS 22   // 6 x 2 state boxes + 2 boundary + 4 control boxes + 4 rate boxes
// Row table (depends on the configuration only; built on the host, lmpc_host_params.h).  Every row is
//   sg * var - [theta] <= bnd,  sg = +1 (upper side) / -1 (lower side),  var = x_c(i), u_c(i) or du_c(i);
// entry [2 c + side] of the x / u / d arrays, [side] of the boundary pair (left, right).  slot = -1: row absent.
struct LmpcRowTab {
  double xbnd[12], ubnd[4], dbnd[4];
  int xslot[12], uslot[4], dslot[4], bslot[2];
  int ib0, pad_;   // first stage of the boundary rows (0 when they are soft, 1 otherwise)
  // the same rows by slot, for the passes that give one lane to one row (LMPC_FOR_ROWS_BY_LANE):
  // sdesc = component of (x, u) [bits 0-3] | lower side [4] | boundary row [5] | rate row [6] | first stage [8] |
  //         last stage is N-1 instead of N-2 [9] | slot in use [10];  sbnd = the constant bound (boundary rows take theirs from the track)
  double sbnd[LMPC_MAX_ROWS];
  int sdesc[LMPC_MAX_ROWS];
};
static_assert(sizeof(LmpcRowTab) % 8 == 0, "row table is copied as doubles");
#define LMPC_ROWS_DOUBLES ((int)(sizeof(LmpcRowTab) / 8))
#define LMPC_TB_SIZE_ (36 + 6 * LMPC_NQ + LMPC_NQ * LMPC_NQ + 6 * LMPC_NQ + LMPC_NQ + 36 + 6 + 6 * LMPC_MB + LMPC_MB + LMPC_MB + LMPC_NQ + 1 + 6 + LMPC_NQ)
LMPC_HD constexpr int lmpc_even(int n) { return (n + 1) & ~1; }   // keep 16-byte alignment
LMPC_HD constexpr LmpcLayout lmpc_layout(int N, int RS, int NW) {
  LmpcLayout L{};
  const int NS = N - 1, d = N | 1;
  int o = 0;
  L.oABG = o; o += lmpc_even(54 * NS);
  L.oS = o; o += lmpc_even(RS * d); L.oY = o; o += lmpc_even(RS * d); L.oISY = o; o += lmpc_even(RS * d);
  L.oX = o; o += lmpc_even(6 * d); L.oU = o; o += lmpc_even(2 * d);
  L.oDXA = o; o += lmpc_even(6 * d); L.oDUA = o; o += lmpc_even(2 * d);
  L.oDXF = o; o += lmpc_even(6 * d); L.oDUF = o; o += lmpc_even(2 * d);
  L.oCZX = o; o += lmpc_even(6 * d); L.oCZTH = o; o += lmpc_even(d); L.oGUD = o; o += lmpc_even(8 * d);
  L.oFAC = o; o += lmpc_even(20 * NS); L.oKFF = o; o += lmpc_even(6 * NS);
  L.oBL = o; o += lmpc_even(d); L.oBR = o; o += lmpc_even(d); L.oVREF = o; o += lmpc_even(d); L.oIT = o; o += lmpc_even(d);
  L.oPM = o; o += 64; L.oL1 = o; o += 8; L.oLTH = o; o += 8; L.oMAB = o; o += 48; L.oAXBW = o; o += 16; L.oYY = o; o += 64;
  L.oRED = o; o += lmpc_even(NW > 1 ? NW * LMPC_NRED : 0);
  L.oTERM = o; o += lmpc_even(LMPC_TB_SIZE_);
  L.oROWS = o; o += lmpc_even(LMPC_ROWS_DOUBLES);   // row table (copied from the parameters: constant-bank indexing is slow)
  L.oMBAR = o; o += 2;   // mbarrier of the bulk stage-in of [A|B|g]
  L.oCHS = o; o += 10;   // channel scales of the step-based acceptance test
  L.total = o;
  return L;
}

struct LmpcQpParams {
  int N, NS, K, learning, soft, hull_slack;
  int nh, hidx[6];
  double Einv[6], chs[6];
  int nxb, xb_c[12];
  double xb_sg[12], xb_h[12];
  int xslot[6][2];               // slot of the (hi, lo) box row of state c, -1 if unbounded
  int RS;                        // row slots per stage = nxb + 10
  double ulo[2], uhi[2], dlo[2], dhi[2];
  int ub_act[4], db_act[4];      // finite flags in slot order (c0 hi, c0 lo, c1 hi, c1 lo)
  double margin, qb;
  double qx[6], qxN[6];          // tracking weights, stage / terminal (x10)
  double Rm[3], Rd[3];
  int max_iter;
  double tol;
  int NSd;                       // odd stage stride of the [.][stage] arrays
  int NW;                        // warps per instance the layout was sized for
  LmpcRowTab rowtab;             // the inequality rows (copied to shared memory: constant-bank indexing is slow)
  LmpcLayout lay;                // shared-memory offsets (doubles)
};

// terminal-block scratch layout inside oTERM (doubles)
#define TB_PHI 0                       // 6x6
#define TB_PHIC (TB_PHI + 36)          // 6 x NQ
#define TB_S2 (TB_PHIC + 6 * LMPC_NQ)  // NQ x NQ (LU)
#define TB_XQ (TB_S2 + LMPC_NQ * LMPC_NQ)  // NQ x 6
#define TB_Q0 (TB_XQ + 6 * LMPC_NQ)    // NQ
#define TB_PT (TB_Q0 + LMPC_NQ)        // 6x6
#define TB_PTV (TB_PT + 36)            // 6
#define TB_BCOL (TB_PTV + 6)           // MB x 6  centred basic columns (compacted components)
#define TB_BD (TB_BCOL + 6 * LMPC_MB)  // MB      y/lambda of basic columns
#define TB_BG (TB_BD + LMPC_MB)        // MB      g_lambda of basic columns
#define TB_PIV (TB_BG + LMPC_MB)       // NQ (as doubles)
#define TB_EQ (TB_PIV + LMPC_NQ + 1)   // 6 + NQ   e and q of the terminal directions (computed by 11 lanes, read by all)
#define TB_SIZE (TB_EQ + 6 + LMPC_NQ)
static_assert(TB_SIZE == LMPC_TB_SIZE_, "terminal scratch size");

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
  const double* ssj;     // K   raw cost-to-go J  (J - J[0] is formed here, racing_mpc.cpp:280)
  const double* cen;     // 6   centre for the columns (the query point X_ref[:, N-1])
  int ss_count;          // 0 => no safe set
  double* scratch;       // LMPC_QP_SCRATCH(N, K) doubles of global memory: centred safe-set columns [K][6] (re-read where needed
                         // instead of living in registers) and the iterate saved before the polish (read back only if it fails)
};
#define LMPC_QP_SCRATCH(N, K) (8 * (N) + 8 * (K) + 6 * LMPC_MAX_SS_PTS + 16)

struct LmpcQpOut {
  double* X;       // N x 6
  double* U;       // NS x 2
  double* dU;      // NS x 2
  double* lam;     // K (may be null)
  double* cost;    // 1 (may be null)
  int* status;     // 1
  int* iters;      // 1
  int* stats;      // optional [4] (diagnostics; null in the product): interior-point iterations, polish rounds, polish attempts, unused
};

struct alignas(16) LmpcD2 { double x, y; };   // 16-byte shared-memory loads (LDS.128)
struct ArrK { double a[LMPC_KPL_MAX]; };
// same access syntax as LaneVar<ArrK> -- v(lane).a[p] -- for per-column values that live in global scratch instead of
// registers (element lane + NT p of a LMPC_MAX_SS_PTS-long array): the step buffers of the lambda block
struct ScrCol { double* b; int stride; LMPC_HDM double& operator[](int p) const { return b[stride * p]; } };
struct ScrLane { ScrCol a; };
template <int NT> struct ScrK { double* base; LMPC_HDM ScrLane operator()(int lane) const { return ScrLane{ScrCol{base + lane, NT}}; } };
struct ArrKx6 { double a[LMPC_KPL_MAX][6]; };
struct ArrKi { int a[LMPC_KPL_MAX]; };

// ---------------------------------------------------------------------------------------- rows
// Row work is organised by *variable*: group g = 0..5 is state c = g (box pair; for e_y also the two boundary rows),
// 6..7 control c = g - 6 (box pair), 8..9 rate c = g - 8 (box pair).  LMPC_FOR_ROWS(i, NA, NF) visits the rows of stage i:
// the variable's value in the iterate (v), in the affine step (va, when NA) and in the final step (vf, when NF) is formed
// once per group, then the unrolled side loop runs ROW_BODY for every existing row of it.  The caller defines three
// object-like macros before the expansion (and undefines them after):
//   ROW_GBEGIN   once per (group, stage), before its rows        in scope: g, i, v, va, vf
//   ROW_BODY     once per existing row                           in scope: + slot, sg, bnd, isb (constant: boundary row)
//   ROW_GEND     once per (group, stage), after its rows
// The value of the row on a vector is  sg * v - (isb ? theta-component : 0).
#define FOR_MY_STAGES(i) for (int i = lane; i < N; i += NT)
#define LMPC_ROW_SIDES_(SLOTS, K2, BNDEXPR, ISB)                                                         \
  LMPC_UNROLL                                                                                            \
  for (int side_ = 0; side_ < 2; side_++) {                                                              \
    const int slot = (SLOTS)[(K2) + side_];                                                              \
    if (slot >= 0) {                                                                                     \
      const double sg = side_ ? -1.0 : 1.0;                                                              \
      const double bnd = (BNDEXPR);                                                                      \
      constexpr bool isb = (ISB);                                                                        \
      constexpr bool on = true;                                                                          \
      (void)sg; (void)bnd; (void)isb; (void)on;                                                          \
      ROW_BODY                                                                                           \
    }                                                                                                    \
  }
#define LMPC_FOR_ROWS(i, NA, NF)                                                                         \
  {                                                                                                      \
    LMPC_NOUNROLL                                                                                        \
    for (int c_ = 0; c_ < 6; c_++) {                                                                     \
      const int g = c_;                                                                                  \
      const double v = X[c_ * d + (i)], va = (NA) ? DXA[c_ * d + (i)] : 0.0, vf = (NF) ? DXF[c_ * d + (i)] : 0.0; \
      (void)g; (void)v; (void)va; (void)vf;                                                              \
      ROW_GBEGIN                                                                                         \
      if ((i) >= 1 && (i) <= N - 2) { LMPC_ROW_SIDES_(RT->xslot, 2 * c_, RT->xbnd[2 * c_ + side_], false) } \
      if (c_ == 1 && (i) >= RT->ib0) { LMPC_ROW_SIDES_(RT->bslot, 0, side_ ? -(BR[i] + P.margin) : (BL[i] - P.margin), true) } \
      ROW_GEND                                                                                           \
    }                                                                                                    \
    if ((i) <= N - 2) {                                                                                  \
      LMPC_NOUNROLL                                                                                      \
      for (int c_ = 0; c_ < 2; c_++) {                                                                   \
        const int g = 6 + c_;                                                                            \
        const double v = U[c_ * d + (i)], va = (NA) ? DUA[c_ * d + (i)] : 0.0, vf = (NF) ? DUF[c_ * d + (i)] : 0.0; \
        (void)g; (void)v; (void)va; (void)vf;                                                            \
        ROW_GBEGIN                                                                                       \
        LMPC_ROW_SIDES_(RT->uslot, 2 * c_, RT->ubnd[2 * c_ + side_], false)                              \
        ROW_GEND                                                                                         \
      }                                                                                                  \
      const double it_ = IT[i];                                                                          \
      LMPC_NOUNROLL                                                                                      \
      for (int c_ = 0; c_ < 2; c_++) {                                                                   \
        const int g = 8 + c_;                                                                            \
        const double v = (U[c_ * d + (i)] - ((i) ? U[c_ * d + (i) - 1] : (c_ ? uic[1] : uic[0]))) * it_; \
        const double va = (NA) ? (DUA[c_ * d + (i)] - ((i) ? DUA[c_ * d + (i) - 1] : 0.0)) * it_ : 0.0;  \
        const double vf = (NF) ? (DUF[c_ * d + (i)] - ((i) ? DUF[c_ * d + (i) - 1] : 0.0)) * it_ : 0.0;  \
        (void)g; (void)v; (void)va; (void)vf;                                                            \
        ROW_GBEGIN                                                                                       \
        LMPC_ROW_SIDES_(RT->dslot, 2 * c_, RT->dbnd[2 * c_ + side_], false)                              \
        ROW_GEND                                                                                         \
      }                                                                                                  \
    }                                                                                                    \
  }

// One lane per row: lane = slot + SL * half; the lane walks the stages of its half of the horizon (SL = 16 and two halves
// when the configuration has at most 16 rows per stage, else 32 and one).  All lanes run the same straight-line code --
// the variable is fetched through the row's descriptor, rows that do not exist at a stage are masked by `on` -- so a
// pass costs (stages per half) bodies instead of (groups x sides) bodies.  One warp per instance only.
// In scope for ROW_BODY: slot, i (a valid stage even when masked), on, sg, bnd, isb, v, va, vf.
#define LMPC_FOR_ROWS_BY_LANE(NA, NF)                                                                    \
  {                                                                                                      \
    const int SL_ = (RSN <= 16) ? 16 : 32;                                                               \
    const int slot_ = lane & (SL_ - 1), half_ = lane / SL_, HS_ = (RSN <= 16) ? (N + 1) / 2 : N;         \
    const int sd_ = RT->sdesc[slot_ < RSN ? slot_ : 0];                                                  \
    const bool rowon_ = slot_ < RSN && (sd_ & 1024) != 0;                                                \
    const int slot = slot_ < RSN ? slot_ : 0;                                                            \
    const double bndc_ = RT->sbnd[slot];                                                                 \
    const int c8_ = sd_ & 15, i0_ = (sd_ >> 8) & 1, i1_ = N - 2 + ((sd_ >> 9) & 1);                      \
    const bool neg_ = (sd_ & 16) != 0, isb = (sd_ & 32) != 0, rate_ = (sd_ & 64) != 0;                   \
    const double sg = neg_ ? -1.0 : 1.0;                                                                 \
    const double uicl_ = (c8_ == 7) ? uic[1] : uic[0];                                                   \
    LMPC_UNROLL2                                                                                         \
    for (int it_ = 0; it_ < HS_; it_++) {                                                                \
      const int iraw_ = half_ * HS_ + it_;                                                               \
      const bool on = rowon_ && iraw_ >= i0_ && iraw_ <= i1_;                                            \
      const int i = on ? iraw_ : i0_;                                                                    \
      const int ip_ = i ? i - 1 : 0;                                                                     \
      const double sc_ = rate_ ? IT[i] : 1.0;                                                            \
      const double cur_ = X[c8_ * d + i], prv_ = X[c8_ * d + ip_];                                       \
      const double v = rate_ ? (cur_ - (i ? prv_ : uicl_)) * sc_ : cur_;                                 \
      double va = 0.0, vf = 0.0;                                                                         \
      if (NA) { const double ca_ = DXA[c8_ * d + i], pa_ = DXA[c8_ * d + ip_]; va = rate_ ? (ca_ - (i ? pa_ : 0.0)) * sc_ : ca_; } \
      if (NF) { const double cf_ = DXF[c8_ * d + i], pf_ = DXF[c8_ * d + ip_]; vf = rate_ ? (cf_ - (i ? pf_ : 0.0)) * sc_ : cf_; } \
      const double bl_ = BL[i], br_ = BR[i];                                                             \
      const double bnd = isb ? (neg_ ? -(br_ + P.margin) : (bl_ - P.margin)) : bndc_;                    \
      (void)v; (void)va; (void)vf; (void)bnd; (void)sg; (void)on;                                        \
      ROW_BODY                                                                                           \
    }                                                                                                    \
  }

// ------------------------------------------------------------------------------------------------
// NW warps per instance, KPL = ceil(K / (32 NW)) safe-set columns per lane (registers).
// NTPL / RSTPL > 0: horizon and rows-per-stage are compile-time (addresses fold to immediates); 0 = runtime.
template <int NW, int KPL, int NTPL, int RSTPL>
LMPC_DEV void lmpc_qp_solve(const LmpcQpParams& P, const LmpcQpIn& in, double* sm, const LmpcQpOut& out) {
  constexpr int NT = 32 * NW;
  constexpr bool FIXED = NTPL > 0;
  constexpr LmpcLayout LC = lmpc_layout(FIXED ? NTPL : 4, FIXED ? RSTPL : 10, NW);
#define LO(f) (FIXED ? LC.f : P.lay.f)
  const int N = FIXED ? NTPL : P.N, NS = N - 1, d = N | 1;
  const bool learn = P.learning != 0, soft = P.soft != 0;
  // Columns beyond the number actually found are copies of the last one (racing_mpc.cpp:263-272); they are
  // dropped here (their lambda stays 0): same optimum in X, U, dU, SS*lambda, and the explicit-column
  // system stays non-singular.
  const int K = learn ? ((in.ss_count > 0 && in.ss_count < P.K) ? in.ss_count : P.K) : 0;
  const int nh = P.nh;
  double* ABG = sm + LO(oABG);
  double* RSs = sm + LO(oS); double* RSy = sm + LO(oY); double* RSi = sm + LO(oISY);
  double* X = sm + LO(oX); double* U = sm + LO(oU);
  double* DXA = sm + LO(oDXA); double* DUA = sm + LO(oDUA); double* DXF = sm + LO(oDXF); double* DUF = sm + LO(oDUF);
  double* HX = DXF;   // alias: the Hessian diagonal is dead once the pass-0 factorisation is done
  double* CZX = sm + LO(oCZX); double* CZTH = sm + LO(oCZTH); double* GUD = sm + LO(oGUD);
  const LmpcRowTab* RT = reinterpret_cast<const LmpcRowTab*>(sm + LO(oROWS));
  const int RSN = FIXED ? RSTPL : P.RS;   // row slots per stage
  double* FAC = sm + LO(oFAC);   // per stage: Kz[16] (2x8 row-major), Sinv[3], pad
  double* KFF = sm + LO(oKFF);   // per stage: kff1[2], kffth[2], Cwth[2]
  double* BL = sm + LO(oBL); double* BR = sm + LO(oBR); double* VREF = sm + LO(oVREF); double* IT = sm + LO(oIT);
  double* PM = sm + LO(oPM); double* L1 = sm + LO(oL1); double* LTH = sm + LO(oLTH);
  double* MAB = sm + LO(oMAB); double* AXBW = sm + LO(oAXBW); double* YY = sm + LO(oYY);
  // scratch of the group reductions: one warp uses the (then idle) YY block of the sweep, NW warps their own NW x NRED area
  double* RED = (NW == 1) ? (sm + LO(oYY)) : (sm + LO(oRED));
  double* TB = sm + LO(oTERM);
// LMPC_FREE_THETA: the boundary slack sigma_b >= 0 (racing_mpc.cpp: slack of the soft track boundary, cost q_b sigma_b^2)
// carried as a FREE variable.  With q_b > 0 the bound is redundant -- a negative slack tightens every boundary row and
// costs q_b sigma_b^2, so it is never optimal -- and the optimum is the same; what goes away is a complementarity pair that
// is degenerate whenever no boundary row is active (sigma_b* = 0 with multiplier 0), i.e. on most instances, which made
// the interior point crawl (sigma_b shrinking 2.5x per iteration while mu fell 10x).
#ifndef LMPC_FREE_THETA
#define LMPC_FREE_THETA 1
#endif
#ifndef LMPC_MU0
#define LMPC_MU0 0.1
#endif
#ifndef LMPC_SFLOOR
#define LMPC_SFLOOR 1e-2
#endif
  const double sfloor = LMPC_SFLOOR, mu0 = LMPC_MU0, th0 = 0.01;

  // ---------------------------------------------------------------- load
  // [A|B|g] of all stages (54 (N-1) doubles, contiguous per instance): one bulk asynchronous copy (TMA) that lands while
  // the group stages the small inputs below
  group_bulk_load_begin(ABG, in.ABg, 54 * NS, sm + LO(oMBAR));
  GLANES_BEGIN(NT)
    for (int idx = lane; idx < LMPC_ROWS_DOUBLES; idx += NT) (sm + LO(oROWS))[idx] = reinterpret_cast<const double*>(&P.rowtab)[idx];
    for (int i = lane; i < N; i += NT) {
      BL[i] = in.bl[i]; BR[i] = in.br[i]; VREF[i] = in.vref[i];
      if (i < NS) {
        IT[i] = 1.0 / in.T[i];
        for (int c = 0; c < 2; c++) U[c * d + i] = fmin(fmax(in.U0[2 * i + c], P.ulo[c]), P.uhi[c]);
      }
    }
    if (lane < 6) X[lane * d] = in.x_ic[lane];
  GLANES_END(NW)
  const double uic[2] = {in.u_ic[0], in.u_ic[1]};
  const double zero2[2] = {0.0, 0.0};

  int status = LMPC_MAX_ITER;
  {
    bool bad = false;   // negated comparisons: a NaN input is rejected here
    for (int c = 0; c < 6; c++) if (!(in.x_ic[c] == in.x_ic[c])) bad = true;
    for (int sl = 0; sl < P.nxb; sl++) if (!(P.xb_sg[sl] * in.x_ic[P.xb_c[sl]] <= P.xb_h[sl])) bad = true;
    if (!soft && !(in.x_ic[1] <= in.bl[0] - P.margin && in.x_ic[1] >= in.br[0] + P.margin)) bad = true;
    if (bad) status = LMPC_INFEASIBLE_IC;
    else if (learn && in.ss_count <= 0) status = LMPC_NO_SAFE_SET;
  }
  int it = 0;

  // ---------------------------------------------------------------- linear rollout from x_ic
  group_bulk_load_wait(sm + LO(oMBAR));
  for (int i = 0; i < NS; i++) {
    GLANES_BEGIN(NT)
      if (lane < 6) {
        const double* A = ABG + 54 * i; const double* B = A + 36; const double* g = A + 48;
        double a = g[lane];
        for (int k = 0; k < 6; k++) a += A[lane + 6 * k] * X[k * d + i];
        for (int k = 0; k < 2; k++) a += B[lane + 6 * k] * U[k * d + i];
        X[lane * d + i + 1] = a;
      }
    GLANES_END(NW)
  }

  // A rollout that leaves the neighbourhood of the linearisation is useless as a start and would set the channel scales
  // of the stopping test (explicit Euler on the stiff lateral dynamics is unstable at dt = 0.025: |x| reaches 1e5 over 19
  // stages).  Any channel beyond 100 x max(1, |x_ic|, |X_ref[N-1]|) => roll out again with controls chosen stage by stage
  // (2x2 least squares, clipped into the box) so that v_x, v_y, omega follow the chord x_ic -> X_ref[N-1]; the iterate
  // stays dynamically feasible.  Never taken on the shipped (RK4) configurations.  Same rule in oracle/oracle_port.c.
  {
    LaneVar<double, NT> dv[1];
    GLANES_BEGIN(NT)
      double far = 0.0;
      if (lane < 6) {
        const double lim = 100.0 * fmax(1.0, fmax(fabs(in.x_ic[lane]), fabs(in.cen[lane])));
        for (int i = 1; i < N; i++) if (!(fabs(X[lane * d + i]) <= lim)) far = 1.0;
      }
      dv[0](lane) = far;
    GLANES_END(NW)
    const int op1[1] = {LMPC_RED_MAX};
    group_reduce<NW, 1>(dv, op1, RED);
    if (dv[0](0) > 0.0) {
      const double ich = 1.0 / (double)(N - 1);
      for (int i = 0; i < NS; i++) {
        GLANES_BEGIN(NT)
          if (lane < 2) {
            const double* A = ABG + 54 * i; const double* B = A + 36; const double* g = A + 48;
            double M00 = 1e-12, M01 = 0.0, M11 = 1e-12, b0 = 0.0, b1 = 0.0;
            for (int c = 3; c < 6; c++) {
              double a = g[c] - (in.x_ic[c] + (in.cen[c] - in.x_ic[c]) * ((double)(i + 1) * ich));
              for (int k = 0; k < 6; k++) a += A[c + 6 * k] * X[k * d + i];
              M00 += B[c] * B[c]; M01 += B[c] * B[c + 6]; M11 += B[c + 6] * B[c + 6];
              b0 -= B[c] * a; b1 -= B[c + 6] * a;
            }
            const double det = M00 * M11 - M01 * M01;
            double uu = lane ? (M00 * b1 - M01 * b0) / det : (M11 * b0 - M01 * b1) / det;
            if (!(uu <= P.uhi[lane])) uu = P.uhi[lane];
            if (!(uu >= P.ulo[lane])) uu = P.ulo[lane];
            U[lane * d + i] = uu;
          }
        GLANES_END(NW)
        GLANES_BEGIN(NT)
          if (lane < 6) {
            const double* A = ABG + 54 * i; const double* B = A + 36; const double* g = A + 48;
            double a = g[lane];
            for (int k = 0; k < 6; k++) a += A[lane + 6 * k] * X[k * d + i];
            for (int k = 0; k < 2; k++) a += B[lane + 6 * k] * U[k * d + i];
            X[lane * d + i + 1] = a;
          }
        GLANES_END(NW)
      }
    }
  }

  // ---------------------------------------------------------------- initial slacks / multipliers, channel scales
  // A start whose rollout leaves the track: the boundary slack sigma_b absorbs the violation from the beginning (boundary
  // rows strictly feasible at the start) instead of being dragged there by an infeasible-start crawl -- IAC tracking
  // max 16 -> 9 iterations, IAC LMPC N = 60 mean 32 -> 18; untouched when the rollout stays inside (all BARC cases).
  double th_start = th0;
  if (soft) {
    LaneVar<double, NT> bv;
    GLANES_BEGIN(NT)
      double vmax = -1e300;
      for (int i = lane; i < N; i += NT) if (i >= RT->ib0) {
        const double ey = X[d + i];
        vmax = lmpc_max(vmax, lmpc_max(ey - (BL[i] - P.margin), (BR[i] + P.margin) - ey));
      }
      bv(lane) = vmax;
    GLANES_END(NW)
    group_max<NW>(bv, RED);
    if (bv(0) + 0.1 > th_start) th_start = bv(0) + 0.1;
  }
  double th = th_start, yth = LMPC_FREE_THETA ? 0.0 : mu0 / th_start, dth = 0.0, dyth = 0.0, dtha = 0.0, dytha = 0.0;
  LaneVar<ArrK, NT> lam, ylam, omg_;   // lambda block: iterate and weights in registers ...
  double* const SCRK = in.scratch + 6 * P.K + 8 * P.N + 2 * P.K;
  const ScrK<NT> dla{SCRK}, dya{SCRK + LMPC_MAX_SS_PTS}, dlf{SCRK + 2 * LMPC_MAX_SS_PTS}, dyf{SCRK + 3 * LMPC_MAX_SS_PTS},
      glam{SCRK + 4 * LMPC_MAX_SS_PTS}, sscv{SCRK + 5 * LMPC_MAX_SS_PTS};   // ... steps, gradient, cost-to-go in global scratch
  double* const chs_ = sm + LO(oCHS);   // [10] channel scales
  LaneVar<ArrKi, NT> isB;
  double* const ST = in.scratch;   // [K][6] centred columns, compacted to the nh hull components (global memory; L2-resident)
  double R0;
  int m_total = 0;
  {
    LaneVar<double, NT> r[12];
    GLANES_BEGIN(NT)
      double r0 = 1.0, cnt = 0.0;
      const double thq = soft ? th : 0.0;
#define ROW_GBEGIN
#define ROW_GEND
#define ROW_BODY                                                           \
  {                                                                        \
    const double slack = bnd - (sg * v - (isb ? thq : 0.0));               \
    const double s = slack > sfloor ? slack : sfloor;                      \
    const double y = mu0 / s;                                              \
    RSs[slot * d + i] = s; RSy[slot * d + i] = y; RSi[slot * d + i] = 1.0 / (s * y); \
    r0 = fmax(r0, y); cnt += 1.0;                                          \
  }
      FOR_MY_STAGES(i) LMPC_FOR_ROWS(i, false, false)
#undef ROW_BODY
      const double j0 = (learn && K > 0) ? in.ssj[0] : 0.0;
      for (int p = 0; p < KPL; p++) {
        const int k = lane + NT * p;
        const bool on = learn && k < K;
        lam(lane).a[p] = on ? 1.0 / K : 0.0; ylam(lane).a[p] = on ? mu0 * K : 0.0;
        dla(lane).a[p] = 0.0; dya(lane).a[p] = 0.0; dlf(lane).a[p] = 0.0; dyf(lane).a[p] = 0.0;
        glam(lane).a[p] = 0.0; omg_(lane).a[p] = 0.0; isB(lane).a[p] = 0;
        sscv(lane).a[p] = on ? in.ssj[k] - j0 : 0.0;
        if (on) for (int a = 0; a < 6; a++) ST[6 * k + a] = (a < nh) ? in.ssx[6 * k + P.hidx[a]] - in.cen[P.hidx[a]] : 0.0;
        if (on) r0 = fmax(r0, fabs(sscv(lane).a[p]));
      }
      double m[10] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
      for (int i = lane; i < N; i += NT) {
        for (int c = 0; c < 6; c++) m[c] = fmax(m[c], fabs(X[c * d + i]));
        if (i < NS) for (int c = 0; c < 2; c++) {
          const double up = i ? U[c * d + i - 1] : uic[c];
          m[6 + c] = fmax(m[6 + c], fabs(U[c * d + i]));
          m[8 + c] = fmax(m[8 + c], fabs(U[c * d + i] - up) * IT[i]);
        }
      }
      r[0](lane) = r0; r[1](lane) = cnt;
      for (int q = 0; q < 10; q++) r[2 + q](lane) = m[q];
    GLANES_END(NW)
    const int ops[12] = {LMPC_RED_MAX, LMPC_RED_SUM, LMPC_RED_MAX, LMPC_RED_MAX, LMPC_RED_MAX, LMPC_RED_MAX, LMPC_RED_MAX,
                         LMPC_RED_MAX, LMPC_RED_MAX, LMPC_RED_MAX, LMPC_RED_MAX, LMPC_RED_MAX};
    group_reduce<NW, 12>(r, ops, RED);
    R0 = r[0](0);
    m_total = (int)(r[1](0) + 0.5) + ((soft && !LMPC_FREE_THETA) ? 1 : 0) + K;
    for (int q = 0; q < 10; q++) chs_[q] = 1.0 / r[2 + q](0);
  }
  if (soft) R0 = fmax(R0, 2.0 * P.qb * th);
  double rho_d = 1.0, prev_stepn = 0.0, mu_m1 = 1e300, mu_m2 = 1e300, mu_m3 = 1e300;
  const double inv_m = 1.0 / (double)m_total;
  // Active-set polish (what OSQP's polish=true does for the reference, racing_mpc.cpp:90-95): once the interior
  // point has identified the active set, one augmented-Lagrangian Newton step on it -- active rows get the
  // weight rho and the gradient y + rho (g'v - h), inactive rows are dropped, basic safe-set columns are free --
  // followed by up to PMAX active-set refinements.  During the polish the row arrays are re-purposed:
  // sign(RSs) < 0 marks an active row, RSy holds the multiplier estimate, RSi keeps the saved y.
  int polishing = 0, polish_tries = 0, classified = 0, n_polish_rounds = 0;
  double tol_step = P.tol, tol_mu = 0.1 * P.tol;   // complementarity level at which the polish takes over
  int numfail_polish = 0;
  double prev_changed = 0.0;   // rows that changed side in the previous polish round
  double* const SAVE_XU = in.scratch + 6 * P.K;          // [N][8] saved iterate (restored if the polish fails); global memory, not registers
  double* const SAVE_L = in.scratch + 6 * P.K + 8 * N;   // [K] lambda, [K] its multiplier
  double th_save = 0.0, yth_save = 0.0;
  int pact_th = 0;
  LaneVar<ArrKi, NT> pnb;

  // ================================================================ interior-point iterations
  for (; status == LMPC_MAX_ITER && (it < P.max_iter || polishing); it++) {
    double sigma = 0.0, alpha = 1.0, Pithth_keep = 0.0, csc = 1.0, mu = 0.0, rpn = 0.0, rnu = 0.0;
    double sig[6] = {0, 0, 0, 0, 0, 0};
    bool fail = false, converged = false, restart = false;
    if (polishing && !classified) {
      // ---------- classify from the interior-point iterate, save what a failed polish must restore
      GLANES_BEGIN(NT)
#define ROW_BODY                                                           \
  {                                                                        \
    const double s = RSs[slot * d + i], y = RSy[slot * d + i];             \
    RSi[slot * d + i] = y;                                                 \
    if (y > s) RSs[slot * d + i] = -s; else RSy[slot * d + i] = 0.0;       \
  }
        FOR_MY_STAGES(i) LMPC_FOR_ROWS(i, false, false)
#undef ROW_BODY
        FOR_MY_STAGES(i) { for (int c = 0; c < 6; c++) SAVE_XU[8 * i + c] = X[c * d + i]; for (int c = 0; c < 2; c++) SAVE_XU[8 * i + 6 + c] = (i < NS) ? U[c * d + i] : 0.0; }
        for (int p = 0; p < KPL; p++) {
          if (lane + NT * p < K) { SAVE_L[lane + NT * p] = lam(lane).a[p]; SAVE_L[K + lane + NT * p] = ylam(lane).a[p]; }
          const bool basic = lam(lane).a[p] >= ylam(lane).a[p];
          pnb(lane).a[p] = basic ? 0 : 1;
          if (basic) ylam(lane).a[p] = 0.0;
        }
      GLANES_END(NW)
      th_save = th; yth_save = yth;
      pact_th = (!LMPC_FREE_THETA && soft && yth > th) ? 1 : 0;
      if (soft && !pact_th) yth = 0.0;
      classified = 1; polish_tries++; prev_changed = 0.0;
    }
    for (int pass = 0; pass < (polishing ? 1 : 2) && !fail; pass++) {
      const double smu = sigma * mu;
      // ---------- rows -> per-stage Hessian / gradient pieces (one lane per (group, stage))
      LaneVar<double, NT> rs[11], rmx;
      GLANES_BEGIN(NT)
        double dth_acc = 0.0, cth_acc = 0.0, msum = 0.0, rpm = 0.0;
        const double thq = soft ? th : 0.0, dthaq = soft ? dtha : 0.0;
        const bool ip0 = !pass && !polishing;   // interior-point predictor pass: complementarity and residual are measured here
#undef ROW_GBEGIN
#undef ROW_GEND
#define ROW_GBEGIN                                                         \
  double hsum = 0.0, gsum = 0.0, cz_th = 0.0;                              \
  if (g < 6 && !learn) { const double w = (i == N - 1) ? P.qxN[g] : P.qx[g]; hsum = 2.0 * w; gsum = 2.0 * w * (v - (g == 3 ? VREF[i] : 0.0)); }
#define ROW_BODY                                                           \
  {                                                                        \
    /* straight-line: the three variants (predictor, corrector with the second-order term of the affine step   \
       ds_a * dy_a, dy_a = -y - y ds_a / s, and the polish) are selects on group-uniform flags, not branches */ \
    const double s = RSs[slot * d + i], y = RSy[slot * d + i], isy = RSi[slot * d + i]; \
    const double is = y * isy;                                             \
    const double dj0 = y * is;                                             \
    const double rp = (sg * v - (isb ? thq : 0.0)) + s - bnd;              \
    const double dsa = -rp - (sg * va - (isb ? dthaq : 0.0));              \
    const double dya_ = -y - dj0 * dsa;                                    \
    const double tip = dj0 * rp + (pass ? (smu - csc * dsa * dya_) * is : 0.0); \
    const bool act = s < 0.0;                                              \
    const double dj = polishing ? (act ? LMPC_PRHO : 0.0) : dj0;           \
    const double t = polishing ? (act ? y + LMPC_PRHO * (rp - s) : 0.0) : tip; \
    msum += ip0 ? s * y : 0.0;                                             \
    rpm = ip0 ? lmpc_max(rpm, fabs(rp)) : rpm;                             \
    hsum += dj; gsum += sg * t;                                            \
    if (isb && soft) { cz_th += -sg * dj; dth_acc += dj; cth_acc += -t; }  \
  }
#define ROW_GEND                                                           \
  if (g < 6) {                                                             \
    if (!pass) HX[g * d + i] = hsum;                                       \
    CZX[g * d + i] = gsum;                                                 \
    if (g == 1) CZTH[i] = cz_th;                                           \
  } else {                                                                 \
    /* GUD rows: 0,1 dub  2,3 tub  4,5 dd  6,7 td */                       \
    const int cc_ = (g - 6) & 1, base_ = (g < 8) ? 0 : 4;                  \
    if (!pass) GUD[(base_ + cc_) * d + i] = hsum;                          \
    GUD[(base_ + 2 + cc_) * d + i] = gsum;                                 \
  }
        FOR_MY_STAGES(i) LMPC_FOR_ROWS(i, pass != 0, false)
#undef ROW_BODY
#undef ROW_GBEGIN
#undef ROW_GEND
#define ROW_GBEGIN
#define ROW_GEND
        double lsum = 0.0, sg6[6] = {0, 0, 0, 0, 0, 0}, nbasic = 0.0;
        if (!pass) for (int p = 0; p < KPL; p++) {
          const double l = lam(lane).a[p];
          msum += l * ylam(lane).a[p]; lsum += l;
          if (lane + NT * p < K) for (int a = 0; a < 6; a++) sg6[a] += ST[6 * (lane + NT * p) + a] * l;
          if (polishing && lane + NT * p < K && !pnb(lane).a[p]) nbasic += 1.0;
        }
        rs[0](lane) = dth_acc; rs[1](lane) = cth_acc; rs[2](lane) = msum; rs[3](lane) = nbasic; rs[4](lane) = lsum;
        for (int a = 0; a < 6; a++) rs[5 + a](lane) = sg6[a];
        rmx(lane) = rpm;
      GLANES_END(NW)
      if (pass == 0) {   // the sums of the iterate (11) + the largest primal residual; the corrector pass only needs the two theta sums
        group_reduce_sum<NW, 11>(rs, RED);
        group_max<NW>(rmx, RED);
      } else {
        LaneVar<double, NT>(&r2)[2] = reinterpret_cast<LaneVar<double, NT>(&)[2]>(rs[0]);
        group_reduce_sum<NW, 2>(r2, RED);
      }
      double Dthth = rs[0](0), cth = rs[1](0);
      if (!pass) {
        mu = (rs[2](0) + ((soft && !LMPC_FREE_THETA) ? th * yth : 0.0)) * inv_m;
        rpn = rmx(0);
        rnu = learn ? rs[4](0) - 1.0 : 0.0;
        if (learn) for (int a = 0; a < 6; a++) sig[a] = (a < nh) ? X[P.hidx[a] * d + N - 1] - in.cen[P.hidx[a]] - rs[5 + a](0) : 0.0;
        // (second alternative: the floor of double precision -- below mu = 1e-13 the products s y are rounding noise and the
        // primal residual cannot follow a tolerance tighter than its own 1e-13: a caller's tol < 1e-12 ends here)
        if (!polishing && ((mu < tol_mu && rpn < tol_mu && rho_d * R0 < tol_mu && fabs(rnu) < tol_mu) ||
                           (mu < 1e-13 && rpn < 1e-9 && rho_d * R0 < 1e-9 && fabs(rnu) < 1e-9))) {
          // complementarity floor reached: polish from here (restart the trip so that the rows are re-assembled)
          polishing = 1; classified = 0; restart = true; break;
        }
        // stalled interior point (Mehrotra limit cycle at a badly centred iterate: mu has not halved over three
        // iterations although the iterate is primal feasible): let the active-set polish decide from here
        if (!polishing && polish_tries < 2 && it >= 3 && mu < 1e-5 && rpn < 1e-8 && fabs(rnu) < 1e-8 && mu > 0.5 * mu_m3) {
          polishing = 1; classified = 0; restart = true; break;
        }
        if (!polishing) { mu_m3 = mu_m2; mu_m2 = mu_m1; mu_m1 = mu; }
        if (polishing && learn && (rs[3](0) > LMPC_MB + 0.5 || rs[3](0) < 0.5)) { fail = true; break; }   // at most MB free columns
      }
      double corr_th = 0.0;
      if (soft) {
        if (polishing) {
          Dthth += 2.0 * P.qb; cth += 2.0 * P.qb * th;
          if (pact_th) { Dthth += LMPC_PRHO; cth += -yth + LMPC_PRHO * th; }
        } else {
          if (pass) corr_th = csc * dtha * dytha;
          if (LMPC_FREE_THETA) { Dthth += 2.0 * P.qb; cth += 2.0 * P.qb * th; }
          else { const double ith = lmpc_rcp(th); Dthth += 2.0 * P.qb + yth * ith; cth += 2.0 * P.qb * th - (smu - corr_th) * ith; }
        }
      }

      // ---------- terminal value  P_{N-1}, l_{N-1}
      GLANES_BEGIN(NT)
        for (int o = lane; o < 64; o += NT) { const int r = o >> 3, c = o & 7; PM[o] = (r == c && r < 6) ? HX[r * d + N - 1] : 0.0; }
        if (lane < 8) { L1[lane] = lane < 6 ? CZX[lane * d + N - 1] : 0.0; LTH[lane] = (lane == 1) ? CZTH[N - 1] : 0.0; }
      GLANES_END(NW)
      if (learn) {
        // ---- per-column weights; pass 0: the MB largest Omega become explicit columns
        if (pass == 0) {
          GLANES_BEGIN(NT)
            for (int p = 0; p < KPL; p++) {
              const int k = lane + NT * p;
              double om = (k < K) ? lam(lane).a[p] * lmpc_rcp(ylam(lane).a[p]) : -1.0;
              if (polishing && k < K) om = pnb(lane).a[p] ? 1.0 / LMPC_PRHO : 1e300;   // free (basic) columns must be explicit
              omg_(lane).a[p] = om; isB(lane).a[p] = 0;
            }
          GLANES_END(NW)
          int kbs[LMPC_MB];
#pragma unroll
          for (int q = 0; q < LMPC_MB; q++) {
            LaneVar<double, NT> bv; LaneVar<int, NT> bi;
            GLANES_BEGIN(NT)
              double v = -2.0; int ix = 1 << 30;
#pragma unroll
              for (int p = 0; p < KPL; p++) {
                const int k = lane + NT * p;
                const bool take = k < K && !isB(lane).a[p] && omg_(lane).a[p] > v;
                v = take ? omg_(lane).a[p] : v; ix = take ? k : ix;
              }
              bv(lane) = v; bi(lane) = ix;
            GLANES_END(NW)
            group_argbest<NW>(bv, bi, true, RED);
            const int kb = bi(0);
            kbs[q] = kb;
            GLANES_BEGIN(NT)
#pragma unroll
              for (int p = 0; p < KPL; p++) if (lane + NT * p == kb) {
                isB(lane).a[p] = 1 + q;
                TB[TB_BD + q] = polishing ? (pnb(lane).a[p] ? LMPC_PRHO : 0.0) : ylam(lane).a[p] * lmpc_rcp(lam(lane).a[p]);
              }
            GLANES_END(NW)
          }
          // the MB explicit columns, one element per lane: a single round trip to the scratch instead of one per column
          GLANES_BEGIN(NT)
            if (lane < 6 * LMPC_MB) {
              const int q = lane / 6, a = lane - 6 * q;
              int kb = kbs[0];
#pragma unroll
              for (int r = 1; r < LMPC_MB; r++) kb = (q == r) ? kbs[r] : kb;
              // an iterate that has gone NaN (it ends as LMPC_NUMERIC) has no largest weight: the arg-max returns its
              // "none" index, which must not become an address
              kb = (kb >= 0 && kb < K) ? kb : 0;
              TB[TB_BCOL + lane] = ST[6 * kb + a];
            }
          GLANES_END(NW)
        }
        // ---- sums over the non-basic columns: b (6), og | W (21), a (6), om1
        LaneVar<double, NT> rt[LMPC_NRED];
        GLANES_BEGIN(NT)
          double W[21], av[6], bv6[6], om1 = 0.0, og = 0.0;
#pragma unroll
          for (int q = 0; q < 21; q++) W[q] = 0.0;
#pragma unroll
          for (int a = 0; a < 6; a++) { av[a] = 0.0; bv6[a] = 0.0; }
#pragma unroll
          for (int p = 0; p < KPL; p++) {
            const int k = lane + NT * p;
            if (k < K) {
              const double l = lam(lane).a[p];
              double tl = smu;
              if (pass) tl -= csc * dla(lane).a[p] * dya(lane).a[p];
              double gl = sscv(lane).a[p] - tl * lmpc_rcp(l);
              if (polishing) gl = pnb(lane).a[p] ? sscv(lane).a[p] - ylam(lane).a[p] + LMPC_PRHO * l : sscv(lane).a[p];
              glam(lane).a[p] = gl;
              // explicit (basic) columns hand their gradient to the LU system and enter the sums with weight zero
              const int ib = isB(lane).a[p];
              if (ib) TB[TB_BG + ib - 1] = gl;
              const double om = ib ? 0.0 : omg_(lane).a[p];
              og += om * gl; om1 += om;
              double sv[6];
#pragma unroll
              for (int a = 0; a < 6; a++) sv[a] = ST[6 * k + a];
#pragma unroll
              for (int a = 0, q = 0; a < 6; a++) {
                const double sa = sv[a];
                bv6[a] += sa * om * gl; av[a] += sa * om;
#pragma unroll
                for (int b = 0; b <= a; b++, q++) W[q] += om * sa * sv[b];
              }
            }
          }
#pragma unroll
          for (int a = 0; a < 6; a++) rt[a](lane) = bv6[a];
          rt[6](lane) = og;
#pragma unroll
          for (int q = 0; q < 21; q++) rt[7 + q](lane) = W[q];
#pragma unroll
          for (int a = 0; a < 6; a++) rt[28 + a](lane) = av[a];
          rt[34](lane) = om1; rt[35](lane) = 0.0;
        GLANES_END(NW)
        double r1[6];
        static_assert(LMPC_NRED == 36, "layout of the terminal reduction");
#ifdef LMPC_DEBUG_TRACE
        if (it == 0 && pass == 0) { LANE0_ONLY(if (LMPC_TRACE_COND) printf("  pre-reduce lane0: W00 %.9e a0 %.9e om1 %.9e | St %.6e %.6e %.6e om %.6e %.6e isB %d %d %d K %d\n", rt[7](0), rt[28](0), rt[34](0), ST[0], ST[6 * NT], ST[12 * NT], omg_(0).a[0], omg_(0).a[1], isB(0).a[0], isB(0).a[1], isB(0).a[2], K);) }
#endif
        if (pass == 0) {   // values that only change with the factorisation are reduced in pass 0 only
          group_reduce_sum<NW, LMPC_NRED>(rt, RED);
        } else {
          LaneVar<double, NT>(&r7)[7] = reinterpret_cast<LaneVar<double, NT>(&)[7]>(rt[0]);
          const int ops[7] = {LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM};
          group_reduce<NW, 7>(r7, ops, RED);
        }
#ifdef LMPC_DEBUG_TRACE
        if (it == 0 && pass == 0) { LANE0_ONLY(if (LMPC_TRACE_COND) printf("  post-reduce: W00 %.9e W10 %.9e a0 %.9e om1 %.9e b0 %.9e og %.9e\n", rt[7](0), rt[8](0), rt[28](0), rt[34](0), rt[0](0), rt[6](0));) }
#endif
#pragma unroll
        for (int a = 0; a < 6; a++) r1[a] = (a < nh) ? sig[a] + rt[a](0) : 0.0;
        const double og_all = rt[6](0);
        if (pass == 0) {
          // Cholesky of Einv + W_N, lower triangle packed (a,b) -> a(a+1)/2 + b, reciprocal diagonal; padded with I
          double Lc[21];
#pragma unroll
          for (int a = 0, q = 0; a < 6; a++)
#pragma unroll
            for (int b = 0; b <= a; b++, q++) Lc[q] = (a < nh && b < nh) ? rt[7 + q](0) + (a == b ? P.Einv[a] : 0.0) : (a == b ? 1.0 : 0.0);
          bool ok = true;
#pragma unroll
          for (int j = 0; j < 6; j++) {
            double dg = Lc[j * (j + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; k++) dg -= Lc[j * (j + 1) / 2 + k] * Lc[j * (j + 1) / 2 + k];
            if (!(dg > 0.0)) { ok = false; dg = 1.0; }
            const double il = lmpc_rsqrt(dg);
            Lc[j * (j + 1) / 2 + j] = il;
#pragma unroll
            for (int i2 = j + 1; i2 < 6; i2++) {
              double a2 = Lc[i2 * (i2 + 1) / 2 + j];
#pragma unroll
              for (int k = 0; k < j; k++) a2 -= Lc[i2 * (i2 + 1) / 2 + k] * Lc[j * (j + 1) / 2 + k];
              Lc[i2 * (i2 + 1) / 2 + j] = a2 * il;
            }
          }
          if (!ok) fail = true;
          // lanes solve in parallel: 0..5 identity columns (Phi), 6 -> -a_N, 7.. -> basic columns
          GLANES_BEGIN(NT)
            if (lane < 6 + LMPC_NQ) {
              double v[6];
#pragma unroll
              for (int a = 0; a < 6; a++) {
                if (lane < 6) v[a] = (a == lane) ? 1.0 : 0.0;
                else if (lane == 6) v[a] = -rt[28 + a](lane);
                else v[a] = TB[TB_BCOL + 6 * (lane - 7) + a];
              }
#pragma unroll
              for (int i2 = 0; i2 < 6; i2++) {
                double a2 = v[i2];
#pragma unroll
                for (int k = 0; k < i2; k++) a2 -= Lc[i2 * (i2 + 1) / 2 + k] * v[k];
                v[i2] = a2 * Lc[i2 * (i2 + 1) / 2 + i2];
              }
#pragma unroll
              for (int i2 = 5; i2 >= 0; i2--) {
                double a2 = v[i2];
#pragma unroll
                for (int k = i2 + 1; k < 6; k++) a2 -= Lc[k * (k + 1) / 2 + i2] * v[k];
                v[i2] = a2 * Lc[i2 * (i2 + 1) / 2 + i2];
              }
              if (lane < 6) {
#pragma unroll
                for (int a = 0; a < 6; a++) TB[TB_PHI + 6 * a + lane] = v[a];
              } else {
#pragma unroll
                for (int a = 0; a < 6; a++) TB[TB_PHIC + a * LMPC_NQ + (lane - 6)] = v[a];
              }
            }
          GLANES_END(NW)
          // S2 = Z - C' Phi C   (NQ x NQ)
          GLANES_BEGIN(NT)
            if (lane < LMPC_NQ * LMPC_NQ) {
              const int q = lane / LMPC_NQ, r = lane % LMPC_NQ;
              double z = 0.0;
              if (q == 0 && r == 0) z = rt[34](lane); else if (q == 0 || r == 0) z = -1.0; else if (q == r) z = -TB[TB_BD + q - 1];
              double s2 = 0.0;
#pragma unroll
              for (int a = 0; a < 6; a++) {
                const double cq = (q == 0) ? -rt[28 + a](lane) : TB[TB_BCOL + 6 * (q - 1) + a];
                s2 += cq * TB[TB_PHIC + a * LMPC_NQ + r];
              }
              TB[TB_S2 + q * LMPC_NQ + r] = z - s2;
            }
          GLANES_END(NW)
          // pivoted LU of S2 (uniform; lane 0 stores)
          {
            double M[LMPC_NQ][LMPC_NQ]; int piv[LMPC_NQ];
#pragma unroll
            for (int q = 0; q < LMPC_NQ; q++)
#pragma unroll
              for (int r = 0; r < LMPC_NQ; r++) M[q][r] = TB[TB_S2 + q * LMPC_NQ + r];
            GROUP_SYNC(NW);
#pragma unroll
            for (int k = 0; k < LMPC_NQ; k++) {
              int pk = k; double mx = fabs(M[k][k]);
#pragma unroll
              for (int i2 = k + 1; i2 < LMPC_NQ; i2++) if (fabs(M[i2][k]) > mx) { mx = fabs(M[i2][k]); pk = i2; }
              piv[k] = pk;
              if (!(mx > 0.0)) fail = true;
#pragma unroll
              for (int i2 = k + 1; i2 < LMPC_NQ; i2++) {
                const bool sw = (pk == i2);
#pragma unroll
                for (int c2 = 0; c2 < LMPC_NQ; c2++) { const double t0 = M[k][c2], t1 = M[i2][c2]; M[k][c2] = sw ? t1 : t0; M[i2][c2] = sw ? t0 : t1; }
              }
              const double ip = lmpc_rcp(fail ? 1.0 : M[k][k]);
#pragma unroll
              for (int i2 = k + 1; i2 < LMPC_NQ; i2++) {
                const double f = M[i2][k] * ip; M[i2][k] = f;
#pragma unroll
                for (int c2 = k + 1; c2 < LMPC_NQ; c2++) M[i2][c2] -= f * M[k][c2];
              }
            }
            LANE0_ONLY(
              for (int q = 0; q < LMPC_NQ; q++) { for (int r = 0; r < LMPC_NQ; r++) TB[TB_S2 + q * LMPC_NQ + r] = M[q][r]; TB[TB_PIV + q] = (double)piv[q]; })
            GROUP_SYNC(NW);
          }
        }
        // ---- solves with the LU: lanes 0..5 -> Xq columns (pass 0), lane 6 -> q0
        GLANES_BEGIN(NT)
          const bool doX = (pass == 0) && lane < nh;
          const bool doQ = lane == 6;
          if (doX || doQ) {
            double v[LMPC_NQ];
#pragma unroll
            for (int q = 0; q < LMPC_NQ; q++) {
              if (doX) v[q] = TB[TB_PHIC + lane * LMPC_NQ + q];
              else {
                double r2 = (q == 0) ? (rnu - og_all) : TB[TB_BG + q - 1];
#pragma unroll
                for (int a = 0; a < 6; a++) r2 -= TB[TB_PHIC + a * LMPC_NQ + q] * r1[a];
                v[q] = r2;
              }
            }
#pragma unroll
            for (int k = 0; k < LMPC_NQ; k++) {
              const int pk = (int)TB[TB_PIV + k];
#pragma unroll
              for (int i2 = k + 1; i2 < LMPC_NQ; i2++) { const bool sw = (pk == i2); const double t0 = v[k], t1 = v[i2]; v[k] = sw ? t1 : t0; v[i2] = sw ? t0 : t1; }
            }
#pragma unroll
            for (int i2 = 0; i2 < LMPC_NQ; i2++) {
              double a2 = v[i2];
#pragma unroll
              for (int k = 0; k < i2; k++) a2 -= TB[TB_S2 + i2 * LMPC_NQ + k] * v[k];
              v[i2] = a2;
            }
#pragma unroll
            for (int i2 = LMPC_NQ - 1; i2 >= 0; i2--) {
              double a2 = v[i2];
#pragma unroll
              for (int k = i2 + 1; k < LMPC_NQ; k++) a2 -= TB[TB_S2 + i2 * LMPC_NQ + k] * v[k];
              v[i2] = a2 * lmpc_rcp(TB[TB_S2 + i2 * LMPC_NQ + i2]);
            }
            if (doX) {
#pragma unroll
              for (int q = 0; q < LMPC_NQ; q++) TB[TB_XQ + q * 6 + lane] = v[q];
            } else {
#pragma unroll
              for (int q = 0; q < LMPC_NQ; q++) TB[TB_Q0 + q] = v[q];
            }
          }
        GLANES_END(NW)
        // ---- PT = Phi + PhiC Xq (pass 0),  pT = Phi r1 - PhiC q0 ; add into P_{N-1}, l_{N-1}
        GLANES_BEGIN(NT)
          if (pass == 0) {
            for (int o = lane; o < 36; o += NT) {
              const int a = o / 6, b = o % 6;
              double s2 = 0.0;
              if (a < nh && b < nh) {
                s2 = TB[TB_PHI + 6 * a + b];
#pragma unroll
                for (int q = 0; q < LMPC_NQ; q++) s2 += TB[TB_PHIC + a * LMPC_NQ + q] * TB[TB_XQ + q * 6 + b];
                PM[8 * P.hidx[a] + P.hidx[b]] += s2;
              }
              TB[TB_PT + o] = s2;
            }
          }
          if (lane >= NT - 6) {   // the last six lanes: pT  (disjoint outputs from the PT lanes above)
            const int a = lane - (NT - 6);
            double s2 = 0.0;
            if (a < nh) {
#pragma unroll
              for (int b = 0; b < 6; b++) s2 += TB[TB_PHI + 6 * a + b] * r1[b];
#pragma unroll
              for (int q = 0; q < LMPC_NQ; q++) s2 -= TB[TB_PHIC + a * LMPC_NQ + q] * TB[TB_Q0 + q];
              L1[P.hidx[a]] += s2;
            }
            TB[TB_PTV + a] = s2;
          }
        GLANES_END(NW)
      }
#ifdef LMPC_DEBUG_TRACE
      if (learn && it == 0) { LANE0_ONLY(if (LMPC_TRACE_COND) printf("  term pass %d: q0 %.9e %.9e %.9e %.9e %.9e | pT %.9e %.9e %.9e | PT00 %.9e PT55 %.9e | Phi00 %.9e S2_00 %.9e piv %g %g %g %g %g | BD %.6e %.6e %.6e %.6e BG %.6e %.6e | L1 %.9e %.9e PM00 %.9e\n", pass, TB[TB_Q0], TB[TB_Q0+1], TB[TB_Q0+2], TB[TB_Q0+3], TB[TB_Q0+4], TB[TB_PTV], TB[TB_PTV+1], TB[TB_PTV+5], TB[TB_PT], TB[TB_PT+35], TB[TB_PHI], TB[TB_S2], TB[TB_PIV], TB[TB_PIV+1], TB[TB_PIV+2], TB[TB_PIV+3], TB[TB_PIV+4], TB[TB_BD], TB[TB_BD+1], TB[TB_BD+2], TB[TB_BD+3], TB[TB_BG], TB[TB_BG+1], L1[0], L1[3], PM[0]);) }
#endif
      if (fail) break;

      // ---------- per-stage control pieces of the sweep, one lane per stage (they do not depend on the recursion):
      // GUD rows 2,3 <- cw, 6,7 <- ev (every pass); 0,1 <- Uq diagonal, 4,5 <- E diagonal (when the factors are built)
      GLANES_BEGIN(NT)
        FOR_MY_STAGES(i) if (i < NS) {
          const double iT = IT[i];
          const double u0 = U[i], u1 = U[d + i];
          const double dc0 = (u0 - (i ? U[i - 1] : uic[0])) * iT, dc1 = (u1 - (i ? U[d + i - 1] : uic[1])) * iT;
          const double ev0 = (2.0 * (P.Rd[0] * dc0 + P.Rd[1] * dc1) + GUD[6 * d + i]) * iT;
          const double ev1 = (2.0 * (P.Rd[1] * dc0 + P.Rd[2] * dc1) + GUD[7 * d + i]) * iT;
          const double cwc0 = 2.0 * (P.Rm[0] * u0 + P.Rm[1] * u1) + GUD[2 * d + i] + ev0;
          const double cwc1 = 2.0 * (P.Rm[1] * u0 + P.Rm[2] * u1) + GUD[3 * d + i] + ev1;
          GUD[6 * d + i] = ev0; GUD[7 * d + i] = ev1; GUD[2 * d + i] = cwc0; GUD[3 * d + i] = cwc1;
          if (pass == 0) {
            GUD[4 * d + i] = (2.0 * P.Rd[0] + GUD[4 * d + i]) * iT * iT; GUD[5 * d + i] = (2.0 * P.Rd[2] + GUD[5 * d + i]) * iT * iT;
            GUD[i] = 2.0 * P.Rm[0] + GUD[i]; GUD[d + i] = 2.0 * P.Rm[2] + GUD[d + i];
          }
        }
      GLANES_END(NW)
      // ---------- backward Riccati sweep
      double Pi1th = cth, Pithth = Dthth;
      for (int i = NS - 1; i >= 0; i--) {
        const double* A = ABG + 54 * i; const double* B = A + 36;
        double* fac = FAC + 20 * i; double* kf = KFF + 6 * i;
        // stage control pieces (prepared per stage before the sweep): E (rate Hessian), Uq, cw, czu
        const double ev0 = GUD[6 * d + i], ev1 = GUD[7 * d + i], cwc0 = GUD[2 * d + i], cwc1 = GUD[3 * d + i];
        if (pass == 0) {
          const double iT = IT[i];
          const double e0 = GUD[4 * d + i], e1 = 2.0 * P.Rd[1] * iT * iT, e2 = GUD[5 * d + i];
          const double uq0 = GUD[i], uq1 = 2.0 * P.Rm[1], uq2 = GUD[d + i];
          // phases a-c produce 8x8 blocks: output (r, c) with c = lane & 7 and r = (lane >> 3) + (NT / 8) * round.
          // The rounds are unrolled with r's range visible to the compiler, so the row tests of the first round fold.
          // phase a: [M_xx A | M_xx B + M_xu] (rows 0..5) and, in rows 6,7, A' l_x | B' l_x + l_u for the two rhs columns.
          // L1 / LTH follow PM in the layout (rows 8, 9 of the same array) and AXBW follows MAB (its rows 6, 7), so all
          // 64 outputs are the same expression: out[r][c] = (c >= 6 ? S[rr][c] : 0) + sum_k S[rr][k] [A|B][k][c], rr = r or r + 2
          GLANES_BEGIN(NT)
#pragma unroll
            for (int rd = 0; rd < LMPC_R8; rd++) {
              const int c = lane & 7, r = ((lane >> 3) & (LMPC_L8 - 1)) + LMPC_L8 * rd;
              if (NT > 64 && lane >= 64) break;
              const LmpcD2* col = reinterpret_cast<const LmpcD2*>(A + 6 * c);   // [A | B] is contiguous: column c of B follows A's six columns
              const double* prow = PM + 8 * (r + ((r >= 6) ? 2 : 0));
              const LmpcD2* pr2 = reinterpret_cast<const LmpcD2*>(prow);
              const double pc = prow[c];   // unconditional load + select: no branch around it
              double a = (c < 6) ? 0.0 : pc;
#pragma unroll
              for (int k = 0; k < 3; k++) { const LmpcD2 pv = pr2[k], cv = col[k]; a += pv.x * cv.x; a += pv.y * cv.y; }
              MAB[8 * r + c] = a;
            }
          GLANES_END(NW)
          // phase b: Yxx = A' MA, Yxu = A' MB, Yuu = B' MB + M_ux B + M_uu.  The first product is the same expression for
          // every (r, c) -- column r of [A|B] against column c of MAB -- the M_ux B + M_uu part is a four-lane tail.
          GLANES_BEGIN(NT)
#pragma unroll
            for (int rd = 0; rd < LMPC_R8; rd++) {
              const int c = lane & 7, r = ((lane >> 3) & (LMPC_L8 - 1)) + LMPC_L8 * rd;
              if (NT > 64 && lane >= 64) break;
              const LmpcD2* ar = reinterpret_cast<const LmpcD2*>(A + 6 * r);
              double a = 0.0;
#pragma unroll
              for (int k = 0; k < 3; k++) { const LmpcD2 av = ar[k]; a += av.x * MAB[8 * (2 * k) + c]; a += av.y * MAB[8 * (2 * k + 1) + c]; }
              if (r >= 6 && c >= 6) {   // the four Yuu entries: + M_uu + M_ux B (one divergent tail instead of one per term)
                double e = PM[8 * r + c];
#pragma unroll
                for (int k = 0; k < 6; k++) e += PM[8 * k + r] * B[k + 6 * (c - 6)];
                a += e;
              }
              // rows 6,7 of Qzw = -E are parked in the unused (r >= 6, c < 2) slots so that phase c reads
              // Qzw[r][j] = QZ(r, j) without selecting between Yxu and -E
              if (r >= 6 && c < 2) a = (r == 6) ? (c == 0 ? -e0 : -e1) : (c == 0 ? -e1 : -e2);
              YY[8 * r + c] = a;
            }
          GLANES_END(NW)
#ifdef LMPC_DEBUG_TRACE
          if (polishing) { LANE0_ONLY(if (LMPC_TRACE_COND) printf("   st %2d Pdiag %.3e %.3e %.3e %.3e %.3e %.3e %.3e %.3e | hx %.2e %.2e %.2e %.2e %.2e | Yuu %.3e %.3e uq %.2e %.2e e %.2e %.2e\n", i, PM[0], PM[9], PM[18], PM[27], PM[36], PM[45], PM[54], PM[63], HX[1*d+i], HX[2*d+i], HX[3*d+i], HX[4*d+i], HX[5*d+i], YY[8*6+6], YY[8*7+7], uq0, uq2, e0, e2);) }
#endif
          // phase c (uniform part): S = Yuu + E + Uq, its inverse, feed-forward terms
          const double q0_ = YY[8 * 6 + 6] + uq0, q1_ = 0.5 * (YY[8 * 6 + 7] + YY[8 * 7 + 6]) + uq1, q2_ = YY[8 * 7 + 7] + uq2;
          const double s0 = q0_ + e0, s1 = q1_ + e1, s2_ = q2_ + e2;
          const double det = s0 * s2_ - s1 * s1;
          if (!(s0 > 0.0) || !(det > 0.0)) {
#ifdef LMPC_DEBUG_TRACE
            LANE0_ONLY(if (LMPC_TRACE_COND) printf("  S not PD at stage %d: s %.6e %.6e %.6e det %.3e | q %.6e %.6e %.6e e %.6e %.6e %.6e uq %.3e %.3e Yuu %.6e %.6e %.6e\n", i, s0, s1, s2_, det, q0_, q1_, q2_, e0, e1, e2, uq0, uq2, YY[8*6+6], YY[8*6+7], YY[8*7+7]);)
#endif
            fail = true; break;
          }
          const double idet = lmpc_rcp(det);
          const double i0 = s2_ * idet, i1 = -s1 * idet, i2_ = s0 * idet;
          const double cw1_0 = cwc0 + AXBW[6], cw1_1 = cwc1 + AXBW[7];
          const double cwt_0 = AXBW[8 + 6], cwt_1 = AXBW[8 + 7];
          const double k1_0 = i0 * cw1_0 + i1 * cw1_1, k1_1 = i1 * cw1_0 + i2_ * cw1_1;
          const double kt_0 = i0 * cwt_0 + i1 * cwt_1, kt_1 = i1 * cwt_0 + i2_ * cwt_1;
          Pithth -= cwt_0 * kt_0 + cwt_1 * kt_1;
          Pi1th -= cwt_0 * k1_0 + cwt_1 * k1_1;
          // P_uu = Q Sinv E  (product form, no cancellation), symmetrised
          const double se00 = i0 * e0 + i1 * e1, se01 = i0 * e1 + i1 * e2, se10 = i1 * e0 + i2_ * e1, se11 = i1 * e1 + i2_ * e2;
          const double p00 = q0_ * se00 + q1_ * se10, p01 = q0_ * se01 + q1_ * se11;
          const double p10 = q1_ * se00 + q2_ * se10, p11 = q1_ * se01 + q2_ * se11;
          const double p01s = 0.5 * (p01 + p10);
          GLANES_BEGIN(NT)
            // P (8x8), then one more 4x8 block: rows 0,1 Kz, row 2 l1, row 3 lth; lane 0 stores the stage scalars
#pragma unroll
            for (int rd = 0; rd < LMPC_R8; rd++) {
              const int c = lane & 7, r = ((lane >> 3) & (LMPC_L8 - 1)) + LMPC_L8 * rd;
              if (NT > 64 && lane >= 64) break;
              // every load is unconditional (all addresses are valid), the cases are selects: no branches in this phase
              const double qr0 = YY[QZ0(r)], qr1 = YY[QZ0(r) + 1], qc0 = YY[QZ0(c)], qc1 = YY[QZ0(c) + 1];
              const double yv = YY[8 * r + c], hx = HX[(r < 6 ? r : 0) * d + i];
              const double qzz = (r < 6 && c < 6) ? yv + (r == c ? hx : 0.0) : 0.0;
              double pv = qzz - (qr0 * (i0 * qc0 + i1 * qc1) + qr1 * (i1 * qc0 + i2_ * qc1));
              if (r >= 6 && c >= 6) pv = (r == 6 && c == 6) ? p00 : ((r == 7 && c == 7) ? p11 : p01s);
              PM[8 * r + c] = pv;
            }
            if (lane < 32) {
              const int c = lane & 7, j = (lane >> 3) & 3;
              const double qc0 = YY[QZ0(c)], qc1 = YY[QZ0(c) + 1];
              const bool isth = (j == 3);
              // rows 0,1: Kz (2x8): Kz[j][c] = Sinv[j][:] . Qzw[c][:];  rows 2,3: l1 and lth:  l = Cz' - Qzw kff
              const double kz = (j == 0) ? (i0 * qc0 + i1 * qc1) : (i1 * qc0 + i2_ * qc1);
              const double kk0 = isth ? kt_0 : k1_0, kk1 = isth ? kt_1 : k1_1;
              const double czth = CZTH[i], czx = CZX[(c < 6 ? c : 0) * d + i], axv = AXBW[(isth ? 8 : 0) + c];
              const double cz = isth ? (c == 1 ? czth : 0.0) : (c < 6 ? czx : (c == 6 ? -ev0 : -ev1));
              const double ax = (c < 6) ? axv : 0.0;
              const double lv = cz + ax - (qc0 * kk0 + qc1 * kk1);
              double* dst = (j < 2) ? (fac + 8 * j + c) : ((isth ? LTH : L1) + c);
              *dst = (j < 2) ? kz : lv;
            }
            if (lane == 0) { fac[16] = i0; fac[17] = i1; fac[18] = i2_; kf[0] = k1_0; kf[1] = k1_1; kf[2] = kt_0; kf[3] = kt_1; kf[4] = cwt_0; kf[5] = cwt_1; }
          GLANES_END(NW)
        } else {
          // pass 1: right-hand side "1" column only (factors unchanged)
          const double i0 = fac[16], i1 = fac[17], i2_ = fac[18];
          double bw0 = L1[6], bw1 = L1[7];
#pragma unroll
          for (int k = 0; k < 6; k++) { bw0 += B[k] * L1[k]; bw1 += B[6 + k] * L1[k]; }
          const double cw0 = cwc0 + bw0, cw1 = cwc1 + bw1;
          const double k0 = i0 * cw0 + i1 * cw1, k1 = i1 * cw0 + i2_ * cw1;
          Pi1th -= kf[4] * k0 + kf[5] * k1;
          LaneVar<double, NT> newl;
          GLANES_BEGIN(NT)
            {   // every lane computes row (lane & 7) without branches; lanes 0..7 store in the next phase
              const int r = lane & 7, r6 = (r < 6) ? r : 0;
              const double czx = CZX[r6 * d + i];
              const double a0 = (r < 6) ? czx : ((r == 6) ? -ev0 : -ev1);
              double a2 = a0;
#pragma unroll
              for (int k = 0; k < 6; k++) a2 += A[k + 6 * r6] * L1[k];
              double a = (r < 6) ? a2 : a0;
              a -= fac[r] * cw0 + fac[8 + r] * cw1;
              newl(lane) = a;
            }
          GLANES_END(NW)
          GLANES_BEGIN(NT)
            if (lane < 8) L1[lane] = newl(lane);
            if (lane == 8) { kf[0] = k0; kf[1] = k1; }
          GLANES_END(NW)
        }
      }
      if (fail) break;
      if (pass == 0) Pithth_keep = Pithth; else Pithth = Pithth_keep;

      // ---------- sigma_b step, forward sweep into DXA/DUA (pass 0) or DXF/DUF (pass 1)
      double* DXo = pass ? DXF : DXA; double* DUo = pass ? DUF : DUA;
      const double dthp = soft ? -Pi1th / Pithth : 0.0;
      GLANES_BEGIN(NT)
        if (lane < 6) DXo[lane * d] = 0.0;
      GLANES_END(NW)
      for (int i = 0; i < NS; i++) {
        const double* A = ABG + 54 * i; const double* B = A + 36;
        const double* fac = FAC + 20 * i; const double* kf = KFF + 6 * i;
        double dz[8];
#pragma unroll
        for (int k = 0; k < 6; k++) dz[k] = DXo[k * d + i];
        const int im = i ? i - 1 : 1;                    // stage 0 has no predecessor: read an entry this phase does not write
        const double pu0 = DUo[im], pu1 = DUo[d + im];   // unconditional loads, then selects
        dz[6] = i ? pu0 : 0.0; dz[7] = i ? pu1 : 0.0;
        double du0 = -kf[0] - kf[2] * dthp, du1 = -kf[1] - kf[3] * dthp;
        const LmpcD2* f2 = reinterpret_cast<const LmpcD2*>(fac);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const LmpcD2 fa = f2[k], fb = f2[4 + k];
          du0 -= fa.x * dz[2 * k]; du1 -= fb.x * dz[2 * k];
          du0 -= fa.y * dz[2 * k + 1]; du1 -= fb.y * dz[2 * k + 1];
        }
        GLANES_BEGIN(NT)
          {
            const int l6 = (lane < 6) ? lane : 0;   // lanes >= 6 repeat row 0 and store nothing
            double a = B[l6] * du0 + B[6 + l6] * du1;
#pragma unroll
            for (int k = 0; k < 6; k++) a += A[l6 + 6 * k] * dz[k];
            if (lane < 6) DXo[lane * d + i + 1] = a;
            if (lane == 6) { DUo[i] = du0; DUo[d + i] = du1; }
          }
        GLANES_END(NW)
      }

      // ---------- terminal directions (lambda), sigma_b dual
      if (learn) {
        // e = pT + PT dx_N (6), q = q0 - Xq dx_N (NQ): one lane per entry, then every lane reads the 11 values
        GLANES_BEGIN(NT)
          if (lane < 6 + LMPC_NQ) {
            double dxn[6];
#pragma unroll
            for (int b = 0; b < 6; b++) dxn[b] = (b < nh) ? DXo[P.hidx[b] * d + N - 1] : 0.0;
            double s2;
            if (lane < 6) {
              s2 = TB[TB_PTV + lane];
#pragma unroll
              for (int b = 0; b < 6; b++) if (b < nh) s2 += TB[TB_PT + 6 * lane + b] * dxn[b];
              if (lane >= nh) s2 = 0.0;
            } else {
              const int q = lane - 6;
              s2 = TB[TB_Q0 + q];
#pragma unroll
              for (int a = 0; a < 6; a++) if (a < nh) s2 -= TB[TB_XQ + q * 6 + a] * dxn[a];
            }
            TB[TB_EQ + lane] = s2;
          }
        GLANES_END(NW)
        double e[6], qv[LMPC_NQ];
#pragma unroll
        for (int a = 0; a < 6; a++) e[a] = TB[TB_EQ + a];
#pragma unroll
        for (int q = 0; q < LMPC_NQ; q++) qv[q] = TB[TB_EQ + 6 + q];
        const double nu = qv[0];
        GLANES_BEGIN(NT)
#pragma unroll
          for (int p = 0; p < KPL; p++) {
            const int k = lane + NT * p;
            if (k < K) {
              const double l = lam(lane).a[p], y = ylam(lane).a[p];
              double tl = smu;
              if (pass) tl -= csc * dla(lane).a[p] * dya(lane).a[p];
              const double il_ = lmpc_rcp(l);
              tl *= il_;
              // both forms are evaluated (the loads of the three columns go out together), the column's kind selects
              const int ib = isB(lane).a[p];
              double dlb = qv[1];
#pragma unroll
              for (int q = 2; q < LMPC_NQ; q++) dlb = (ib == q) ? qv[q] : dlb;
              double se = 0.0;
#pragma unroll
              for (int a = 0; a < 6; a++) se += ST[6 * k + a] * e[a];
              const double dln = omg_(lane).a[p] * (se - glam(lane).a[p] - nu);
              const double dl = ib ? dlb : dln;
              const double dy = tl - y - dl * (y * il_);
              if (pass) { dlf(lane).a[p] = dl; dyf(lane).a[p] = dy; } else { dla(lane).a[p] = dl; dya(lane).a[p] = dy; }
            }
          }
        GLANES_END(NW)
      }
      double dthc = dthp, dythc = 0.0;
      if (polishing) { dtha = dthc; break; }   // the polish takes the full step below, no ratio test
      if (soft && !LMPC_FREE_THETA) { const double ith = lmpc_rcp(th); const double tt = (smu - corr_th) * ith; dythc = tt - yth - (yth * ith) * dthc; }
      if (pass) { dth = dthc; dyth = dythc; } else { dtha = dthc; dytha = dythc; }

      // ---------- row directions, step length
      LaneVar<double, NT> ra[2];
      GLANES_BEGIN(NT)
        double rmax = 0.0, cross = 0.0;   // rmax = max over rows of (-ds/s, -dy/y)  ->  amax = 1 / rmax
        const double thq = soft ? th : 0.0, dthaq = soft ? dtha : 0.0, dthq = soft ? dth : 0.0;
#define ROW_BODY                                                           \
  {                                                                        \
    const double s = RSs[slot * d + i], y = RSy[slot * d + i], isy = RSi[slot * d + i]; \
    const double is = y * isy, iy = s * isy;                               \
    const double rp = (sg * v - (isb ? thq : 0.0)) + s - bnd;              \
    const double dsa = -rp - (sg * va - (isb ? dthaq : 0.0));              \
    const double dya_ = -y - y * is * dsa;                                 \
    const double dsf = -rp - (sg * vf - (isb ? dthq : 0.0));               \
    const double dyf_ = (-(s * y - smu + csc * dsa * dya_) - y * dsf) * is; \
    const double ds = pass ? dsf : dsa, dy = pass ? dyf_ : dya_;           \
    cross += (pass || !on) ? 0.0 : dsa * dya_;                             \
    rmax = on ? lmpc_max(lmpc_max(rmax, -ds * is), -dy * iy) : rmax;       \
  }
        if constexpr (NW == 1) LMPC_FOR_ROWS_BY_LANE(true, pass != 0)
        else { FOR_MY_STAGES(i) LMPC_FOR_ROWS(i, true, pass != 0) }
#undef ROW_BODY
#pragma unroll
        for (int p = 0; p < KPL; p++) {
          const int k = lane + NT * p;
          if (k < K) {
            const double dl = pass ? dlf(lane).a[p] : dla(lane).a[p], dy = pass ? dyf(lane).a[p] : dya(lane).a[p];
            rmax = lmpc_max(lmpc_max(rmax, -dl * lmpc_rcp(lam(lane).a[p])), -dy * lmpc_rcp(ylam(lane).a[p]));
            if (!pass) cross += dl * dy;
          }
        }
        ra[0](lane) = rmax; ra[1](lane) = cross;
      GLANES_END(NW)
      {
        const int ops[2] = {LMPC_RED_MAX, LMPC_RED_SUM};
        group_reduce<NW, 2>(ra, ops, RED);
      }
      double rmax = ra[0](0), cross = ra[1](0);
      if (soft && !LMPC_FREE_THETA) {
        rmax = fmax(rmax, fmax(-dthc / th, -dythc / yth));
        if (!pass) cross += dthc * dythc;
      }
      const double amax = rmax > 0.0 ? 1.0 / rmax : 1e300;
      if (pass == 0) {
        // mu_aff: sum (s + a ds)(y + a dy) = (1 - a) sum s y + a^2 sum ds dy   (affine step: s dy + y ds = -s y)
        const double aa = amax < 1.0 ? amax : 1.0;
        const double mua = (1.0 - aa) * mu + aa * aa * cross * inv_m;
        const double rt_ = mua / mu;
        sigma = rt_ * rt_ * rt_;
        csc = (aa < 0.2) ? aa : 1.0;   // Mehrotra's second-order term is harmful when the affine step is short
#ifdef LMPC_DEBUG_TRACE
        LANE0_ONLY(if (LMPC_TRACE_COND) printf("  aff it %d: rmax_rows %.9e cross_rows %.9e aa %.6f dtha %.9e dytha %.9e Pi1th %.9e Pithth %.9e mua %.6e DXA[3,N-1] %.9e DUA0 %.9e dla0 %.9e\n", it, ra[0](0), ra[1](0), aa, dtha, dytha, Pi1th, Pithth, mua, DXA[3 * d + N - 1], DUA[0], dla(0).a[0]);)
#endif
      } else {
        const double tau = 1.0 - fmin(0.005, mu);
        alpha = tau * amax; if (alpha > 1.0) alpha = 1.0;
      }
    }  // pass
    if (restart) { it--; continue; }
    bool polish_failed = false;
    if (polishing && fail) polish_failed = true;
    else if (fail) {
      // numerical floor of the barrier-weighted recursion: the iterate is intact, let the polish finish from it
      if (!numfail_polish && mu < 1e-5) { numfail_polish = 1; polishing = 1; classified = 0; polish_tries = 0; continue; }
      status = LMPC_NUMERIC;
      break;
    }
    if (polishing) n_polish_rounds++;
    if (polishing && !polish_failed) {
      // ---------- full Newton step of the augmented-Lagrangian model, multiplier update, active-set refinement
      const double ftol = 1e-10, dtol = 1e-9;
      GLANES_BEGIN(NT)
        FOR_MY_STAGES(i) {
          if (i >= 1) for (int c = 0; c < 6; c++) X[c * d + i] += DXA[c * d + i];
          if (i < NS) for (int c = 0; c < 2; c++) U[c * d + i] += DUA[c * d + i];
        }
      GLANES_END(NW)
      double changed = 0.0, dymax_u = 0.0;
      if (soft) {
        th += dtha;
        if (pact_th) { dymax_u = fabs(LMPC_PRHO * th) / (1.0 + fabs(yth)); yth += LMPC_PRHO * (-th); if (yth < -dtol) { pact_th = 0; yth = 0.0; changed += 1.0; } }
        else if (!LMPC_FREE_THETA && th < -ftol) { pact_th = 1; changed += 1.0; }
      }
      LaneVar<double, NT> rc2[3];
      GLANES_BEGIN(NT)
        double ch = 0.0, viol = 0.0, dym = 0.0;
        const double thq = soft ? th : 0.0;
#define ROW_BODY                                                           \
  {                                                                        \
    const double s = RSs[slot * d + i];                                    \
    const double r = (sg * v - (isb ? thq : 0.0)) - bnd;                   \
    bool act = s < 0.0;                                                    \
    if (act) {                                                             \
      const double yo = RSy[slot * d + i];                                 \
      const double yn = yo + LMPC_PRHO * r;                                \
      dym = fmax(dym, fabs(LMPC_PRHO * r) / (1.0 + fabs(yo)));             \
      if (yn < -dtol * fmax(1.0, fabs(yn))) { act = false; RSs[slot * d + i] = -s; RSy[slot * d + i] = 0.0; ch += 1.0; } \
      else RSy[slot * d + i] = yn;                                         \
    } else if (r > ftol) { act = true; RSs[slot * d + i] = -s; ch += 1.0; } \
    if (!act && r > 1e-9) viol += 1.0;                                     \
  }
        FOR_MY_STAGES(i) LMPC_FOR_ROWS(i, false, false)
#undef ROW_BODY
        for (int p = 0; p < KPL; p++) {
          const int k = lane + NT * p;
          if (k < K) {
            const double l = lam(lane).a[p] + dla(lane).a[p];
            lam(lane).a[p] = l;
            if (pnb(lane).a[p]) { dym = fmax(dym, fabs(LMPC_PRHO * l) / (1.0 + fabs(ylam(lane).a[p]))); const double yn = ylam(lane).a[p] + LMPC_PRHO * (-l); if (yn < -dtol) { pnb(lane).a[p] = 0; ylam(lane).a[p] = 0.0; ch += 1.0; } else ylam(lane).a[p] = yn; }
            else if (l < -ftol) { pnb(lane).a[p] = 1; ch += 1.0; }
          }
        }
        rc2[0](lane) = ch; rc2[1](lane) = viol; rc2[2](lane) = dym;
      GLANES_END(NW)
      {
        const int ops[3] = {LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_MAX};
        group_reduce<NW, 3>(rc2, ops, RED);
      }
      changed += rc2[0](0);
      const double dymax = fmax(dymax_u, rc2[2](0));   // a second AL step removes the bias left by inexact multipliers
#ifdef LMPC_DEBUG_TRACE
      LANE0_ONLY(if (LMPC_TRACE_COND) printf("  polish round %d: changed %.0f viol %.0f th %.3e pact_th %d dymax %.3e (u %.3e)\n", polishing, changed, rc2[1](0), th, pact_th, dymax, dymax_u);)
#endif
      // a polish whose active set is coming apart (more rows change side than in the round before, and more than a
      // handful) will not settle: give up now instead of spending the remaining rounds (one such instance per ~1000 set the
      // time of a one-wave batch: 6 wasted rounds = 3.3 iteration-equivalents)
      const bool diverging = polishing >= 2 && changed > 8.5 && changed > 2.0 * prev_changed;
      prev_changed = changed;
      // stop after the FIRST round only when it moved the multipliers by less than LMPC_PDY (its feasibility error is that
      // change / rho); from the second round on the bar is LMPC_PDY2, the change a second augmented-Lagrangian step leaves
      if (!diverging && (changed > 0.5 || dymax > (polishing >= 2 ? LMPC_PDY2 : LMPC_PDY)) && polishing < LMPC_PMAX) { polishing++; continue; }
      if (!diverging && changed < 0.5 && rc2[1](0) < 0.5) { status = LMPC_SOLVED; it++; break; }
      polish_failed = true;
    }
    if (polish_failed) {
#ifdef LMPC_DEBUG_TRACE
      LANE0_ONLY(if (LMPC_TRACE_COND) printf("  polish FAILED (fail=%d) tries %d it %d\n", (int)fail, polish_tries, it);)
#endif
      // no consistent active set: restore the interior-point iterate; the first time keep iterating with a
      // 100x tighter tolerance and try once more, the second time return the interior-point solution
      GLANES_BEGIN(NT)
#define ROW_BODY                                                           \
  {                                                                        \
    const double s = fabs(RSs[slot * d + i]), y = RSi[slot * d + i];       \
    RSs[slot * d + i] = s; RSy[slot * d + i] = y; RSi[slot * d + i] = 1.0 / (s * y); \
  }
        FOR_MY_STAGES(i) LMPC_FOR_ROWS(i, false, false)
#undef ROW_BODY
        FOR_MY_STAGES(i) { for (int c = 0; c < 6; c++) X[c * d + i] = SAVE_XU[8 * i + c]; if (i < NS) for (int c = 0; c < 2; c++) U[c * d + i] = SAVE_XU[8 * i + 6 + c]; }
        for (int p = 0; p < KPL; p++) if (lane + NT * p < K) { lam(lane).a[p] = SAVE_L[lane + NT * p]; ylam(lane).a[p] = SAVE_L[K + lane + NT * p]; }
      GLANES_END(NW)
      th = th_save; yth = yth_save;
      polishing = 0; classified = 0;
      if (numfail_polish) { status = LMPC_NUMERIC; it++; break; }
      if (polish_tries >= 2 || it >= P.max_iter) { status = LMPC_SOLVED_INACCURATE; it++; break; }
      tol_step *= 1e-2; tol_mu *= 1e-2;
      continue;
    }
    // ---------- update the iterate: rows first (they read X/U of neighbouring stages), then the primal
    {
      const double smu = sigma * mu;
      LaneVar<double, NT> rstep;
      GLANES_BEGIN(NT)
        const double thq = soft ? th : 0.0, dthaq = soft ? dtha : 0.0, dthq = soft ? dth : 0.0;
#define ROW_BODY                                                           \
  {                                                                        \
    const double s = RSs[slot * d + i], y = RSy[slot * d + i], isy = RSi[slot * d + i]; \
    const double is = y * isy;                                             \
    const double rp = (sg * v - (isb ? thq : 0.0)) + s - bnd;              \
    const double dsa = -rp - (sg * va - (isb ? dthaq : 0.0));              \
    const double dya_ = -y - y * is * dsa;                                 \
    const double ds = -rp - (sg * vf - (isb ? dthq : 0.0));                \
    const double dy = (-(s * y - smu + csc * dsa * dya_) - y * ds) * is;   \
    const double sn = s + alpha * ds, yn = y + alpha * dy;                 \
    if (on) { RSs[slot * d + i] = sn; RSy[slot * d + i] = yn; RSi[slot * d + i] = lmpc_rcp(sn * yn); } \
  }
        if constexpr (NW == 1) LMPC_FOR_ROWS_BY_LANE(true, true)
        else { FOR_MY_STAGES(i) LMPC_FOR_ROWS(i, true, true) }
#undef ROW_BODY
#pragma unroll
        for (int p = 0; p < KPL; p++) {
          const int k = lane + NT * p;
          if (k < K) { lam(lane).a[p] += alpha * dlf(lane).a[p]; ylam(lane).a[p] += alpha * dyf(lane).a[p]; }
        }
        // step-based acceptance: the primal step per channel, relative to max(1, |channel|)
        double m = 0.0;
        for (int i = lane; i < N; i += NT) {
          for (int c = 0; c < 6; c++) m = lmpc_max(m, fabs(DXF[c * d + i]) * chs_[c]);
          if (i < NS) for (int c = 0; c < 2; c++) {
            const double dup = i ? DUF[c * d + i - 1] : 0.0;
            m = lmpc_max(m, fabs(DUF[c * d + i]) * chs_[6 + c]);
            m = lmpc_max(m, fabs(DUF[c * d + i] - dup) * IT[i] * chs_[8 + c]);
          }
        }
        rstep(lane) = m;
      GLANES_END(NW)
      GLANES_BEGIN(NT)
        for (int i = lane; i < N; i += NT) {
          if (i >= 1) for (int c = 0; c < 6; c++) X[c * d + i] += alpha * DXF[c * d + i];
          if (i < NS) for (int c = 0; c < 2; c++) U[c * d + i] += alpha * DUF[c * d + i];
        }
      GLANES_END(NW)
      if (soft) { th += alpha * dth; yth += alpha * dyth; }
      rho_d *= (1.0 - alpha);
      group_max<NW>(rstep, RED);
      const double stepn = alpha * rstep(0);
      const double ratio = (prev_stepn > 0.0) ? stepn / prev_stepn : 1.0;
      const double est = (ratio < 0.9) ? stepn * ratio / (1.0 - ratio) : 1e300;   // geometric-tail estimate
      prev_stepn = stepn;
#ifdef LMPC_DEBUG_TRACE
      LANE0_ONLY(if (LMPC_TRACE_COND) printf("it %2d mu %.6e rp %.3e sigma %.6e alpha %.6f th %.6e stepn %.3e csc %.3f\n", it, mu, rpn, sigma, alpha, th, stepn, csc);)
#endif
      if (stepn < tol_step && est < tol_step && alpha > 0.5 && mu < 1e-6 && rpn < 1e-9 && fabs(rnu) < 1e-9) { polishing = 1; classified = 0; }
    }
  }  // iterations

  // ---------------------------------------------------------------- outputs
  for (int i = 0; i < NS; i++) {   // consistent rollout of the linear dynamics
    GLANES_BEGIN(NT)
      if (lane < 6) {
        const double* A = ABG + 54 * i; const double* B = A + 36; const double* g = A + 48;
        double a = g[lane];
        for (int k = 0; k < 6; k++) a += A[lane + 6 * k] * X[k * d + i];
        for (int k = 0; k < 2; k++) a += B[lane + 6 * k] * U[k * d + i];
        X[lane * d + i + 1] = a;
      }
    GLANES_END(NW)
  }
  LaneVar<double, NT> ro[7];
  GLANES_BEGIN(NT)
    double cst = 0.0;
    for (int i = lane; i < N; i += NT) {
      for (int c = 0; c < 6; c++) out.X[6 * i + c] = X[c * d + i];
      if (i < NS) {
        const double u0 = U[i], u1 = U[d + i];
        const double iT = IT[i];
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
      const int k = lane + NT * p;
      if (k < K) {
        const double l = lam(lane).a[p];
        cst += sscv(lane).a[p] * l;
        for (int a = 0; a < 6; a++) sg[a] += ST[6 * k + a] * l;
      }
      if (out.lam && k < P.K) out.lam[k] = (k < K) ? fmax(lam(lane).a[p], 0.0) : 0.0;   // polished non-basic columns sit at -y/rho ~ 1e-20
    }
    ro[0](lane) = cst;
    for (int a = 0; a < 6; a++) ro[1 + a](lane) = sg[a];
  GLANES_END(NW)
  {
    const int ops[7] = {LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM, LMPC_RED_SUM};
    group_reduce<NW, 7>(ro, ops, RED);
  }
  double cost = ro[0](0);
  if (soft) cost += P.qb * th * th;
  if (learn && P.hull_slack)
    for (int a = 0; a < nh; a++) { const int c = P.hidx[a]; const double sh = X[c * d + N - 1] - in.cen[c] - ro[1 + a](0); cost += P.chs[c] * sh * sh; }
  LANE0_ONLY(if (out.cost) *out.cost = cost; *out.status = status; *out.iters = it;
             if (out.stats) { out.stats[0] = it - n_polish_rounds; out.stats[1] = n_polish_rounds; out.stats[2] = polish_tries; out.stats[3] = 0; })
#undef LO
#undef ROW_GBEGIN
#undef ROW_GEND
}
