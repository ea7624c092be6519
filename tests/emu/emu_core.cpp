// Test-only lane-loop emulator of the product's warp kernels (LMPC_EMULATE build of the same
// .cuh sources).  Lets the kernel logic be checked against the CPU oracle without a GPU, and run
// with reversed lane order to expose phase-ordering (race) mistakes.  Never linked into the product.
#define LMPC_EMULATE 1
#include <stdlib.h>
#include <vector>
#include "../../racing-lmpc-ros2_b200/csrc/lmpc_host_params.h"
#include "../../racing-lmpc-ros2_b200/csrc/lmpc_ss_core.cuh"
#include "../../racing-lmpc-ros2_b200/csrc/lmpc_reg_core.cuh"
int g_lmpc_emu_reverse = 0;

extern "C" void emu_set_reverse(int r) { g_lmpc_emu_reverse = r; }
static int g_emu_stats[4];   // diagnostics of the last emu_qp_solve: interior-point iterations, polish rounds, polish attempts
extern "C" const int* emu_last_stats() { return g_emu_stats; }

extern "C" void emu_linearise(const lmpc_vehicle_params* v, const double* x, const double* u, double kappa, double dt,
                              double* A, double* B, double* g, double* xn) {
  LmpcModel P = lmpc_make_model(*v);
  lmpc_linearise(P, x, u, kappa, dt, A, B, g, xn);
}

// the linearisation kernel's form (tangent sums accumulated in the output, strided): out [54] = A | B | g
extern "C" void emu_linearise_staged(const lmpc_vehicle_params* v, const double* x, const double* u, double kappa, double dt,
                                     double* out, int stride) {
  LmpcModel P = lmpc_make_model(*v);
  lmpc_linearise_staged(P, x, u, kappa, dt, out, stride);
}

// one lap at a time: (query, lap) exactly as the kernel does it
extern "C" void emu_ss_query_lap(int m, const double* ps, const double* pe, const double* xr, const double* J, const int* canon,
                                 int take, int out_off, double qs, double qe, int max_total, double* ss_x, double* ss_j,
                                 int last, int count, int pad_to) {
  LmpcLapView lap = {ps, pe, xr, J, canon, m, take, out_off};
  lmpc_ss_query_warp(lap, qs, qe, max_total, ss_x, ss_j, last != 0, count, pad_to);
}

// full QP solve of one instance given its linearisation and safe-set columns; nw = warps per instance
extern "C" int emu_qp_solve(const lmpc_mpc_config* c, const lmpc_vehicle_params* v, const double* x_ic, const double* u_ic,
                            const double* U0, const double* T, const double* bl, const double* br, const double* vref,
                            const double* ABg, const double* ssx, const double* ssj_raw, const double* cen, int ss_count,
                            double* X, double* U, double* dU, double* lam, double* cost, int* status, int* iters,
                            int* smem_doubles, int nw) {
  LmpcQpParams P;
  int rc = lmpc_make_qp_params(*c, *v, &P, nw);
  if (rc != LMPC_OK) return rc;
  if (smem_doubles) *smem_doubles = P.lay.total;
  std::vector<double> sm((size_t)P.lay.total, 0.0 / 0.0);   // NaN-filled: reads of unwritten scratch show up
  std::vector<double> scr((size_t)LMPC_QP_SCRATCH(P.N, P.K > 0 ? P.K : 1), 0.0 / 0.0);
  LmpcQpIn in = {x_ic, u_ic, U0, T, bl, br, vref, ABg, ssx, ssj_raw, cen, ss_count, scr.data()};
  LmpcQpOut out = {X, U, dU, lam, cost, status, iters, g_emu_stats};
  const int kpl = (P.K + 32 * nw - 1) / (32 * nw);
  const bool fixed = (nw & 0x100) == 0 && P.RS == 16 && (P.N == 20 || P.N == 40);   // same rule as the C ABI
  if (nw == 1) {
    if (fixed && P.N == 20 && kpl == 3) lmpc_qp_solve<1, 3, 20, 16>(P, in, sm.data(), out);
    else if (fixed && P.N == 40 && kpl <= 1) lmpc_qp_solve<1, 1, 40, 16>(P, in, sm.data(), out);
    else if (fixed && P.N == 20 && kpl <= 1) lmpc_qp_solve<1, 1, 20, 16>(P, in, sm.data(), out);
    else if (kpl <= 1) lmpc_qp_solve<1, 1, 0, 0>(P, in, sm.data(), out);
    else if (kpl == 2) lmpc_qp_solve<1, 2, 0, 0>(P, in, sm.data(), out);
    else if (kpl == 3) lmpc_qp_solve<1, 3, 0, 0>(P, in, sm.data(), out);
    else lmpc_qp_solve<1, 4, 0, 0>(P, in, sm.data(), out);
  } else if (nw == 2) {
    if (kpl <= 1) lmpc_qp_solve<2, 1, 0, 0>(P, in, sm.data(), out); else lmpc_qp_solve<2, 2, 0, 0>(P, in, sm.data(), out);
  } else {
    lmpc_qp_solve<4, 1, 0, 0>(P, in, sm.data(), out);
  }
  return LMPC_OK;
}

// ---- track interpolation functions (host builder + the device evaluation code, compiled for the CPU)
#include "../../racing-lmpc-ros2_b200/csrc/lmpc_track.cuh"
static LmpcTrackHost g_trk;
extern "C" int emu_track_build(int n, int ncols, const double* table) { return lmpc_track_build(n, ncols, table, &g_trk) ? 0 : -1; }
extern "C" int emu_track_m() { return g_trk.m; }
extern "C" double emu_track_length() { return g_trk.L; }
extern "C" void emu_track_eval(int n, const double* s, double* out) {   // out [n][7]: left right curvature vel x y yaw
  const LmpcTrack T = g_trk.view();
  for (int i = 0; i < n; i++) {
    LmpcTrackPoint p;
    lmpc_track_eval(T, s[i], &p);
    double* o = out + 7 * (size_t)i;
    o[0] = p.left; o[1] = p.right; o[2] = p.curvature; o[3] = p.vel; o[4] = p.x; o[5] = p.y; o[6] = p.yaw;
  }
}
extern "C" void emu_track_f2g(int n, const double* f, double* g) { const LmpcTrack T = g_trk.view(); for (int i = 0; i < n; i++) lmpc_frenet_to_global(T, f + 3 * i, g + 3 * i); }
extern "C" void emu_track_g2f(int n, const double* g, double* f) { const LmpcTrack T = g_trk.view(); for (int i = 0; i < n; i++) lmpc_global_to_frenet(T, g + 3 * i, f + 3 * i); }

// error-dynamics regression of one query item over M prepared points (Z [M][8], E [M][6]); A, B column-major, in place
// the plan the host builds from a spec: which regressions the tiled kernel scans as a pair (lead / follower per row)
extern "C" int emu_reg_plan_pairs(const lmpc_reg_spec* sp, int* lead, int* follower) {
  LmpcRegPlan plan;
  if (!lmpc_make_reg_plan(sp, &plan)) return -1;
  for (int r = 0; r < plan.n_out; r++) { lead[r] = plan.row[r].lead; follower[r] = plan.row[r].follower; }
  return plan.n_out;
}

extern "C" int emu_regress(const lmpc_reg_spec* sp, int M, const double* Z, const double* E, const double* zq, double* A, double* B,
                           double* C, int* npts) {
  LmpcRegPlan plan;
  if (!lmpc_make_reg_plan(sp, &plan)) return -1;
  std::vector<double> Zc((size_t)8 * M), Ec((size_t)6 * M);   // the device keeps the samples by column
  for (int p = 0; p < M; p++) { for (int c = 0; c < 8; c++) Zc[(size_t)c * M + p] = Z[8 * (size_t)p + c]; for (int c = 0; c < 6; c++) Ec[(size_t)c * M + p] = E[6 * (size_t)p + c]; }
  LmpcRegView v = {Zc.data(), Ec.data(), M, M, -1};
  lmpc_regress_warp(plan, v, zq, A, B, C, npts);
  return 0;
}
