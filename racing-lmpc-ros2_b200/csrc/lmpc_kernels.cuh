// lmpc_kernels.cuh -- __global__ entry points of the batched LMPC solve (sm_100a).
//
//   K1 lmpc_linearise_kernel   thread per (instance, stage): abscissa alignment + RK4 Jacobians -> A,B,g
//   K2 lmpc_ss_query_kernel    warp per (instance, lap): exact k-NN in the safe-set slab + cost-to-go gather
//   K3 lmpc_qp_kernel          warp group (1, 2 or 4 warps = one CTA) per instance: interior-point / Riccati solve in shared memory
//
// Batch arrays are instance-major, so a warp's (or thread's) reads of its own instance are contiguous.
#pragma once
#include <cuda_runtime.h>
#include "lmpc_model.cuh"
#include "lmpc_qp_core.cuh"
#include "lmpc_ss_core.cuh"

#define LMPC_MAX_LAPS_USED 64   // laps one query can draw from (the table travels as a kernel parameter, 3.6 KB)

struct LmpcLapTable {
  LmpcLapView lap[LMPC_MAX_LAPS_USED];
  int n_used;     // laps that contribute columns
  int count;      // columns found = min(sum take, max_total)
};

// ---- K1: generic linearisation of n independent items (the C-ABI's lmpc_linearise_batch)
__global__ void lmpc_linearise_items_kernel(LmpcModel M, int n, const double* __restrict__ x, const double* __restrict__ u,
                                            const double* __restrict__ kappa, const double* __restrict__ dt,
                                            double* __restrict__ A, double* __restrict__ Bm, double* __restrict__ g,
                                            double* __restrict__ xnext) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double xl[6], ul[2], Al[36], Bl[12], gl[6], xn[6];
  for (int k = 0; k < 6; k++) xl[k] = x[6 * (size_t)t + k];
  ul[0] = u[2 * (size_t)t]; ul[1] = u[2 * (size_t)t + 1];
  lmpc_linearise(M, xl, ul, kappa[t], dt[t], Al, Bl, gl, xn);
  for (int k = 0; k < 36; k++) A[36 * (size_t)t + k] = Al[k];
  for (int k = 0; k < 12; k++) Bm[12 * (size_t)t + k] = Bl[k];
  for (int k = 0; k < 6; k++) g[6 * (size_t)t + k] = gl[k];
  if (xnext) for (int k = 0; k < 6; k++) xnext[6 * (size_t)t + k] = xn[k];
}

__global__ void lmpc_step_items_kernel(LmpcModel M, int n, const double* __restrict__ x, const double* __restrict__ u,
                                       const double* __restrict__ kappa, const double* __restrict__ dt,
                                       double* __restrict__ xnext) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double xl[6], ul[2], xn[6];
  for (int k = 0; k < 6; k++) xl[k] = x[6 * (size_t)t + k];
  ul[0] = u[2 * (size_t)t]; ul[1] = u[2 * (size_t)t + 1];
  lmpc_step(M, xl, ul, kappa[t], dt[t], xn);
  for (int k = 0; k < 6; k++) xnext[6 * (size_t)t + k] = xn[k];
}

// ---- K1 (solve path): per (instance b, stage i): align X_ref abscissa to x_ic (racing_mpc.cpp:219-223),
// linearise at (X_ref_i, U_ref_i, kappa_i, T_i) (racing_mpc.cpp:169-176), write [A|B|g] (54 doubles).
// Stage 0's thread also writes the aligned query / centre point X_ref[:, N-1].
__global__ void lmpc_linearise_kernel(LmpcModel M, int B, int N, const double* __restrict__ x_ic,
                                      const double* __restrict__ X_ref, const double* __restrict__ U_ref,
                                      const double* __restrict__ T_ref, const double* __restrict__ kappa,
                                      const double* __restrict__ total_length, double* __restrict__ ABg,
                                      double* __restrict__ cen) {
  const int NS = N - 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * NS) return;
  const int b = t / NS, i = t - b * NS;
  const double L = total_length[b], s0 = x_ic[6 * (size_t)b];
  const double* xr = X_ref + (6 * (size_t)N) * b + 6 * i;
  double xl[6], ul[2], Al[36], Bl[12], gl[6];
  for (int k = 0; k < 6; k++) xl[k] = xr[k];
  xl[0] = lmpc_align_abscissa(xl[0], s0, L);
  ul[0] = U_ref[(2 * (size_t)NS) * b + 2 * i]; ul[1] = U_ref[(2 * (size_t)NS) * b + 2 * i + 1];
  lmpc_linearise(M, xl, ul, kappa[(size_t)N * b + i], T_ref[(size_t)NS * b + i], Al, Bl, gl, nullptr);
  double* o = ABg + (54 * (size_t)NS) * b + 54 * i;
  for (int k = 0; k < 36; k++) o[k] = Al[k];
  for (int k = 0; k < 12; k++) o[36 + k] = Bl[k];
  for (int k = 0; k < 6; k++) o[48 + k] = gl[k];
  if (i == 0) {
    const double* xe = X_ref + (6 * (size_t)N) * b + 6 * (N - 1);
    double* c = cen + 6 * (size_t)b;
    c[0] = lmpc_align_abscissa(xe[0], s0, L);
    for (int k = 1; k < 6; k++) c[k] = xe[k];
  }
}

// ---- K2: one warp per (query b, lap slot j).  query is [B][qstride] with (s, e_y) in its first two entries.
__global__ void lmpc_ss_query_kernel(LmpcLapTable tab, int B, const double* __restrict__ query, int qstride,
                                     int max_total, int pad_to, double* __restrict__ ss_x, double* __restrict__ ss_j) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= B * tab.n_used) return;
  const int b = w / tab.n_used, j = w - b * tab.n_used;
  const double qs = query[(size_t)qstride * b], qe = query[(size_t)qstride * b + 1];
  lmpc_ss_query_warp(tab.lap[j], qs, qe, max_total, ss_x + (6 * (size_t)pad_to) * b, ss_j + (size_t)pad_to * b,
                     j == tab.n_used - 1, tab.count, pad_to);
}

// ---- K3: one CTA of NW warps per instance
struct LmpcQpBatch {
  const double *x_ic, *u_ic, *U0, *T_ref, *bl, *br, *vref, *ABg, *ssx, *ssj, *cen;
  double *X, *U, *dU, *lam, *cost;
  int *status, *iters;
  int ss_count;
  int B;
};

template <int NW, int KPL, int NTPL, int RSTPL>
__global__ void __launch_bounds__(32 * NW, 7) lmpc_qp_kernel(const __grid_constant__ LmpcQpParams P, const __grid_constant__ LmpcQpBatch a) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x;
  if (b >= a.B) return;
  const int N = P.N, NS = P.NS, K = P.K;
  LmpcQpIn in;
  in.x_ic = a.x_ic + 6 * (size_t)b; in.u_ic = a.u_ic + 2 * (size_t)b;
  in.U0 = a.U0 + (2 * (size_t)NS) * b; in.T = a.T_ref + (size_t)NS * b;
  in.bl = a.bl + (size_t)N * b; in.br = a.br + (size_t)N * b; in.vref = a.vref + (size_t)N * b;
  in.ABg = a.ABg + (54 * (size_t)NS) * b;
  in.ssx = P.learning ? a.ssx + (6 * (size_t)K) * b : nullptr;
  in.ssj = P.learning ? a.ssj + (size_t)K * b : nullptr;
  in.cen = a.cen + 6 * (size_t)b; in.ss_count = a.ss_count;
  LmpcQpOut out;
  out.X = a.X + (6 * (size_t)N) * b; out.U = a.U + (2 * (size_t)NS) * b; out.dU = a.dU + (2 * (size_t)NS) * b;
  out.lam = (a.lam && P.learning) ? a.lam + (size_t)K * b : nullptr;
  out.cost = a.cost ? a.cost + b : nullptr;
  out.status = a.status + b; out.iters = a.iters + b;
  lmpc_qp_solve<NW, KPL, NTPL, RSTPL>(P, in, sm, out);
}
