// lmpc_qp_kernel.cuh -- K3, the __global__ entry point of the QP solve (template; instantiated in lmpc_qp_tu*.cu).
#pragma once
#include <cuda_runtime.h>
#include "lmpc_qp_core.cuh"

// ---- K3: one CTA of NW warps per instance
struct LmpcQpBatch {
  const double *x_ic, *u_ic, *U0, *T_ref, *bl, *br, *vref, *ABg, *ssx, *ssj, *cen;
  double *X, *U, *dU, *lam, *cost;
  double* scratch;   // [B][LMPC_QP_SCRATCH(N, K)] workspace
  int *status, *iters;
  int ss_count;
  int B;
  const int* skip;   // optional [B]: non-zero = leave this instance's outputs untouched (converged SQP instances)
};

template <int NW, int KPL, int NTPL, int RSTPL>
__global__ void __launch_bounds__(32 * NW, 7) lmpc_qp_kernel(const __grid_constant__ LmpcQpParams P, const __grid_constant__ LmpcQpBatch a) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x;
  if (b >= a.B) return;
  if (a.skip && a.skip[b]) return;
  const int N = P.N, NS = P.NS, K = P.K;
  LmpcQpIn in;
  in.x_ic = a.x_ic + 6 * (size_t)b; in.u_ic = a.u_ic + 2 * (size_t)b;
  in.U0 = a.U0 + (2 * (size_t)NS) * b; in.T = a.T_ref + (size_t)NS * b;
  in.bl = a.bl + (size_t)N * b; in.br = a.br + (size_t)N * b; in.vref = a.vref + (size_t)N * b;
  in.ABg = a.ABg + (54 * (size_t)NS) * b;
  in.ssx = P.learning ? a.ssx + (6 * (size_t)K) * b : nullptr;
  in.ssj = P.learning ? a.ssj + (size_t)K * b : nullptr;
  in.cen = a.cen + 6 * (size_t)b; in.ss_count = a.ss_count;
  in.scratch = a.scratch + (size_t)LMPC_QP_SCRATCH(N, K) * b;
  LmpcQpOut out;
  out.X = a.X + (6 * (size_t)N) * b; out.U = a.U + (2 * (size_t)NS) * b; out.dU = a.dU + (2 * (size_t)NS) * b;
  out.lam = (a.lam && P.learning) ? a.lam + (size_t)K * b : nullptr;
  out.cost = a.cost ? a.cost + b : nullptr;
  out.status = a.status + b; out.iters = a.iters + b;
  lmpc_qp_solve<NW, KPL, NTPL, RSTPL>(P, in, sm, out);
}
