// lmpc_kernels.cuh -- __global__ entry points of the batched LMPC solve (sm_100a).
//
//   K1 lmpc_linearise_kernel   thread per (instance, stage): abscissa alignment + RK4 Jacobians -> A,B,g
//   K2 lmpc_ss_query_kernel    warp per (instance, lap): exact k-NN in the safe-set slab + cost-to-go gather
//   KR lmpc_regress_tiled_kernel  warp per (instance, stage), 8 per block: error-dynamics regression added to [A|B|g] (optional, between K1 and K3)
//   K3 lmpc_qp_kernel          warp group (1, 2 or 4 warps = one CTA) per instance: interior-point / Riccati solve in shared memory
//
// Batch arrays are instance-major, so a warp's (or thread's) reads of its own instance are contiguous.
#pragma once
#include <cuda_runtime.h>
#include "lmpc_model.cuh"
#include "lmpc_qp_kernel.cuh"
#include "lmpc_ss_core.cuh"
#include "lmpc_reg_core.cuh"
#include "lmpc_agents.cuh"

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
// Stage 0's thread also writes the aligned query / centre point X_ref[:, N-1] (when cen is given).
// skip: optional per-instance mask (converged SQP instances keep their linearisation).
#define LMPC_K1_THREADS 64
__global__ void __launch_bounds__(LMPC_K1_THREADS) lmpc_linearise_kernel(LmpcModel M, int B, int N, const double* __restrict__ x_ic,
                                      const double* __restrict__ X_ref, const double* __restrict__ U_ref,
                                      const double* __restrict__ T_ref, const double* __restrict__ kappa,
                                      const double* __restrict__ total_length, double* __restrict__ ABg,
                                      double* __restrict__ cen, const int* __restrict__ skip) {
  // a thread's 54 results are 432 bytes apart from its neighbour's: staged by value index ([54][65], conflict-free both
  // ways) and written out as the block's one contiguous run of 64 x 54 doubles
  __shared__ double stage[54 * (LMPC_K1_THREADS + 1)];
  __shared__ int wrote[LMPC_K1_THREADS];
  const int NS = N - 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool mine = t < B * NS && !(skip && skip[t / NS]);
  wrote[threadIdx.x] = mine ? 1 : 0;
  if (mine) {
    const int b = t / NS, i = t - b * NS;
    const double L = total_length[b], s0 = x_ic[6 * (size_t)b];
    const double* xr = X_ref + (6 * (size_t)N) * b + 6 * i;
    double xl[6], ul[2];
    for (int k = 0; k < 6; k++) xl[k] = xr[k];
    xl[0] = lmpc_align_abscissa(xl[0], s0, L);
    ul[0] = U_ref[(2 * (size_t)NS) * b + 2 * i]; ul[1] = U_ref[(2 * (size_t)NS) * b + 2 * i + 1];
    // the thread's staging column is also where the tangents are accumulated (lmpc_linearise_staged)
    lmpc_linearise_staged(M, xl, ul, kappa[(size_t)N * b + i], T_ref[(size_t)NS * b + i], stage + threadIdx.x, LMPC_K1_THREADS + 1);
    if (i == 0 && cen) {
      const double* xe = X_ref + (6 * (size_t)N) * b + 6 * (N - 1);
      double* c = cen + 6 * (size_t)b;
      c[0] = lmpc_align_abscissa(xe[0], s0, L);
      for (int k = 1; k < 6; k++) c[k] = xe[k];
    }
  }
  __syncthreads();
  double* o = ABg + 54 * (size_t)blockIdx.x * LMPC_K1_THREADS;   // item t's block is ABg + 54 t
  for (int e = threadIdx.x; e < 54 * LMPC_K1_THREADS; e += LMPC_K1_THREADS) {
    const int th = e / 54, k = e - 54 * th;
    if (wrote[th]) o[e] = stage[k * (LMPC_K1_THREADS + 1) + th];
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

// ---- K2 inside the tick: the query point is the abscissa-aligned X_ref[:, N-1] (racing_mpc.cpp:219-223,249-255), formed
// here from the tick's inputs so that the kernel does not depend on K1 and can run beside it on a second stream
__global__ void lmpc_ss_query_tick_kernel(LmpcLapTable tab, int B, int N, const double* __restrict__ x_ic,
                                          const double* __restrict__ X_ref, const double* __restrict__ total_length,
                                          int max_total, int pad_to, double* __restrict__ ss_x, double* __restrict__ ss_j) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= B * tab.n_used) return;
  const int b = w / tab.n_used, j = w - b * tab.n_used;
  const double* xe = X_ref + (6 * (size_t)N) * b + 6 * (N - 1);
  const double qs = lmpc_align_abscissa(xe[0], x_ic[6 * (size_t)b], total_length[b]), qe = xe[1];
  lmpc_ss_query_warp(tab.lap[j], qs, qe, max_total, ss_x + (6 * (size_t)pad_to) * b, ss_j + (size_t)pad_to * b,
                     j == tab.n_used - 1, tab.count, pad_to);
}

// ---- SQP to convergence (the B200 counterpart of the reference's one-off full-dynamics IPOPT solve,
// racing_mpc.cpp:67-84,162-166): the tick's QP re-linearised at its own solution.  One thread per instance.
// init: linearisation point <- (abscissa-aligned X_ref, U_ref)
__global__ void lmpc_sqp_init_kernel(int B, int N, const double* __restrict__ x_ic, const double* __restrict__ X_ref,
                                     const double* __restrict__ U_ref, const double* __restrict__ total_length,
                                     double* __restrict__ Xk, double* __restrict__ Uk, double* __restrict__ Dprev,
                                     double* __restrict__ alpha_, int* __restrict__ done, int* __restrict__ its) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NS = N - 1;
  const double L = total_length[b], s0 = x_ic[6 * (size_t)b];
  const double* xr = X_ref + (6 * (size_t)N) * b; double* xk = Xk + (6 * (size_t)N) * b;
  for (int i = 0; i < N; i++) {
    xk[6 * i] = lmpc_align_abscissa(xr[6 * i], s0, L);
    for (int c = 1; c < 6; c++) xk[6 * i + c] = xr[6 * i + c];
  }
  for (int q = 0; q < 2 * NS; q++) Uk[(2 * (size_t)NS) * b + q] = U_ref[(2 * (size_t)NS) * b + q];
  for (int q = 0; q < 6 * N + 2 * NS; q++) Dprev[(size_t)(6 * N + 2 * NS) * b + q] = 0.0;
  alpha_[b] = 1.0; done[b] = 0; its[b] = 0;
}

// update: d = QP solution - linearisation point; step = max |d| / max(1, |new|) over X and U; the linearisation point
// moves by alpha d, alpha following successive displacements: (anti)parallel (|cos| > 0.9) -> the secant step length that
// cancels the dominant mode; else cos < -0.25 (oscillation of the Gauss-Newton iteration) -> halve, >= 1/8; cos > 0.25 ->
// double, <= 1.
// done: 1 = converged (step < tol), 2 = the QP failed (its status stays in the output).
__global__ void lmpc_sqp_update_kernel(int B, int N, double tol, const double* __restrict__ X, const double* __restrict__ U,
                                       const int* __restrict__ status, double* __restrict__ Xk, double* __restrict__ Uk,
                                       double* __restrict__ Dprev, double* __restrict__ alpha_, int* __restrict__ done,
                                       int* __restrict__ its, int* __restrict__ n_active) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || done[b]) return;
  const int k = its[b];
  its[b] = k + 1;
  if (status[b] != LMPC_SOLVED && status[b] != LMPC_SOLVED_INACCURATE) { done[b] = 2; return; }
  const int NS = N - 1, nx = 6 * N, nd = 6 * N + 2 * NS;
  const double* xn = X + (size_t)nx * b; const double* un = U + (2 * (size_t)NS) * b;
  double* xk = Xk + (size_t)nx * b; double* uk = Uk + (2 * (size_t)NS) * b; double* dpv = Dprev + (size_t)nd * b;
  double step = 0.0, dd = 0.0, dp = 0.0, pp = 0.0;
  for (int q = 0; q < nd; q++) {
    const double nv = q < nx ? xn[q] : un[q - nx], ov = q < nx ? xk[q] : uk[q - nx];
    const double d = nv - ov, pv = dpv[q];
    step = fmax(step, fabs(d) / fmax(1.0, fabs(nv)));
    dd += d * d; dp += d * pv; pp += pv * pv;
    dpv[q] = d;
  }
  if (step < tol) { done[b] = 1; return; }
  double alpha = alpha_[b];
  if (k > 0) {
    // successive displacements (anti)parallel: one mode d_k = (1 - alpha (1 + rho)) d_{k-1} dominates; the secant step
    // alpha / (1 - d_k.d_{k-1} / |d_{k-1}|^2) = 1 / (1 + rho) cancels it.  Otherwise halve on oscillation, double on progress.
    const double cs = dp / sqrt(dd * pp + 1e-300);
    const double sr = dp / (pp + 1e-300);
    if (fabs(cs) > 0.9) alpha = (1.0 - sr > 0.1) ? fmin(fmax(alpha / (1.0 - sr), 0.125), 1.0) : 1.0;
    else if (cs < -0.25) alpha = fmax(0.5 * alpha, 0.125);
    else if (cs > 0.25) alpha = fmin(2.0 * alpha, 1.0);
    alpha_[b] = alpha;
  }
  for (int q = 0; q < nx; q++) xk[q] += alpha * dpv[q];
  for (int q = 0; q < 2 * NS; q++) uk[q] += alpha * dpv[nx + q];
  atomicAdd(n_active, 1);
}

// after the last pass: an instance whose step test never passed is not a solution of the nonlinear problem (IPOPT's
// "Maximum_Iterations_Exceeded" under error_on_fail, racing_mpc.cpp:71): LMPC_SQP_MAX_ITER
__global__ void lmpc_sqp_finalize_kernel(int B, const int* __restrict__ done, int* __restrict__ status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (done[b] == 0 && (status[b] == LMPC_SOLVED || status[b] == LMPC_SOLVED_INACCURATE)) status[b] = LMPC_SQP_MAX_ITER;
}

// defect of the nonlinear dynamics at the returned trajectory: max_i,c |x_{i+1} - f_d(x_i, u_i, kappa_i, T_i)|_c
__global__ void lmpc_sqp_defect_kernel(LmpcModel M, int B, int N, const double* __restrict__ X, const double* __restrict__ U,
                                       const double* __restrict__ T_ref, const double* __restrict__ kappa, double* __restrict__ defect) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int NS = N - 1;
  double dmax = 0.0;
  for (int i = 0; i < NS; i++) {
    double xl[6], ul[2], xn[6];
    for (int c = 0; c < 6; c++) xl[c] = X[(6 * (size_t)N) * b + 6 * i + c];
    ul[0] = U[(2 * (size_t)NS) * b + 2 * i]; ul[1] = U[(2 * (size_t)NS) * b + 2 * i + 1];
    lmpc_step(M, xl, ul, kappa[(size_t)N * b + i], T_ref[(size_t)NS * b + i], xn);
    for (int c = 0; c < 6; c++) dmax = fmax(dmax, fabs(xn[c] - X[(6 * (size_t)N) * b + 6 * (i + 1) + c]));
  }
  defect[b] = dmax;
}

// ---- error-dynamics regression (lmpc_reg_core.cuh)
// prepare: thread per stored sample p with a successor: E[:, p] = x_{p+1} - f_d(x_p, u_p, k_p, t_{p+1} - t_p).
// Z [8][ld], Xn [6][ld], E [6][ld] by column.
__global__ void lmpc_reg_prepare_kernel(LmpcModel M, int n, int ld, const double* __restrict__ Z, const double* __restrict__ Xn,
                                        const double* __restrict__ kappa, const double* __restrict__ dt, double* __restrict__ E) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double xl[6], ul[2], xn[6];
  for (int c = 0; c < 6; c++) xl[c] = Z[(size_t)c * ld + p];
  ul[0] = Z[(size_t)6 * ld + p]; ul[1] = Z[(size_t)7 * ld + p];
  lmpc_step(M, xl, ul, kappa[p], dt[p], xn);
  for (int c = 0; c < 6; c++) E[(size_t)c * ld + p] = Xn[(size_t)c * ld + p] - xn[c];
}

// KR: every query scans the stored samples (its window of them when they are sorted).  A block of LMPC_REG_WARPS warps (one
// query each) streams the samples through shared-memory tiles that hold, by column, only the regression's own inputs and
// its output's error: loaded once per block (8x less L2 traffic than every warp pulling the samples itself), by cp.async
// into the other of two buffers while the current tile is scanned, read conflict-free by the scan (lane = sample).
// DD = size class of the plan (5: at most four regressors + 1 -- 20 accumulators per lane; 9: the general case, 54).
#define LMPC_REG_TILE_OF(DD) ((DD) == 5 ? 480 : 256)   // two buffers of 4 + 2 columns stay inside the 48 KB of static shared memory
#define LMPC_REG_WARPS 8
__device__ __forceinline__ void lmpc_cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void lmpc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void lmpc_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
struct LmpcRegItems {          // where the items of a launch live
  int n;                       // items
  int tick;                    // 1: item = (instance b, stage i) of a tick, query at the aligned linearisation point, [A|B|g] in ABg
  const double *xq, *uq; double *A, *Bm, *C; int* npts;                                                  // generic items
  int B, N; const double *x_ic, *X_ref, *U_ref, *total_length; double* ABg; const int* skip;              // tick items
};
template <int DD, int MINB = 1>
__global__ void __launch_bounds__(32 * LMPC_REG_WARPS, MINB) lmpc_regress_tiled_kernel(LmpcRegPlan plan, LmpcRegView v, LmpcRegItems it) {
  constexpr int TILE = LMPC_REG_TILE_OF(DD);
  __shared__ __align__(16) double tZ[2][(DD - 1) * TILE];
  __shared__ __align__(16) double tE[2][TILE];
  __shared__ __align__(16) double tE2[DD == 5 ? 2 : 1][DD == 5 ? TILE : 1];   // the paired regression's output (size class 5 only)
  constexpr int D = DD, NQ = D * (D + 1) / 2, NV = NQ + D + 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int item = blockIdx.x * LMPC_REG_WARPS + w;
  bool live = item < it.n;
  double zq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double *A = nullptr, *Bm = nullptr, *C = nullptr; int* npts = nullptr;
  if (live) {
    if (it.tick) {
      const int NS = it.N - 1, b = item / NS, i = item - b * NS;
      if (it.skip && it.skip[b]) live = false;
      const double* xr = it.X_ref + (6 * (size_t)it.N) * b + 6 * i;
      for (int c = 0; c < 6; c++) zq[c] = xr[c];
      zq[0] = lmpc_align_abscissa(zq[0], it.x_ic[6 * (size_t)b], it.total_length[b]);
      zq[6] = it.U_ref[(2 * (size_t)NS) * b + 2 * i]; zq[7] = it.U_ref[(2 * (size_t)NS) * b + 2 * i + 1];
      A = it.ABg + (54 * (size_t)NS) * b + 54 * i; Bm = A + 36; C = A + 48;
    } else {
      for (int c = 0; c < 6; c++) zq[c] = it.xq[6 * (size_t)item + c];
      zq[6] = it.uq[2 * (size_t)item]; zq[7] = it.uq[2 * (size_t)item + 1];
      A = it.A + 36 * (size_t)item; Bm = it.Bm + 12 * (size_t)item; C = it.C + 6 * (size_t)item;
      npts = it.npts ? it.npts + (size_t)plan.n_out * item : nullptr;
    }
  }
  const double h = plan.h, ih = 1.0 / h, kc = 0.75 / h;
  // the block's window of the sorted samples: the union of its queries' windows |z[sort_dim] - z_q[sort_dim]| < h (two
  // binary searches per query, lane 0), the whole range when the samples are not sorted
  __shared__ int win[2];
  if (threadIdx.x == 0) { win[0] = v.sort_dim >= 0 ? v.M : 0; win[1] = v.sort_dim >= 0 ? 0 : v.M; }
  __syncthreads();
  if (v.sort_dim >= 0 && live && lane == 0) {
    const double* key = v.Z + (size_t)v.sort_dim * v.ld;
    const double ql = zq[v.sort_dim] - h, qh = zq[v.sort_dim] + h;
    int a = 0, b = v.M;
    while (a < b) { const int m = (a + b) >> 1; if (key[m] < ql) a = m + 1; else b = m; }   // first key >= ql
    const int lo = a;
    b = v.M;
    while (a < b) { const int m = (a + b) >> 1; if (key[m] <= qh) a = m + 1; else b = m; }  // first key > qh
    atomicMin(&win[0], lo); atomicMax(&win[1], a);
  }
  __syncthreads();
  for (int r = 0; r < plan.n_out; r++) {
    const LmpcRegRow& row = plan.row[r];
    // a regression with the input lists of an earlier one was scanned with it (same regressors, same weights, same M'KM)
    const int fo = (DD == 5) ? row.follower : -1;
    if (DD == 5 && row.lead != r) continue;
    double q[D], Q[NQ], bv[D], bv2[DD == 5 ? D : 1], cnt = 0.0;
    bool uses_key = false;
#pragma unroll
    for (int a = 0; a < D; a++) { q[a] = (row.sel[a] < 8) ? zq[row.sel[a]] : 0.0; bv[a] = 0.0; if (DD == 5) bv2[a] = 0.0; uses_key |= (row.sel[a] == v.sort_dim); }
#pragma unroll
    for (int k = 0; k < NQ; k++) Q[k] = 0.0;
    // a regression that does not use the sort component must see every sample
    const int t_begin = (uses_key && v.sort_dim >= 0) ? (win[0] & ~31) : 0, t_end = (uses_key && v.sort_dim >= 0) ? win[1] : v.M;
    // tile k + 1 is in flight (cp.async into the other buffer) while tile k is scanned; one barrier per tile: past it every
    // warp has finished tile k - 1, whose buffer the next copies overwrite
    auto issue = [&](int buf, int t0) {
      const int count = min(TILE, t_end - t0);
      for (int a = 0; a < row.D - 1; a++) {
        const double* src = v.Z + (size_t)row.sel[a] * v.ld + t0;
        for (int e = threadIdx.x; e < count; e += 32 * LMPC_REG_WARPS) lmpc_cp_async8(&tZ[buf][a * TILE + e], src + e);
      }
      const double* src = v.E + (size_t)row.out * v.ld + t0;
      for (int e = threadIdx.x; e < count; e += 32 * LMPC_REG_WARPS) lmpc_cp_async8(&tE[buf][e], src + e);
      if (DD == 5 && fo >= 0) {
        const double* src2 = v.E + (size_t)plan.row[fo].out * v.ld + t0;
        for (int e = threadIdx.x; e < count; e += 32 * LMPC_REG_WARPS) lmpc_cp_async8(&tE2[buf][e], src2 + e);
      }
      lmpc_cp_async_commit();
    };
    __syncthreads();   // the previous regression's last tile has been scanned by every warp
    if (t_begin < t_end) issue(0, t_begin);
    int buf = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += TILE, buf ^= 1) {
      const int count = min(TILE, t_end - t0);
      lmpc_cp_async_wait_all();
      __syncthreads();
      if (t0 + TILE < t_end) issue(buf ^ 1, t0 + TILE);
      if (live) {
        const double *cZ = tZ[buf], *cE = tE[buf];
        if constexpr (DD == 5) {   // exact-size scans (warp-uniform switch): no index lists in the loop
          const double* cE2 = tE2[buf];
#define LMPC_REG_SCAN_(DE_)                                                                                              \
  if (fo >= 0) lmpc_reg_scan_tile<DE_, DD, TILE, true>(h, ih, kc, q, cZ, cE, cE2, count, lane, Q, bv, bv2, cnt);         \
  else lmpc_reg_scan_tile<DE_, DD, TILE, false>(h, ih, kc, q, cZ, cE, cE2, count, lane, Q, bv, bv2, cnt);
          switch (row.D) {
            case 5: LMPC_REG_SCAN_(5) break;
            case 4: LMPC_REG_SCAN_(4) break;
            case 3: LMPC_REG_SCAN_(3) break;
            case 2: LMPC_REG_SCAN_(2) break;
            default: LMPC_REG_SCAN_(1) break;
          }
#undef LMPC_REG_SCAN_
        } else lmpc_reg_scan_lane<DD, true>(row, h, ih, kc, q, cZ, cE, TILE, count, lane, Q, bv, cnt);
      }
    }
    if (live) {   // warp-uniform: a warp is one item
      LaneVar<double> acc[NV];
#pragma unroll
      for (int k = 0; k < NQ; k++) acc[k].v = Q[k];
#pragma unroll
      for (int a = 0; a < D; a++) acc[NQ + a].v = bv[a];
      acc[NQ + D].v = cnt;
      lmpc_reg_finish<DD>(plan, row, acc, A, Bm, C, npts ? npts + r : nullptr);
      if (DD == 5 && fo >= 0) {   // the paired regression: the same sums, its own M'Ky
#pragma unroll
        for (int k = 0; k < NQ; k++) acc[k].v = Q[k];
#pragma unroll
        for (int a = 0; a < D; a++) acc[NQ + a].v = bv2[DD == 5 ? a : 0];
        acc[NQ + D].v = cnt;
        lmpc_reg_finish<DD>(plan, plan.row[fo], acc, A, Bm, C, npts ? npts + fo : nullptr);
      }
    }
  }
}
