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
  const int* ss_count_v;   // optional [B]: columns found per instance (per-agent safe sets); overrides ss_count
  int B;
  const int* skip;   // optional [B]: non-zero = leave this instance's outputs untouched (converged SQP instances)
  // Fused result exchange over NVLink peer memory (lmpc_solve_gather_batch).  The trajectory outputs X, U,
  // dU, cost, status point into this rank's block of its gather buffer; every peer maps a buffer of the same layout, so
  // adding mirror_off[m] (bytes) to an output address gives the same element in peer m's buffer.  After its local
  // stores each CTA stores its result into the peers (fire-and-forget writes through the NVSwitch), fences at system
  // scope and counts itself on `done`; the CTA that completes the count publishes `seq` in every peer's flag slot of this
  // rank.  The peers' wait kernel spins on its LOCAL flags only.  n_mirror = 0: plain solve.
  int n_mirror;
  long long mirror_off[LMPC_MAX_PEERS - 1];
  unsigned long long* peer_flag[LMPC_MAX_PEERS - 1];   // address (in peer m's memory) of this rank's flag slot
  unsigned int* done;                                  // local CTA counter (reset by the last CTA); null: no completion protocol
  unsigned long long seq;
  // mirror_all != 0: the (single) mirror target is the caller's pinned HOST arena (zero-copy stores over PCIe, lmpc_solve_batch
  // with host buffers): lambda and the iteration count are mirrored too, so that nothing of the instance's result is left to a
  // device-to-host copy after the kernel -- the results of instances that finish early travel while the slow ones still iterate
  int mirror_all;
};

LMPC_DEV void lmpc_st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

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
  in.cen = a.cen + 6 * (size_t)b; in.ss_count = a.ss_count_v ? a.ss_count_v[b] : a.ss_count;
  in.scratch = a.scratch + (size_t)LMPC_QP_SCRATCH(N, K) * b;
  LmpcQpOut out;
  out.X = a.X + (6 * (size_t)N) * b; out.U = a.U + (2 * (size_t)NS) * b; out.dU = a.dU + (2 * (size_t)NS) * b;
  out.lam = (a.lam && P.learning) ? a.lam + (size_t)K * b : nullptr;
  out.cost = a.cost ? a.cost + b : nullptr;
  out.status = a.status + b; out.iters = a.iters + b; out.stats = nullptr;
  lmpc_qp_solve<NW, KPL, NTPL, RSTPL>(P, in, sm, out);
  if (a.n_mirror > 0) {
    // ---- the collective, fused: this instance's 1.6 KB of results go straight into every peer's gather buffer
    if (NW == 1) __syncwarp(); else __syncthreads();   // the group's own global stores are visible to all its lanes
    const int tid = (int)threadIdx.x, NT = 32 * NW;
    // Values are read ONCE into registers (independent loads: one L2 round trip), then stored to every peer: a loop of
    // load -> remote store per peer serialises on the load latency because the compiler may not move a load above a store
    // it cannot prove disjoint (measured on 8 GPUs: +3.5 us per peer before this form).
#define LMPC_MIR(ptr, off) (*reinterpret_cast<double*>(reinterpret_cast<char*>(ptr) + (off)))
    for (int base = 0; base < 6 * N; base += 4 * NT) {
      double v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) { const int q = base + tid + NT * j; v[j] = (q < 6 * N) ? out.X[q] : 0.0; }
      for (int m = 0; m < a.n_mirror; m++) {
        const long long off = a.mirror_off[m];
#pragma unroll
        for (int j = 0; j < 4; j++) { const int q = base + tid + NT * j; if (q < 6 * N) LMPC_MIR(out.X + q, off) = v[j]; }
      }
    }
    for (int base = 0; base < 2 * NS; base += 2 * NT) {
      double vu[2], vd[2];
#pragma unroll
      for (int j = 0; j < 2; j++) { const int q = base + tid + NT * j; const bool on = q < 2 * NS; vu[j] = on ? out.U[q] : 0.0; vd[j] = on ? out.dU[q] : 0.0; }
      for (int m = 0; m < a.n_mirror; m++) {
        const long long off = a.mirror_off[m];
#pragma unroll
        for (int j = 0; j < 2; j++) { const int q = base + tid + NT * j; if (q < 2 * NS) { LMPC_MIR(out.U + q, off) = vu[j]; LMPC_MIR(out.dU + q, off) = vd[j]; } }
      }
    }
    if (a.mirror_all && out.lam) {
      for (int base = 0; base < K; base += 4 * NT) {
        double v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { const int q = base + tid + NT * j; v[j] = (q < K) ? out.lam[q] : 0.0; }
        const long long off = a.mirror_off[0];
#pragma unroll
        for (int j = 0; j < 4; j++) { const int q = base + tid + NT * j; if (q < K) LMPC_MIR(out.lam + q, off) = v[j]; }
      }
    }
    if (tid == 0) {
      const double cv = out.cost ? *out.cost : 0.0;
      const int sv = *out.status, iv = *out.iters;
      for (int m = 0; m < a.n_mirror; m++) {
        const long long off = a.mirror_off[m];
        if (out.cost) LMPC_MIR(out.cost, off) = cv;
        *reinterpret_cast<int*>(reinterpret_cast<char*>(out.status) + off) = sv;
        if (a.mirror_all) *reinterpret_cast<int*>(reinterpret_cast<char*>(out.iters) + off) = iv;
      }
    }
#undef LMPC_MIR
    if (NW == 1) __syncwarp(); else __syncthreads();
    if (tid == 0 && a.done) {
      __threadfence_system();                                   // this group's peer stores before the count
      const unsigned int prev = atomicAdd(a.done, 1u);
      if (prev == (unsigned int)a.B - 1u) {                     // last CTA of the launch: everyone's stores are ordered before this
        *a.done = 0u;                                           // ready for the next launch (stream-ordered after this one)
        __threadfence_system();
        for (int m = 0; m < a.n_mirror; m++) lmpc_st_release_sys_u64(a.peer_flag[m], a.seq);
      }
    }
  }
}
