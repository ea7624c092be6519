// lmpc_agents.cuh -- per-agent safe sets and lap recording on the device (Monte-Carlo LMPC: every agent learns from
// its OWN laps).
//
// In the reference every RacingMPC owns a SafeSetManager and a SafeSetRecorder (racing_mpc.hpp:99-100) and feeds the
// recorder on every solve (racing_mpc.cpp:245-246).  For B agents in one closed loop that is B recorders and B safe sets:
//   recorder   SafeSetRecorder::step (safe_set.cpp:278-322): lap segmentation by the abscissa wrap, the samples before the
//              first wrap discarded, a completed lap handed to add_lap, the new lap started with the wrapping sample
//   add_lap    SSTrajectory::process_lap_data (safe_set.cpp:116-137): x_repeat = [x - L e0, x, x + L e0],
//              J = [J + n - 1, J, J - n + 1], J_j = n - 1 - j; boost::circular_buffer of max_lap_stored laps (:139-151)
//   query      SafeSetManager::query(SSQuery) (:153-180): newest lap first, num_ss_pts_per_lap nearest per lap, until
//              num_ss_pts columns are found -- lmpc_ss_query_warp of lmpc_ss_core.cuh on the agent's own slots
// Storage per agent: `slots` lap slots of 3 * cap tripled points (SoA keys + payload, as the shared slab) and one
// recording buffer of cap samples; 8192 agents x 3 slots x 1024 samples = 5.7 GB of the 180 GB.
#pragma once
#include "lmpc_ss_core.cuh"

struct LmpcAgentSets {
  int B, slots, cap;     // agents, lap slots per agent (= max_lap_stored), samples per lap (capacity)
  // stored laps, [B][slots][3 cap] (xr: [..][6])
  double *ps, *pe, *J, *xr;
  int* canon;
  int* n;                // [B][slots] samples of the lap in the slot
  int* head;             // [B] slot of the oldest stored lap
  int* count;            // [B] laps stored (<= slots)
  // recorder
  double *rec_x, *rec_u, *rec_k, *rec_t;   // [B][cap][6|2|1|1] the lap being recorded
  double* carry;         // [B][10] the sample that started the next lap while a completed lap waits to be stored
  int* rec_n;            // [B]
  double* last_px;       // [B] abscissa of the last sample
  int* flags;            // [B] bit 0 last_x_valid, 1 initialized, 2 a completed lap waits (rec_n samples), 3 a lap overflowed cap
  int* lap_count;        // [B] SafeSetRecorder::lap_count_
};

#define LMPC_AG_VALID 1
#define LMPC_AG_INIT 2
#define LMPC_AG_PENDING 4
#define LMPC_AG_OVERFLOW 8

LMPC_HD size_t lmpc_ag_slot(const LmpcAgentSets& A, int b, int slot) { return ((size_t)b * A.slots + slot) * (size_t)(3 * A.cap); }

#if !defined(LMPC_EMULATE)
// ---- recorder: one thread per agent and tick.  Fed with what RacingMPC::solve feeds SafeSetRecorder::step:
// (x_ic, u_ic, curvatures(0), t_ic, total_length) (racing_mpc.cpp:245-246).  log_rec (optional) [ticks][B][10] keeps exactly
// those values so that a host recorder can be replayed on them.
__global__ void lmpc_agents_record_kernel(LmpcAgentSets A, const double* __restrict__ x_ic, const double* __restrict__ u_ic,
                                          const double* __restrict__ kap, int N, const double* __restrict__ total_length, double t0,
                                          double dt, const int* __restrict__ tick, double* __restrict__ log_rec, int log_ticks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= A.B) return;
  const int tk = *tick;
  double s[10];
  for (int c = 0; c < 6; c++) s[c] = x_ic[6 * (size_t)b + c];
  s[6] = u_ic[2 * (size_t)b]; s[7] = u_ic[2 * (size_t)b + 1];
  s[8] = kap[(size_t)N * b];                    // curvatures(0)
  s[9] = t0 + dt * (double)tk;                  // t_ic
  if (log_rec && tk < log_ticks) for (int c = 0; c < 10; c++) log_rec[((size_t)tk * A.B + b) * 10 + c] = s[c];
  int fl = A.flags[b];
  const double L = total_length[b];
  if (!(fl & LMPC_AG_VALID)) {                  // safe_set.cpp:282-286: the first sample only arms the wrap test
    A.last_px[b] = s[0]; A.flags[b] = fl | LMPC_AG_VALID;
    return;
  }
  if (A.last_px[b] - s[0] > 0.5 * L) {          // :290 new lap
    if (fl & LMPC_AG_INIT) {                    // :293-309 the completed lap goes to the safe set (next kernel)
      fl |= LMPC_AG_PENDING;
      for (int c = 0; c < 10; c++) A.carry[10 * (size_t)b + c] = s[c];
    } else {
      fl |= LMPC_AG_INIT;
      A.rec_n[b] = 0;
    }
    A.lap_count[b] += 1;
    if (!(fl & LMPC_AG_PENDING)) {              // :312-315 the new lap starts with this sample
      const size_t o = (size_t)b * A.cap;
      for (int c = 0; c < 6; c++) A.rec_x[6 * o + c] = s[c];
      A.rec_u[2 * o] = s[6]; A.rec_u[2 * o + 1] = s[7]; A.rec_k[o] = s[8]; A.rec_t[o] = s[9];
      A.rec_n[b] = 1;
    }
  } else if (fl & LMPC_AG_INIT) {               // :317-320 (samples before the first wrap are never used: not stored)
    const int q = A.rec_n[b];
    if (q < A.cap) {
      const size_t o = (size_t)b * A.cap + q;
      for (int c = 0; c < 6; c++) A.rec_x[6 * o + c] = s[c];
      A.rec_u[2 * o] = s[6]; A.rec_u[2 * o + 1] = s[7]; A.rec_k[o] = s[8]; A.rec_t[o] = s[9];
      A.rec_n[b] = q + 1;
    } else fl |= LMPC_AG_OVERFLOW;              // lap longer than the capacity: truncated, flagged
  }
  A.last_px[b] = s[0];
  A.flags[b] = fl;
}

// ---- add_lap: one warp per agent; does nothing unless the agent's recorder has a completed lap waiting.
// Also used to seed every agent with a lap given by the host (src_* non-null: n_src samples shared by all agents).
__global__ void lmpc_agents_add_lap_kernel(LmpcAgentSets A, const double* __restrict__ total_length, const double* __restrict__ src_x,
                                           int n_src, double L_src) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= A.B) return;
  const bool seed = src_x != nullptr;
  if (!seed && !(A.flags[b] & LMPC_AG_PENDING)) return;
  const int n = seed ? n_src : A.rec_n[b];
  const double L = seed ? L_src : total_length[b];
  const double* x = seed ? src_x : A.rec_x + 6 * (size_t)b * A.cap;
  // boost::circular_buffer::push_back (safe_set.cpp:150): the oldest lap is overwritten when the buffer is full
  const int cnt = A.count[b], hd = A.head[b];
  const int slot = (cnt < A.slots) ? (hd + cnt) % A.slots : hd;
  const size_t o = lmpc_ag_slot(A, b, slot);
  const int m = 3 * n;
  for (int q = lane; q < m; q += 32) {          // SSTrajectory::process_lap_data (safe_set.cpp:116-137)
    const int rep = q / n, j = q - rep * n;
    for (int c = 0; c < 6; c++) A.xr[6 * (o + q) + c] = x[6 * (size_t)j + c];
    const double sx = x[6 * (size_t)j] + (double)(rep - 1) * L;
    A.xr[6 * (o + q)] = sx;
    A.ps[o + q] = sx; A.pe[o + q] = x[6 * (size_t)j + 1];
    A.J[o + q] = (double)(n - 1 - j) + (double)(1 - rep) * (double)(n - 1);
  }
  __syncwarp();
  // exact-duplicate keys resolve to the first inserted index (trajectory_kd_tree.cpp:38)
  for (int q = lane; q < m; q += 32) {
    const double a = A.ps[o + q], e = A.pe[o + q];
    int first = q;
    for (int r = 0; r < q; r++) if (A.ps[o + r] == a && A.pe[o + r] == e) { first = r; break; }
    A.canon[o + q] = first;
  }
  __syncwarp();
  if (lane == 0) {
    A.n[(size_t)b * A.slots + slot] = n;
    if (cnt < A.slots) A.count[b] = cnt + 1; else A.head[b] = (hd + 1) % A.slots;
    if (!seed) {                                // the new lap starts with the sample that wrapped (:312-315)
      const size_t r0 = (size_t)b * A.cap;
      const double* s = A.carry + 10 * (size_t)b;
      for (int c = 0; c < 6; c++) A.rec_x[6 * r0 + c] = s[c];
      A.rec_u[2 * r0] = s[6]; A.rec_u[2 * r0 + 1] = s[7]; A.rec_k[r0] = s[8]; A.rec_t[r0] = s[9];
      A.rec_n[b] = 1;
      A.flags[b] &= ~LMPC_AG_PENDING;
    }
  }
}

// ---- query: one warp per (agent, j-th newest lap), j < max_used = min(slots, ceil(max_total / per_lap)).
// query_stride / N as in the tick kernel: the query point is X_ref[:, N-1] with its abscissa aligned to x_ic, or (when
// x_ic is null) the raw pair query[b][0..1].  ss_count [B]: columns found per agent (0: no laps -> LMPC_NO_SAFE_SET).
__global__ void lmpc_agents_query_kernel(LmpcAgentSets A, int max_used, int per_lap, int N, const double* __restrict__ x_ic,
                                         const double* __restrict__ X_ref, const double* __restrict__ total_length,
                                         const double* __restrict__ query, int max_total, int pad_to, double* __restrict__ ss_x,
                                         double* __restrict__ ss_j, int* __restrict__ ss_count) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= A.B * max_used) return;
  const int b = w / max_used, j = w - b * max_used;
  const int cnt = A.count[b], hd = A.head[b];
  // laps newest first while fewer than max_total columns have been found (safe_set.cpp:164)
  int off = 0, my_off = -1, my_take = 0, my_slot = 0, total = 0, last_j = -1;
  for (int q = 0; q < cnt && q < max_used && total < max_total; q++) {
    const int slot = (hd + cnt - 1 - q) % A.slots;
    const int m = 3 * A.n[(size_t)b * A.slots + slot];
    const int take = per_lap < m ? per_lap : m;
    if (q == j) { my_off = off; my_take = take; my_slot = slot; }
    off += take; total += take; last_j = q;
  }
  const int found = total < max_total ? total : max_total;
  if (j == 0 && (threadIdx.x & 31) == 0) ss_count[b] = found;
  if (my_off < 0) return;   // this lap does not contribute
  double qs, qe;
  if (x_ic) {
    const double* xe = X_ref + (6 * (size_t)N) * b + 6 * (N - 1);
    qs = lmpc_align_abscissa(xe[0], x_ic[6 * (size_t)b], total_length[b]); qe = xe[1];
  } else { qs = query[2 * (size_t)b]; qe = query[2 * (size_t)b + 1]; }
  const size_t o = lmpc_ag_slot(A, b, my_slot);
  LmpcLapView lap = {A.ps + o, A.pe + o, A.xr + 6 * o, A.J + o, A.canon + o, 3 * A.n[(size_t)b * A.slots + my_slot], my_take, my_off};
  lmpc_ss_query_warp(lap, qs, qe, max_total, ss_x + (6 * (size_t)pad_to) * b, ss_j + (size_t)pad_to * b, j == last_j, found, pad_to);
}
#endif
