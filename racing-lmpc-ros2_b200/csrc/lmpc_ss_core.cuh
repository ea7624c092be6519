// lmpc_ss_core.cuh -- safe-set nearest-neighbour query, one warp per (query, lap).
//
// Replaces SafeSetManager::query(SSQuery) -> SSTrajectory::query -> CGAL
// Orthogonal_k_neighbor_search (reference src/vehicle_dynamics_models/racing_trajectory/src/
// safe_set.cpp:42-54,153-180 and trajectory_kd_tree.cpp:53-63): for every stored lap, newest first,
// the `per_lap` Euclidean-nearest points to (s, e_y) in the lap's tripled set (s-L, s, s+L),
// nearest first, concatenated and truncated to `max_total` columns.  The search is exact brute force
// over the lap's device-resident slab: every lane keeps its T best candidates from a strided scan,
// the warp merges them by repeated arg-min; if a lane ran out of candidates while the merge was
// still selecting (its (T+1)-th best might qualify) the warp falls back to an exact threshold scan.
// Ties are broken by the lower point index; exact duplicates of (s, e_y) resolve to the first
// inserted index through the canon[] table, as the reference's coordinate hash does
// (trajectory_kd_tree.cpp:38,60-62).
#pragma once
#include "lmpc_warp.cuh"

#define LMPC_SS_T 4   // candidates kept per lane
#ifndef LMPC_SS_PPT
#define LMPC_SS_PPT 8 // points per lane and trip of the scan
#endif

struct LmpcLapView {
  const double* ps;     // [m] s of the tripled points
  const double* pe;     // [m] e_y
  const double* xr;     // [m][6] payload (x_repeat)
  const double* J;      // [m] cost-to-go
  const int* canon;     // [m] first index with the same (s, e_y)
  int m;                // 3 n
  int take;             // min(per_lap, m)
  int out_off;          // first output column of this lap
};

// One (query, lap).  Writes columns [out_off, min(out_off + take, max_total)) of ss_x / ss_j and,
// when `last` is set, pads the tail [count, pad_to) with the last written column.
LMPC_DEV void lmpc_ss_query_warp(const LmpcLapView& lap, double qs, double qe, int max_total, double* ss_x,
                                 double* ss_j, bool last, int count, int pad_to) {
  const int m = lap.m;
  // ---- per-lane sorted top-T of a strided scan
  LaneVar<double> c0, c1, c2, c3;   // ascending distances
  LaneVar<int> j0, j1, j2, j3;
  LANES_BEGIN
    double a0 = 1e300, a1 = 1e300, a2 = 1e300, a3 = 1e300;
    int b0 = 1 << 30, b1 = 1 << 30, b2 = 1 << 30, b3 = 1 << 30;
    // LMPC_SS_PPT points per trip: their loads go out together (the scan waits for the L2, not for arithmetic), then the
    // (branchy, usually one-test) insertions in index order
    for (int base = lane; base < m; base += 32 * LMPC_SS_PPT) {
      double vv[LMPC_SS_PPT];
#pragma unroll
      for (int u = 0; u < LMPC_SS_PPT; u++) {
        const int idx = base + 32 * u;
        const int ic = idx < m ? idx : base;   // clamp: always a valid address
        const double ds = qs - lap.ps[ic], de = qe - lap.pe[ic];
        vv[u] = idx < m ? ds * ds + de * de : 1e300;
      }
#pragma unroll
      for (int u = 0; u < LMPC_SS_PPT; u++) {
        const int idx = base + 32 * u;
        const double v = vv[u];
        // strided indices increase, so on equal distance the earlier (lower) index stays ahead
        if (v < a3) {
          if (v < a2) {
            a3 = a2; b3 = b2;
            if (v < a1) {
              a2 = a1; b2 = b1;
              if (v < a0) { a1 = a0; b1 = b0; a0 = v; b0 = idx; } else { a1 = v; b1 = idx; }
            } else { a2 = v; b2 = idx; }
          } else { a3 = v; b3 = idx; }
        }
      }
    }
    c0(lane) = a0; c1(lane) = a1; c2(lane) = a2; c3(lane) = a3;
    j0(lane) = b0; j1(lane) = b1; j2(lane) = b2; j3(lane) = b3;
  LANES_END
  // ---- merge: take rounds of warp arg-min over the lane heads
  LaneVar<int> sel;        // lane r keeps the index of rank r (take <= 32) -- ranks beyond 31 are written directly
  LaneVar<int> exhausted;
  LANES_BEGIN
    sel(lane) = -1; exhausted(lane) = 0;
  LANES_END
  bool inexact = false;
  double last_d = -1.0; int last_i = -1;
  int r = 0;
  for (; r < lap.take; r++) {
    LaneVar<double> hv; LaneVar<int> hi;
    LANES_BEGIN
      hv(lane) = c0(lane); hi(lane) = j0(lane);
    LANES_END
    warp_argmin_nonneg(hv, hi);
    const int wi = hi(0); const double wv = hv(0);
    LaneVar<int> anyex;
    LANES_BEGIN
      if (j0(lane) == wi) {   // pop the winner's head
        c0(lane) = c1(lane); j0(lane) = j1(lane); c1(lane) = c2(lane); j1(lane) = j2(lane);
        c2(lane) = c3(lane); j2(lane) = j3(lane); c3(lane) = 1e300; j3(lane) = 1 << 30;
        if (j0(lane) == (1 << 30) && (lane + 32 * LMPC_SS_T) < m) exhausted(lane) = 1;   // had more points than T
      }
      anyex(lane) = exhausted(lane);
    LANES_END
    warp_or(anyex);
    last_d = wv; last_i = wi;
    // ranks 0..31: lane r keeps the winner of rank r, the payload is gathered for all ranks at once after the merge
    // (a gather per round would put two dependent global loads on the critical path of every round)
    if (r < 32) {
      LANES_BEGIN
        if (lane == r) sel(lane) = wi;
      LANES_END
    } else {
      const int col = lap.out_off + r;
      if (col < max_total) {
        LANES_BEGIN
          const int src = (wi >= 0 && wi < m) ? lap.canon[wi] : 0;   // NaN query: no valid winner
          if (lane < 6) ss_x[6 * col + lane] = lap.xr[6 * src + lane];
          else if (lane == 6) ss_j[col] = lap.J[src];
        LANES_END
      }
    }
    // the winner above is exact (every head was valid); a lane that just ran out of kept candidates
    // may hold unseen closer points, so the remaining ranks take the exact path
    if (anyex(0)) { inexact = true; r++; break; }
  }
  if (inexact) {
    // exact continuation: rank r.. by repeated scan for the smallest (d2, idx) strictly after the last one
    for (; r < lap.take; r++) {
      LaneVar<double> hv; LaneVar<int> hi;
      LANES_BEGIN
        double bv = 1e300; int bi = 1 << 30;
        for (int idx = lane; idx < m; idx += 32) {
          const double ds = qs - lap.ps[idx], de = qe - lap.pe[idx];
          const double v = ds * ds + de * de;
          const bool after = (v > last_d) || (v == last_d && idx > last_i);
          if (after && (v < bv || (v == bv && idx < bi))) { bv = v; bi = idx; }
        }
        hv(lane) = bv; hi(lane) = bi;
      LANES_END
      warp_argmin_nonneg(hv, hi);
      last_d = hv(0); last_i = hi(0);
      const int wi = last_i;
      if (r < 32) {
        LANES_BEGIN
          if (lane == r) sel(lane) = wi;
        LANES_END
      } else {
        const int col = lap.out_off + r;
        if (col < max_total) {
          LANES_BEGIN
            const int src = (wi >= 0 && wi < m) ? lap.canon[wi] : 0;   // NaN query: no valid winner
            if (lane < 6) ss_x[6 * col + lane] = lap.xr[6 * src + lane];
            else if (lane == 6) ss_j[col] = lap.J[src];
          LANES_END
        }
      }
    }
  }
  // ---- payload of ranks 0..31, one lane per rank
  LANES_BEGIN
    const int col = lap.out_off + lane;
    if (lane < lap.take && col < max_total) {
      const int wi = sel(lane);
      const int src = (wi >= 0 && wi < m) ? lap.canon[wi] : 0;   // NaN query: no valid winner
      for (int c = 0; c < 6; c++) ss_x[6 * col + c] = lap.xr[6 * src + c];
      ss_j[col] = lap.J[src];
    }
  LANES_END
  // ---- pad with the last column (racing_mpc.cpp:263-272)
  if (last && count > 0 && count < pad_to) {
    LANES_BEGIN
      for (int o = lane; o < 7 * (pad_to - count); o += 32) {
        const int col = count + o / 7, c = o % 7;
        if (c < 6) ss_x[6 * col + c] = ss_x[6 * (count - 1) + c]; else ss_j[col] = ss_j[count - 1];
      }
    LANES_END
  }
}
