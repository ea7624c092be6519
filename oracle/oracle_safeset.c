/*
 * oracle_safeset.c -- CPU ORACLE (test infrastructure only; see lmpc_oracle.h).
 *
 * Restates SafeSetManager / SSTrajectory / TrajectoryKDTree of the reference
 * (src/vehicle_dynamics_models/racing_trajectory/src/safe_set.cpp:33-54,116-180,260-276 and
 *  src/trajectory_kd_tree.cpp:27-63).  CGAL's Orthogonal_k_neighbor_search (exact, sorted
 * ascending, Euclidean) is absent here; it is restated as a brute-force exact k-NN with a
 * stable (distance, index) order.  The coordinate-hash index recovery of the reference
 * (trajectory_kd_tree.cpp:38,60-62: duplicates of (s,e_y) resolve to the FIRST inserted
 * index) is reproduced through the canon[] table.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lmpc_oracle.h"

typedef struct {
  int n;          /* samples in the lap */
  double* xr;     /* x_repeat: 3n rows of 6: [x - L e0, x, x + L e0]  (safe_set.cpp:122-125) */
  double* J;      /* 3n: [J + n-1, J, J - n+1], J_j = n-1-j           (safe_set.cpp:121,128) */
  int* canon;     /* 3n: first index carrying the same (s, e_y) key   */
} lap_t;

struct orc_safe_set {
  int cap, count, head; /* circular buffer: newest at (head+count-1)%cap */
  lap_t* laps;
};

orc_safe_set* orc_ss_create(int max_lap_stored) {
  orc_safe_set* ss = (orc_safe_set*)calloc(1, sizeof *ss);
  ss->cap = max_lap_stored > 0 ? max_lap_stored : 1;
  ss->laps = (lap_t*)calloc((size_t)ss->cap, sizeof(lap_t));
  return ss;
}

static void lap_free(lap_t* l) { free(l->xr); free(l->J); free(l->canon); memset(l, 0, sizeof *l); }

void orc_ss_destroy(orc_safe_set* ss) {
  if (!ss) return;
  for (int i = 0; i < ss->cap; i++) lap_free(&ss->laps[i]);
  free(ss->laps);
  free(ss);
}

int orc_ss_num_laps(const orc_safe_set* ss) { return ss->count; }

typedef struct { double s, e; int idx; } key_t_;
static int key_cmp(const void* a, const void* b) {
  const key_t_* x = (const key_t_*)a; const key_t_* y = (const key_t_*)b;
  if (x->s != y->s) return x->s < y->s ? -1 : 1;
  if (x->e != y->e) return x->e < y->e ? -1 : 1;
  return x->idx - y->idx;
}

int orc_ss_add_lap(orc_safe_set* ss, int n, const double* x, const double* u, const double* k,
                   const double* t, double L) {
  (void)u; (void)k; (void)t; /* only the (unused) regression query reads them */
  if (n <= 0) return -1;
  lap_t lap; lap.n = n;
  lap.xr = (double*)malloc(sizeof(double) * 18 * (size_t)n);
  lap.J = (double*)malloc(sizeof(double) * 3 * (size_t)n);
  lap.canon = (int*)malloc(sizeof(int) * 3 * (size_t)n);
  for (int rep = 0; rep < 3; rep++) {
    for (int j = 0; j < n; j++) {
      double* dst = lap.xr + 6 * ((size_t)rep * n + j);
      memcpy(dst, x + 6 * (size_t)j, 6 * sizeof(double));
      dst[0] += (rep - 1) * L;
      const double Jj = (double)(n - 1 - j);                 /* linspace(n-1, 0, n) */
      lap.J[(size_t)rep * n + j] = Jj + (1 - rep) * (double)(n - 1);
    }
  }
  /* first-index-wins map for exactly equal keys */
  key_t_* keys = (key_t_*)malloc(sizeof(key_t_) * 3 * (size_t)n);
  for (int i = 0; i < 3 * n; i++) { keys[i].s = lap.xr[6 * i]; keys[i].e = lap.xr[6 * i + 1]; keys[i].idx = i; }
  qsort(keys, (size_t)(3 * n), sizeof(key_t_), key_cmp);
  for (int i = 0; i < 3 * n;) {
    int j = i;
    while (j < 3 * n && keys[j].s == keys[i].s && keys[j].e == keys[i].e) { lap.canon[keys[j].idx] = keys[i].idx; j++; }
    i = j;
  }
  free(keys);
  /* boost::circular_buffer::push_back (safe_set.cpp:150): overwrite the oldest when full */
  if (ss->count == ss->cap) {
    lap_free(&ss->laps[ss->head]);
    ss->laps[ss->head] = lap;
    ss->head = (ss->head + 1) % ss->cap;
  } else {
    ss->laps[(ss->head + ss->count) % ss->cap] = lap;
    ss->count++;
  }
  return 0;
}

static double* read_matrix(const char* path, int cols, int* rows_out) {
  FILE* f = fopen(path, "r");
  if (!f) return NULL;
  size_t cap = 1024, cnt = 0;
  double* buf = (double*)malloc(cap * sizeof(double));
  double v;
  while (fscanf(f, "%lf", &v) == 1) {
    if (cnt == cap) { cap *= 2; buf = (double*)realloc(buf, cap * sizeof(double)); }
    buf[cnt++] = v;
  }
  fclose(f);
  if (cnt == 0 || cnt % (size_t)cols) { free(buf); return NULL; }
  *rows_out = (int)(cnt / (size_t)cols);
  return buf;
}

/* SafeSetRecorder::load for one prefix (safe_set.cpp:260-276): <prefix>_{x,u,k,t}.txt */
int orc_ss_load(orc_safe_set* ss, const char* prefix, double L) {
  char path[4096];
  int nx = 0, nu = 0, nk = 0, nt = 0;
  snprintf(path, sizeof path, "%s_x.txt", prefix); double* x = read_matrix(path, 6, &nx);
  snprintf(path, sizeof path, "%s_u.txt", prefix); double* u = read_matrix(path, 2, &nu);
  snprintf(path, sizeof path, "%s_k.txt", prefix); double* k = read_matrix(path, 1, &nk);
  snprintf(path, sizeof path, "%s_t.txt", prefix); double* t = read_matrix(path, 1, &nt);
  int rc = -1;
  if (x && u && k && t && nx == nu && nx == nk && nx == nt) rc = orc_ss_add_lap(ss, nx, x, u, k, t, L);
  free(x); free(u); free(k); free(t);
  return rc;
}

typedef struct { double d2; int idx; } cand_t;
static int cand_cmp(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->d2 != y->d2) return x->d2 < y->d2 ? -1 : 1;
  return x->idx - y->idx;
}

int orc_ss_query(const orc_safe_set* ss, double qs, double qey, int max_total, int max_per_lap,
                 double* ss_x, double* ss_j) {
  int total = 0; /* may exceed max_total before the final truncation, like the reference */
  int written = 0;
  /* newest -> oldest while num_total < max_num_total (safe_set.cpp:164) */
  for (int li = ss->count - 1; li >= 0 && total < max_total; li--) {
    const lap_t* lap = &ss->laps[(ss->head + li) % ss->cap];
    const int m = 3 * lap->n;
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)m);
    for (int i = 0; i < m; i++) {
      const double ds = qs - lap->xr[6 * i], de = qey - lap->xr[6 * i + 1];
      c[i].d2 = ds * ds + de * de; c[i].idx = i;
    }
    qsort(c, (size_t)m, sizeof(cand_t), cand_cmp);
    const int take = max_per_lap < m ? max_per_lap : m;
    for (int j = 0; j < take; j++) {
      if (written < max_total) { /* horzcat then truncate to max_total (safe_set.cpp:175-178) */
        const int src = lap->canon[c[j].idx];
        memcpy(ss_x + 6 * (size_t)written, lap->xr + 6 * (size_t)src, 6 * sizeof(double));
        ss_j[written] = lap->J[src];
        written++;
      }
    }
    total += take;
    free(c);
  }
  return written;
}

int orc_ss_query_padded(const orc_safe_set* ss, double qs, double qey, int K, int per_lap,
                        double* ss_x, double* ss_cost) {
  const int cnt = orc_ss_query(ss, qs, qey, K, per_lap, ss_x, ss_cost);
  if (cnt == 0) return 0;
  /* pad with the last column (racing_mpc.cpp:263-272); truncation already applied */
  for (int j = cnt; j < K; j++) {
    memcpy(ss_x + 6 * (size_t)j, ss_x + 6 * (size_t)(cnt - 1), 6 * sizeof(double));
    ss_cost[j] = ss_cost[cnt - 1];
  }
  const double J0 = ss_cost[0]; /* ss_j - ss_j(:,0) (racing_mpc.cpp:280) */
  for (int j = 0; j < K; j++) ss_cost[j] -= J0;
  return cnt;
}
