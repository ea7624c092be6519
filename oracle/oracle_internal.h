/* oracle_internal.h -- CPU ORACLE internals (test infrastructure only; see lmpc_oracle.h). */
#ifndef ORACLE_INTERNAL_H_
#define ORACLE_INTERNAL_H_
#include "lmpc_oracle.h"

#define ORC_NMAX 128   /* horizon cap of the oracle */
#define ORC_KMAX 256   /* safe-set columns cap */

/* Everything RacingMPC::solve hands to the QP after its own preprocessing
 * (racing_mpc.cpp:215-340): aligned reference, linearisation, safe-set columns. */
typedef struct orc_prob {
  int N, K, learning, soft_boundary, hull_slack;
  double x_ic[6], u_ic[2];
  double Xref[6 * ORC_NMAX];                 /* abscissa-aligned */
  double A[36 * ORC_NMAX], B[12 * ORC_NMAX], g[6 * ORC_NMAX], T[ORC_NMAX];
  double bl[ORC_NMAX], br[ORC_NMAX], vref[ORC_NMAX];
  double ssx[6 * ORC_KMAX], ssc[ORC_KMAX];
  int ss_count;
  double margin;                             /* config margin + chassis b / 2 (racing_mpc.cpp:531) */
  double ulo[2], uhi[2];                     /* merged u_min/u_max and actuator box */
  double dlo[2], dhi[2];                     /* actuator rate box (single_track_planar_model.cpp:146-151) */
} orc_prob;

/* returns ORC_OK / ORC_NO_SAFE_SET / ORC_INFEASIBLE_IC */
int orc_build_prob(const orc_vehicle* v, const orc_config* c, const orc_safe_set* ss,
                   const orc_step_in* in, orc_prob* p);
/* objective of the reference QP at a candidate (racing_mpc.cpp:442-543) */
double orc_eval_cost(const orc_config* c, const orc_prob* p, const double* X, const double* U,
                     const double* dU, double sigma_b, const double* lambda, const double* sigma_h);
#endif
