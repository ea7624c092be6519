// QP kernel instantiations: one warp per instance, compile-time layouts for the long shipped horizons --
// N = 60 (barc_tracking_mpc.param.yaml n: 60; iac_car_lmpc.param.yaml n: 60, K = 96) and
// N = 80 (iac_car_tracking_mpc.param.yaml n: 80)
#include "lmpc_qp_launch.h"
LMPC_QP_TU_DECL(4) {
  LMPC_QP_CASE(1, 1, 60, 16) LMPC_QP_CASE(1, 3, 60, 16) LMPC_QP_CASE(1, 1, 80, 16)
  return false;
}
