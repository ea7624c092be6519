// QP kernel instantiations: one warp per instance, compile-time layout for N = 20 (BARC LMPC / tracking)
#include "lmpc_qp_launch.h"
LMPC_QP_TU_DECL(0) {
  LMPC_QP_CASE(1, 1, 20, 16) LMPC_QP_CASE(1, 2, 20, 16) LMPC_QP_CASE(1, 3, 20, 16) LMPC_QP_CASE(1, 4, 20, 16)
  return false;
}
