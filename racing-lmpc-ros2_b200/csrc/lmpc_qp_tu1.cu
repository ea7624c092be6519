// QP kernel instantiations: one warp per instance, compile-time layout for N = 40 (IAC tracking)
#include "lmpc_qp_launch.h"
LMPC_QP_TU_DECL(1) {
  LMPC_QP_CASE(1, 1, 40, 16) LMPC_QP_CASE(1, 2, 40, 16) LMPC_QP_CASE(1, 3, 40, 16) LMPC_QP_CASE(1, 4, 40, 16)
  return false;
}
