// QP kernel instantiations: one warp per instance, run-time layout (any horizon / row count)
#include "lmpc_qp_launch.h"
LMPC_QP_TU_DECL(2) {
  LMPC_QP_CASE(1, 1, 0, 0) LMPC_QP_CASE(1, 2, 0, 0) LMPC_QP_CASE(1, 3, 0, 0) LMPC_QP_CASE(1, 4, 0, 0)
  return false;
}
