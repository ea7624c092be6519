// QP kernel instantiations: two / four warps per instance (selectable with LMPC_WARPS_PER_INSTANCE)
#include "lmpc_qp_launch.h"
LMPC_QP_TU_DECL(3) {
  LMPC_QP_CASE(2, 1, 20, 16) LMPC_QP_CASE(2, 2, 20, 16) LMPC_QP_CASE(2, 1, 0, 0) LMPC_QP_CASE(2, 2, 0, 0) LMPC_QP_CASE(4, 1, 0, 0)
  return false;
}

bool lmpc_qp_set_smem(int nw, int kpl, int nf, size_t smem, cudaError_t* err) {
  if (err) *err = cudaSuccess;
  return lmpc_qp_tu0(0, nw, kpl, nf, 0, smem, nullptr, nullptr, nullptr, err) || lmpc_qp_tu1(0, nw, kpl, nf, 0, smem, nullptr, nullptr, nullptr, err) ||
         lmpc_qp_tu2(0, nw, kpl, nf, 0, smem, nullptr, nullptr, nullptr, err) || lmpc_qp_tu3(0, nw, kpl, nf, 0, smem, nullptr, nullptr, nullptr, err) ||
         lmpc_qp_tu4(0, nw, kpl, nf, 0, smem, nullptr, nullptr, nullptr, err);
}
bool lmpc_qp_launch(int nw, int kpl, int nf, int nblocks, size_t smem, cudaStream_t s, const LmpcQpParams& P, const LmpcQpBatch& a) {
  return lmpc_qp_tu0(1, nw, kpl, nf, nblocks, smem, s, &P, &a, nullptr) || lmpc_qp_tu1(1, nw, kpl, nf, nblocks, smem, s, &P, &a, nullptr) ||
         lmpc_qp_tu2(1, nw, kpl, nf, nblocks, smem, s, &P, &a, nullptr) || lmpc_qp_tu3(1, nw, kpl, nf, nblocks, smem, s, &P, &a, nullptr) ||
         lmpc_qp_tu4(1, nw, kpl, nf, nblocks, smem, s, &P, &a, nullptr);
}
