// lmpc_qp_launch.h -- host-side launch table of the QP kernel instantiations.
//
// lmpc_qp_kernel<NW, KPL, N, RS> is instantiated in four translation units (lmpc_qp_tu{0..3}.cu) so that
// the library builds in parallel; lmpc_capi.cu reaches them through these two functions only.
#pragma once
#include <cuda_runtime.h>
#include "lmpc_qp_kernel.cuh"

// nf: compile-time horizon of the layout (20, 40; 60, 80 for the shipped long horizons) or 0 for the run-time layout.  Both return false when no
// instantiation matches (nw, kpl, nf).
bool lmpc_qp_set_smem(int nw, int kpl, int nf, size_t smem_bytes, cudaError_t* err);
bool lmpc_qp_launch(int nw, int kpl, int nf, int nblocks, size_t smem_bytes, cudaStream_t stream,
                    const LmpcQpParams& P, const LmpcQpBatch& a);

// per-translation-unit pieces (op 0 = set the dynamic shared-memory attribute, 1 = launch)
#define LMPC_QP_TU_DECL(n) \
  bool lmpc_qp_tu##n(int op, int nw, int kpl, int nf, int nblocks, size_t smem, cudaStream_t s, const LmpcQpParams* P, const LmpcQpBatch* a, cudaError_t* err)
LMPC_QP_TU_DECL(0);
LMPC_QP_TU_DECL(1);
LMPC_QP_TU_DECL(2);
LMPC_QP_TU_DECL(3);
LMPC_QP_TU_DECL(4);

#define LMPC_QP_CASE(NW_, KPL_, NF_, RS_)                                                                          \
  if (nw == NW_ && kpl == KPL_ && nf == NF_) {                                                                     \
    if (op == 0) { cudaError_t e = cudaFuncSetAttribute(lmpc_qp_kernel<NW_, KPL_, NF_, RS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (err) *err = e; } \
    else lmpc_qp_kernel<NW_, KPL_, NF_, RS_><<<nblocks, 32 * NW_, smem, s>>>(*P, *a);                               \
    return true;                                                                                                   \
  }
