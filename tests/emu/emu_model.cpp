// Test-only: compiles the product's model header for the host (LMPC_EMULATE) so its analytic
// Jacobians can be checked against the oracle / sympy without a GPU.  Never linked into the product.
#define LMPC_EMULATE 1
#include "../../racing-lmpc-ros2_b200/csrc/lmpc_model.cuh"
#include "../../racing-lmpc-ros2_b200/csrc/lmpc_host_params.h"
int g_lmpc_emu_reverse = 0;
extern "C" void emu_linearise(const lmpc_vehicle_params* v, const double* x, const double* u, double kappa, double dt,
                              double* A, double* B, double* g, double* xn) {
  LmpcModel P = lmpc_make_model(*v);
  lmpc_linearise(P, x, u, kappa, dt, A, B, g, xn);
}
extern "C" void emu_step(const lmpc_vehicle_params* v, const double* x, const double* u, double kappa, double dt, double* xn) {
  LmpcModel P = lmpc_make_model(*v);
  lmpc_step(P, x, u, kappa, dt, xn);
}
