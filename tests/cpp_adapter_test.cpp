// Compile/link check of the header-only C++ adapter (racing_mpc_b200.hpp) against liblmpc_b200.so.
// On a box without a GPU construction must throw (no CPU fallback); with a GPU (GPU test) it solves
// one tick read from a raw file produced by the Python side (argv[1]); argv[2] = "full" selects full_dynamics.
#include <cstdio>
#include <cstring>
#include <fstream>
#include "racing_mpc_b200.hpp"
using namespace lmpc_b200;
int main(int argc, char** argv) {
  auto cfg = std::make_shared<RacingMPCConfig>();
  auto mdl = std::make_shared<SingleTrackPlanarModel>();
  if (argc < 2) { std::printf("usage\n"); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  f.read((char*)&cfg->c, sizeof cfg->c);
  f.read((char*)&mdl->p, sizeof mdl->p);
  const bool full = argc > 2 && !std::strcmp(argv[2], "full");   // full_dynamics = true: the SQP path (racing_mpc.cpp:67-84)
  try {
    RacingMPC mpc(cfg, mdl, full, 0, 1);
    const int N = cfg->c.N;
    int nl = 0; f.read((char*)&nl, sizeof nl);
    for (int l = 0; l < nl; l++) {
      int n = 0; double L = 0; f.read((char*)&n, sizeof n); f.read((char*)&L, sizeof L);
      Matrix x(6, n), u(2, n), k(1, n), t(1, n);
      f.read((char*)x.data.data(), 8 * 6 * n); f.read((char*)u.data.data(), 8 * 2 * n);
      f.read((char*)k.data.data(), 8 * n); f.read((char*)t.data.data(), 8 * n);
      mpc.add_lap(x, u, k, t, L);
    }
    MatrixDict in, out; StatsDict st;
    auto rd = [&](const char* key, int r, int c) { Matrix m(r, c); f.read((char*)m.data.data(), 8 * r * c); in[key] = m; };
    rd("total_length", 1, 1); rd("x_ic", 6, 1); rd("u_ic", 2, 1); rd("t_ic", 1, 1); rd("X_ref", 6, N); rd("U_ref", 2, N - 1);
    rd("T_ref", 1, N - 1); rd("bound_left", 1, N); rd("bound_right", 1, N); rd("curvatures", 1, N); rd("vel_ref", 1, N);
    bool threw = false;
    try { mpc.solve(in, out, st); } catch (const std::runtime_error&) { threw = true; }   // no warm start, not solved yet
    if (!threw) { std::printf("FAIL: expected runtime_error without warm start\n"); return 1; }
    in["X_optm_ref"] = in["X_ref"]; in["U_optm_ref"] = in["U_ref"]; in["T_optm_ref"] = in["T_ref"]; in["dU_optm_ref"] = Matrix(2, N - 1);
    mpc.solve(in, out, st);
    if (!out.count("X_optm") || !mpc.solved()) { std::printf("FAIL: not solved status=%g\n", st["status"]); return 1; }
    std::printf("OK iters=%g cost=%.12g x1=%.12g\n", st["iter_count"], st["cost"], out["X_optm"](3, N - 1));
    in.erase("X_optm_ref");   // second tick: warm start from the previous solution (T_ref path)
    mpc.solve(in, out, st);
    std::printf("OK2 iters=%g\n", st["iter_count"]);
    return 0;
  } catch (const std::runtime_error& e) {
    std::printf("CTOR_THROW %s\n", e.what());
    return 3;
  }
}
