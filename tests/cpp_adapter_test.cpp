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
    std::printf("SS cols=%d found_rows=%d\n", out.at("ss_x").size2(), out.at("ss_x").size1());
    in.erase("X_optm_ref");   // second tick: warm start from the previous solution (T_ref path)
    mpc.solve(in, out, st);
    std::printf("OK2 iters=%g\n", st["iter_count"]);
    // ---- the model surface (vehicle_model_factory + BaseVehicleModel functions) and create_warm_start
    if (vehicle_model_factory::load_vehicle_model("double_track_planar_model", mdl->p) != nullptr) { std::printf("FAIL: factory accepted an unknown model\n"); return 1; }
    auto m2 = vehicle_model_factory::load_vehicle_model("single_track_planar_model", mdl->p);
    MatrixDict mi;
    Matrix x(6, 1), u(2, 1);
    for (int c = 0; c < 6; c++) x.data[c] = in["X_ref"](c, 0);
    u.data[0] = in["U_ref"](0, 0); u.data[1] = in["U_ref"](1, 0);
    mi["x"] = x; mi["u"] = u; mi["k"] = Matrix(in["curvatures"].data[0]); mi["dt"] = Matrix(in["T_ref"].data[0]);
    const auto dd = m2->discrete_dynamics(mi);
    const auto jj = m2->discrete_dynamics_jacobian(mi);
    double chk = 0.0;   // g = xip1 - A x - B u (single_track_planar_model.cpp:379)
    for (int r = 0; r < 6; r++) {
      double a = jj.at("g")(r, 0);
      for (int c = 0; c < 6; c++) a += jj.at("A")(r, c) * x.data[c];
      for (int c = 0; c < 2; c++) a += jj.at("B")(r, c) * u.data[c];
      chk = std::fmax(chk, std::fabs(a - dd.at("xip1")(r, 0)));
    }
    const auto ub = m2->to_base_control(mi);
    MatrixDict bi2; bi2["x"] = x; bi2["u"] = ub.at("u_out");
    const auto ud = m2->from_base_control(bi2);
    std::printf("MODEL xip1_3=%.12g g_resid=%.3e fd=%.12g fb=%.12g back=%.12g same_state=%d\n", dd.at("xip1")(3, 0), chk, ub.at("u_out")(0, 0),
                ub.at("u_out")(1, 0), ud.at("u_out")(0, 0), (int)(m2->to_base_state(mi).at("x_out")(3, 0) == x.data[3]));
    MatrixDict wi, wo;
    Matrix P0(2, N), Yaws(1, N), Radii(1, N);
    for (int i = 0; i < N; i++) { P0(0, i) = 0.5 * i; P0(1, i) = 0.01 * i * i; Yaws.data[i] = 0.04 * i; Radii.data[i] = 12.0; }
    wi["P0"] = P0; wi["Yaws"] = Yaws; wi["Radii"] = Radii; wi["current_vel"] = Matrix(1.0); wi["target_vel"] = Matrix(2.0);
    mpc.create_warm_start(wi, wo);
    std::printf("WARM vx_last=%.12g omega1=%.12g u0=%.12g steer=%.12g\n", wo.at("X_ref")(3, N - 1), wo.at("X_ref")(5, 1), wo.at("U_ref")(0, 0), wo.at("U_ref")(1, 0));
    bool le = false, re = false;
    wi["P0"] = Matrix(2, N + 1);
    try { mpc.create_warm_start(wi, wo); } catch (const std::length_error&) { le = true; }
    wi["P0"] = P0; wi["current_vel"] = Matrix(0.0);
    try { mpc.create_warm_start(wi, wo); } catch (const std::range_error&) { re = true; }
    std::printf("WARM_ERRORS %d %d\n", (int)le, (int)re);
    return 0;
  } catch (const std::runtime_error& e) {
    std::printf("CTOR_THROW %s\n", e.what());
    return 3;
  }
}
