// Compile test of the reference-side binding INTEGRATION.md shows (section 3): the reference's RacingMPCConfig /
// model configuration objects converted to the adapter's PODs, and the casadi::DMDict overload of solve().  Built
// against tests/stub_casadi (CasADi is not installable here) with -DLMPC_HAVE_CASADI; never executed on a GPU-less box
// beyond the constructor's refusal.  The struct below restates the fields of racing_mpc_config.hpp:37-82 the binding reads.
#include <cstdio>
#include "racing_mpc_b200.hpp"

struct RefRacingMPCConfig {   // lmpc::mpc::racing_mpc::RacingMPCConfig (racing_mpc_config.hpp:37-82)
  size_t N; bool learning; double margin;
  double q_contour, q_heading, q_vel, q_vy, q_vyaw, q_boundary;
  casadi::DM R, R_d, x_max, x_min, u_max, u_min, convex_hull_slack;
  size_t num_ss_pts, num_ss_pts_per_lap, max_lap_stored;
  bool record, load; std::string path_prefix; std::vector<std::string> load_path;
};

static lmpc_b200::RacingMPCConfig::SharedPtr to_b200(const RefRacingMPCConfig& c) {
  auto o = std::make_shared<lmpc_b200::RacingMPCConfig>();
  o->c.N = (int32_t)c.N; o->c.learning = c.learning; o->c.margin = c.margin;
  o->c.q_contour = c.q_contour; o->c.q_heading = c.q_heading; o->c.q_vel = c.q_vel; o->c.q_vy = c.q_vy; o->c.q_vyaw = c.q_vyaw;
  o->c.q_boundary = c.q_boundary;
  for (int i = 0; i < 4; i++) { o->c.R[i] = c.R.nz(i); o->c.R_d[i] = c.R_d.nz(i); }
  for (int i = 0; i < 6; i++) { o->c.x_max[i] = c.x_max(i); o->c.x_min[i] = c.x_min(i); o->c.convex_hull_slack[i] = c.convex_hull_slack(i); }
  for (int i = 0; i < 2; i++) { o->c.u_max[i] = c.u_max(i); o->c.u_min[i] = c.u_min(i); }
  o->c.num_ss_pts = (int32_t)c.num_ss_pts; o->c.num_ss_pts_per_lap = (int32_t)c.num_ss_pts_per_lap; o->c.max_lap_stored = (int32_t)c.max_lap_stored;
  o->record = c.record; o->path_prefix = c.path_prefix; o->load = c.load; o->load_path = c.load_path;
  return o;   // max_iter / tol left 0 -> defaults
}

int main() {
  RefRacingMPCConfig rc{};
  rc.N = 20; rc.learning = false; rc.margin = 0.1; rc.q_contour = 1; rc.q_heading = 1; rc.q_vel = 0.2; rc.q_vy = 1e-3; rc.q_vyaw = 1e-3; rc.q_boundary = 20;
  rc.R = casadi::DM(std::vector<double>{0.01, 0, 0, 0.01}); rc.R_d = rc.R;
  rc.x_max = casadi::DM(std::vector<double>{1e20, 1e20, 1e20, 6, 1, 3}); rc.x_min = casadi::DM(std::vector<double>{-1e20, -1e20, -1e20, 0.1, -1, -3});
  rc.u_max = casadi::DM(std::vector<double>{0.01, 0.33}); rc.u_min = casadi::DM(std::vector<double>{-0.01, -0.33});
  rc.convex_hull_slack = casadi::DM(std::vector<double>{20, 20, 2, 20, 20, 2});
  rc.num_ss_pts = 96; rc.num_ss_pts_per_lap = 32; rc.max_lap_stored = 3;
  lmpc_vehicle_params vp{};
  vp.mass = 2.2187; vp.moi = 0.02723; vp.wheel_base = 0.324; vp.cg_ratio = 0.5; vp.cg_height = 0.07; vp.fr = 0.012; vp.chassis_b = 0.281;
  vp.kb = 0.5; vp.air_density = 1.2; vp.frontal_area = 1; vp.mu = 0.9; vp.Bf = vp.Br = 5; vp.Cf = vp.Cr = 2.28;
  vp.Fd_max = 15; vp.Fb_max = -15; vp.Td = vp.Tb = 0.1; vp.max_steer = 0.314159; vp.max_steer_rate = 10;
  auto model = lmpc_b200::vehicle_model_factory::load_vehicle_model("single_track_planar_model", vp);
  try {
    auto mpc = std::make_shared<lmpc_b200::RacingMPC>(to_b200(rc), model);
    casadi::DMDict sol_in, sol_out; casadi::Dict stats;
    try { mpc->solve(sol_in, sol_out, stats); } catch (const std::out_of_range&) { std::printf("MISSING_KEY_THROWS\n"); }   // .at() on a missing key, as in the reference
  } catch (const std::runtime_error& e) {
    std::printf("CTOR_THROW %s\n", e.what());
    return 3;
  }
  return 0;
}
