// racing_mpc_b200.hpp -- header-only C++ adapter over the C ABI (include/lmpc_b200.h) with the
// reference's class surface, so that RacingMPCNode can link against it instead of the CasADi/OSQP
// implementation:
//
//   lmpc::mpc::racing_mpc::RacingMPC           racing_mpc.hpp:37-108   -> lmpc_b200::RacingMPC
//     RacingMPC(config, model, full_dynamics)   racing_mpc.hpp:46-49
//     void solve(in, out, stats)                racing_mpc.hpp:52, racing_mpc.cpp:209-372
//     const bool & solved() const               racing_mpc.hpp:58
//   lmpc::vehicle_model::racing_trajectory::SafeSetManager::add_lap   safe_set.hpp:119-121
//
// Same key strings, same "X_optm absent on failure" contract, stats["iter_count"] filled.  The matrix
// type is a minimal dense column-major fp64 matrix (the layout of casadi::DM); with
// -DLMPC_HAVE_CASADI an overload taking casadi::DMDict is compiled as well (CasADi is not installable in
// the build environment of this repo, so that overload is compile-gated and untested here).
//
// One instance per call is the reference's usage; solve_batch() exposes the batched entry for
// Monte-Carlo callers.  Not thread-safe (the reference serialises solve() under a mutex,
// racing_mpc_node.cpp:158).
#pragma once

#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "lmpc_b200.h"

#ifdef LMPC_HAVE_CASADI
#include <casadi/casadi.hpp>
#endif

namespace lmpc_b200 {

// dense column-major matrix == memory layout of a dense casadi::DM
struct Matrix {
  int rows = 0, cols = 0;
  std::vector<double> data;
  Matrix() = default;
  Matrix(int r, int c, double v = 0.0) : rows(r), cols(c), data((size_t)r * c, v) {}
  explicit Matrix(double scalar) : rows(1), cols(1), data(1, scalar) {}
  double& operator()(int r, int c) { return data[(size_t)c * rows + r]; }
  double operator()(int r, int c) const { return data[(size_t)c * rows + r]; }
  int size1() const { return rows; }
  int size2() const { return cols; }
};
using MatrixDict = std::map<std::string, Matrix>;
using StatsDict = std::map<std::string, double>;

struct RacingMPCConfig {   // racing_mpc_config.hpp:37-82 (fields read by the solve path)
  lmpc_mpc_config c{};
  bool record = false;                  // :76-77  SafeSetRecorder(to_file = record, file_prefix = path_prefix)
  std::string path_prefix;
  bool load = false;                    // :79-80  laps loaded on the first solve (racing_mpc.cpp:240-243)
  std::vector<std::string> load_path;
  typedef std::shared_ptr<RacingMPCConfig> SharedPtr;
};

struct SingleTrackPlanarModel {   // the "single_track_planar_model" entry of vehicle_model_factory.cpp:39-41
  lmpc_vehicle_params p{};
  typedef std::shared_ptr<SingleTrackPlanarModel> SharedPtr;
  size_t nx() const { return 6; }
  size_t nu() const { return 2; }
};

class RacingMPC {
 public:
  typedef std::shared_ptr<RacingMPC> SharedPtr;

  RacingMPC(RacingMPCConfig::SharedPtr config, SingleTrackPlanarModel::SharedPtr model, const bool& full_dynamics = false,
            int device = 0, int max_batch = 1)
      : config_(config), model_(model), max_batch_(max_batch), full_dynamics_(full_dynamics) {
    // full_dynamics: the nonlinear-equality problem the reference gives to IPOPT (racing_mpc.cpp:67-84) is solved by
    // SQP on the tick's kernels (lmpc_solve_sqp_batch)
    const int rc = lmpc_create(&config_->c, &model_->p, device, max_batch, &h_);
    if (rc != LMPC_OK) throw std::runtime_error(std::string("lmpc_create: ") + lmpc_status_string(rc));
    check(lmpc_recorder_config(h_, config_->record ? 1 : 0, config_->path_prefix.c_str()), "recorder_config");   // racing_mpc.cpp:59-61
  }
  ~RacingMPC() { if (h_) lmpc_destroy(h_); }
  RacingMPC(const RacingMPC&) = delete;
  RacingMPC& operator=(const RacingMPC&) = delete;

  const RacingMPCConfig& get_config() const { return *config_; }
  SingleTrackPlanarModel& get_model() { return *model_; }
  const bool& solved() const { return solved_; }

  // SafeSetManager::add_lap(x, u, k, t, total_length): x is 6 x n, u 2 x n, k and t 1 x n
  void add_lap(const Matrix& x, const Matrix& u, const Matrix& k, const Matrix& t, double total_length) {
    check(lmpc_safe_set_add_lap(h_, x.cols, x.data.data(), u.data.data(), k.data.data(), t.data.data(), total_length), "add_lap");
  }
  void load_laps(const std::vector<std::string>& prefixes, double total_length) {   // SafeSetRecorder::load
    for (const auto& p : prefixes) check(lmpc_safe_set_load(h_, p.c_str(), total_length), "load");
  }

  // RacingMPC::solve: keys as racing_mpc.cpp:215-228 (inputs) and :256-257,347-352 (outputs)
  void solve(const MatrixDict& in, MatrixDict& out, StatsDict& stats) {
    const int N = config_->c.N, K = config_->c.num_ss_pts;
    const Matrix& L = in.at("total_length");
    const Matrix& x_ic = in.at("x_ic");
    const Matrix& u_ic = in.at("u_ic");
    const Matrix& t_ic = in.at("t_ic");
    const Matrix& X_ref = in.at("X_ref");
    const Matrix& U_ref = in.at("U_ref");
    const Matrix& bl = in.at("bound_left");
    const Matrix& br = in.at("bound_right");
    const Matrix& kap = in.at("curvatures");
    const Matrix& vref = in.at("vel_ref");
    const Matrix* T = nullptr;
    const Matrix* Uw = nullptr;
    if (in.count("X_optm_ref")) {                          // racing_mpc.cpp:293-305
      T = &in.at("T_optm_ref");
      Uw = &in.at("U_optm_ref");
      (void)in.at("dU_optm_ref");
    } else {
      if (!solved_) throw std::runtime_error("No warm start given and no previous solution found.");   // :312-314
      T = &in.at("T_ref");
    }
    if (!ss_loaded_ && config_->load) { load_laps(config_->load_path, L.data[0]); ss_loaded_ = true; }   // racing_mpc.cpp:240-243
    // add current state to safe set (racing_mpc.cpp:245-246): lap segmentation, completed laps join the device slab
    check(lmpc_recorder_step(h_, x_ic.data.data(), u_ic.data.data(), kap.data[0], t_ic.data[0], L.data[0], nullptr), "recorder_step");
    lmpc_batch_in bi{};
    bi.x_ic = x_ic.data.data(); bi.u_ic = u_ic.data.data(); bi.X_ref = X_ref.data.data(); bi.U_ref = U_ref.data.data();
    bi.T_ref = T->data.data(); bi.bound_left = bl.data.data(); bi.bound_right = br.data.data();
    bi.curvatures = kap.data.data(); bi.vel_ref = vref.data.data(); bi.total_length = L.data.data();
    bi.U_warm = Uw ? Uw->data.data() : (have_last_ ? last_U_.data.data() : nullptr);
    Matrix X(6, N), U(2, N - 1), dU(2, N - 1), lam(K > 0 ? K : 1, 1), ssx(6, K > 0 ? K : 1), ssj(1, K > 0 ? K : 1);
    double cost = 0.0;
    int32_t status = 0, iters = 0;
    lmpc_batch_out bo{};
    bo.X_optm = X.data.data(); bo.U_optm = U.data.data(); bo.dU_optm = dU.data.data();
    bo.convex_combi_optm = lam.data.data(); bo.ss_x = ssx.data.data(); bo.ss_j = ssj.data.data();
    bo.cost = &cost; bo.status = &status; bo.iters = &iters;
    if (full_dynamics_) check(lmpc_solve_sqp_batch(h_, 1, &bi, &bo, 30, 1e-9, nullptr, nullptr, LMPC_MEM_HOST), "solve_sqp_batch");
    else check(lmpc_solve_batch(h_, 1, &bi, &bo, LMPC_MEM_HOST), "solve_batch");
    if (config_->c.learning) { out["ss_x"] = ssx; out["ss_j"] = ssj; }            // racing_mpc.cpp:256-257
    stats["iter_count"] = iters;
    stats["status"] = status;
    stats["cost"] = cost;
    if (status == LMPC_SOLVED || status == LMPC_SOLVED_INACCURATE) {                                                   // racing_mpc.cpp:345-352
      solved_ = true;
      out["X_optm"] = X; out["U_optm"] = U; out["dU_optm"] = dU;
      if (config_->c.learning) out["convex_combi_optm"] = lam;
      last_U_ = U; have_last_ = true;
    }   // on failure the keys are simply absent (racing_mpc.cpp:358-371)
  }

  // Batched entry for Monte-Carlo callers: raw instance-major arrays, host or device memory.
  int solve_batch(int B, const lmpc_batch_in& in, const lmpc_batch_out& out, int memspace) {
    return lmpc_solve_batch(h_, B, &in, &out, memspace);
  }

#ifdef LMPC_HAVE_CASADI
  void solve(const casadi::DMDict& in, casadi::DMDict& out, casadi::Dict& stats) {
    MatrixDict min, mout; StatsDict s;
    for (const auto& kv : in) {
      Matrix m((int)kv.second.size1(), (int)kv.second.size2());
      m.data = casadi::DM::densify(kv.second).get_elements();
      min[kv.first] = m;
    }
    solve(min, mout, s);
    for (const auto& kv : mout) out[kv.first] = casadi::DM::reshape(casadi::DM(kv.second.data), kv.second.rows, kv.second.cols);
    for (const auto& kv : s) stats[kv.first] = kv.second;
  }
#endif

  lmpc_handle* handle() { return h_; }

 private:
  void check(int rc, const char* what) {
    if (rc != LMPC_OK) throw std::runtime_error(std::string(what) + ": " + lmpc_status_string(rc) + " " + lmpc_last_error(h_));
  }
  RacingMPCConfig::SharedPtr config_;
  SingleTrackPlanarModel::SharedPtr model_;
  lmpc_handle* h_ = nullptr;
  int max_batch_ = 1;
  bool full_dynamics_ = false;
  bool ss_loaded_ = false;
  bool solved_ = false;
  bool have_last_ = false;
  Matrix last_U_;
};

}  // namespace lmpc_b200
