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
#include <cstdio>
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

// The vehicle model surface of the reference (BaseVehicleModel virtuals that return casadi::Function objects with named
// inputs / outputs, base_vehicle_model.hpp; single_track_planar_model.cpp:370-417) as methods that evaluate on the GPU
// through the C ABI.  The handle is created on first use (the node builds the model before the MPC, racing_mpc_node.cpp:43-47).
//   discrete_dynamics()           {x, u, k, dt} -> {xip1}          (the Fx_ij / Fy_ij / Fz_ij diagnostic outputs are not provided)
//   discrete_dynamics_jacobian()  {x, u, k, dt} -> {A, B, g}
//   to_base_control / from_base_control / to_base_state / from_base_state   {x, u} -> {u_out} / {x_out}
struct SingleTrackPlanarModel {   // the "single_track_planar_model" entry of vehicle_model_factory.cpp:39-41
  lmpc_vehicle_params p{};
  int device = 0;
  typedef std::shared_ptr<SingleTrackPlanarModel> SharedPtr;
  SingleTrackPlanarModel() = default;
  explicit SingleTrackPlanarModel(const lmpc_vehicle_params& params, int dev = 0) : p(params), device(dev) {}
  ~SingleTrackPlanarModel() { if (h_) lmpc_destroy(h_); }
  SingleTrackPlanarModel(const SingleTrackPlanarModel&) = delete;
  SingleTrackPlanarModel& operator=(const SingleTrackPlanarModel&) = delete;
  size_t nx() const { return 6; }
  size_t nu() const { return 2; }

  // casadi::Function-style call: named inputs in, named outputs out (missing key -> std::out_of_range, as .at() does)
  MatrixDict discrete_dynamics(const MatrixDict& in) {
    const Matrix &x = in.at("x"), &u = in.at("u"), &k = in.at("k"), &dt = in.at("dt");
    Matrix xn(6, 1);
    check(lmpc_discrete_dynamics_batch(handle(), 1, x.data.data(), u.data.data(), k.data.data(), dt.data.data(), xn.data.data(), LMPC_MEM_HOST), "discrete_dynamics");
    return MatrixDict{{"xip1", xn}};
  }
  MatrixDict discrete_dynamics_jacobian(const MatrixDict& in) {
    const Matrix &x = in.at("x"), &u = in.at("u"), &k = in.at("k"), &dt = in.at("dt");
    Matrix A(6, 6), B(6, 2), g(6, 1);
    check(lmpc_linearise_batch(handle(), 1, x.data.data(), u.data.data(), k.data.data(), dt.data.data(), A.data.data(), B.data.data(), g.data.data(), nullptr, LMPC_MEM_HOST), "discrete_dynamics_jacobian");
    return MatrixDict{{"A", A}, {"B", B}, {"g", g}};
  }
  MatrixDict to_base_control(const MatrixDict& in) {     // single_track_planar_model.cpp:390-400
    const Matrix& u = in.at("u");
    Matrix ub(3, 1);
    check(lmpc_to_base_control_batch(handle(), 1, u.data.data(), ub.data.data(), LMPC_MEM_HOST), "to_base_control");
    return MatrixDict{{"u_out", ub}};
  }
  MatrixDict from_base_control(const MatrixDict& in) {   // :401-407
    const Matrix& ub = in.at("u");
    Matrix u(2, 1);
    check(lmpc_from_base_control_batch(handle(), 1, ub.data.data(), u.data.data(), LMPC_MEM_HOST), "from_base_control");
    return MatrixDict{{"u_out", u}};
  }
  MatrixDict to_base_state(const MatrixDict& in) const { return MatrixDict{{"x_out", in.at("x")}}; }     // identity, :411-412
  MatrixDict from_base_state(const MatrixDict& in) const { return MatrixDict{{"x_out", in.at("x")}}; }   // identity, :413-414
  // batched forms for Monte-Carlo callers (instance-major arrays, host or device memory)
  int discrete_dynamics_batch(int n, const double* x, const double* u, const double* k, const double* dt, double* xip1, int memspace) {
    return lmpc_discrete_dynamics_batch(handle(), n, x, u, k, dt, xip1, memspace);
  }
  int discrete_dynamics_jacobian_batch(int n, const double* x, const double* u, const double* k, const double* dt, double* A, double* B, double* g, int memspace) {
    return lmpc_linearise_batch(handle(), n, x, u, k, dt, A, B, g, nullptr, memspace);
  }

 private:
  lmpc_handle* handle() {
    if (!h_) {
      const int rc = lmpc_model_create(&p, device, &h_);
      if (rc != LMPC_OK) { h_ = nullptr; throw std::runtime_error(std::string("lmpc_model_create: ") + lmpc_status_string(rc)); }
    }
    return h_;
  }
  void check(int rc, const char* what) {
    if (rc != LMPC_OK) throw std::runtime_error(std::string(what) + ": " + lmpc_status_string(rc) + " " + lmpc_last_error(h_));
  }
  lmpc_handle* h_ = nullptr;
};

namespace vehicle_model_factory {
// vehicle_model_factory::load_vehicle_model(model_name, node) (vehicle_model_factory.cpp:31-50): the parameters come as a
// POD instead of a ROS node.  Only "single_track_planar_model" has a B200 counterpart; the reference logs FATAL and returns
// nullptr for an unknown name, and so does this (the other two models of the reference are outside the hot path).
inline SingleTrackPlanarModel::SharedPtr load_vehicle_model(const std::string& model_name, const lmpc_vehicle_params& params, int device = 0) {
  if (model_name == "single_track_planar_model") return std::make_shared<SingleTrackPlanarModel>(params, device);
  std::fprintf(stderr, "Vehicle model %s cannot be found.\n", model_name.c_str());
  return nullptr;
}
}  // namespace vehicle_model_factory

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
    const int N = config_->c.N;
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
    const int Kc = config_->c.num_ss_pts > 0 ? config_->c.num_ss_pts : 1;
    Matrix X(6, N), U(2, N - 1), dU(2, N - 1), lam(Kc, 1), ssx(6, Kc), ssj(1, Kc);
    double cost = 0.0, defect = 0.0;
    int32_t status = 0, iters = 0, sqp_iters = 0, found = 0;
    lmpc_batch_out bo{};
    bo.X_optm = X.data.data(); bo.U_optm = U.data.data(); bo.dU_optm = dU.data.data();
    bo.convex_combi_optm = lam.data.data(); bo.ss_x = ssx.data.data(); bo.ss_j = ssj.data.data();
    bo.cost = &cost; bo.status = &status; bo.iters = &iters;
    // full_dynamics: IPOPT in the reference runs with max_iter 1000 and error_on_fail (racing_mpc.cpp:67-84): a run that does
    // not converge throws, X_optm stays absent, solved() stays false and the node retries on the next tick.  Here the SQP's
    // own convergence test decides (status LMPC_SQP_MAX_ITER when it did not pass within max_sqp_iter).
    if (full_dynamics_) check(lmpc_solve_sqp_batch(h_, 1, &bi, &bo, max_sqp_iter, sqp_tol, &sqp_iters, &defect, LMPC_MEM_HOST), "solve_sqp_batch");
    else check(lmpc_solve_batch(h_, 1, &bi, &bo, LMPC_MEM_HOST), "solve_batch");
    // out["ss_x"] / out["ss_j"]: the UNPADDED query result, in every mode (racing_mpc.cpp:249-257)
    check(lmpc_safe_set_tick_count(h_, &found), "safe_set_tick_count");
    if (!config_->c.learning && found > 0) {   // tracking mode: the QP does not use the columns; the reference still returns them
      const double q[2] = {X_ref(0, N - 1) /* aligned below */, X_ref(1, N - 1)};
      double qa[2] = {align_abscissa(q[0], x_ic.data[0], L.data[0]), q[1]};
      check(lmpc_safe_set_query_batch(h_, 1, qa, Kc, config_->c.num_ss_pts_per_lap, ssx.data.data(), ssj.data.data(), nullptr, LMPC_MEM_HOST), "safe_set_query");
    }
    Matrix sx(6, found), sj(1, found);
    for (int q = 0; q < 6 * found; q++) sx.data[q] = ssx.data[q];
    for (int q = 0; q < found; q++) sj.data[q] = ssj.data[q];
    out["ss_x"] = sx; out["ss_j"] = sj;
    stats["iter_count"] = full_dynamics_ ? sqp_iters : iters;
    stats["status"] = status;
    stats["cost"] = cost;
    if (full_dynamics_) stats["defect"] = defect;
    if (status == LMPC_SOLVED || status == LMPC_SOLVED_INACCURATE) {                // racing_mpc.cpp:345-352
      solved_ = true;
      out["X_optm"] = X; out["U_optm"] = U; out["dU_optm"] = dU;
      if (config_->c.learning) out["convex_combi_optm"] = lam;
      last_U_ = U; have_last_ = true;
    }   // on failure the keys are simply absent (racing_mpc.cpp:358-371)
  }

  // RacingMPC::create_warm_start (racing_mpc.cpp:374-430): a reference trajectory from a path (P0 2 x N, Yaws 1 x N,
  // Radii 1 x N) and a speed ramp.  X_ref: rows 0,1 <- P0, row 2 <- Yaws, v_x <- linspace(current_vel, target_vel, N),
  // omega <- v_x / Radii.  Controls: force from Newton's second law over each segment, steering by pure pursuit.  As
  // written the reference indexes the BASE control rows (Fd, Fb, steer = 0, 1, 2) into a matrix with model->nu() = 2 rows,
  // which CasADi rejects (the function has no caller in the reference); here the base control (3 rows) is formed as written
  // and mapped to the model's control with from_base_control (:401-407).  Same exceptions (:385-396).
  void create_warm_start(const MatrixDict& in, MatrixDict& out) {
    const Matrix &P0 = in.at("P0"), &Yaws = in.at("Yaws"), &Radii = in.at("Radii"), &cv = in.at("current_vel"), &tv = in.at("target_vel");
    const int N = config_->c.N;
    if (P0.size2() != N) throw std::length_error("create_warm_start: P0 dimension does not match MPC dimension.");
    if (Yaws.size2() != N) throw std::length_error("create_warm_start: Yaws dimension does not match MPC dimension.");
    if (cv.data[0] <= 0.0) throw std::range_error("Current velocity cannot be smaller than or equal to zero.");
    if (tv.data[0] <= 0.0) throw std::range_error("Target velocity cannot be smaller than or equal to zero.");
    Matrix X_ref(6, N), U_ref(2, N - 1), T_ref(N - 1, 1);
    for (int i = 0; i < N; i++) {
      X_ref(0, i) = P0(0, i); X_ref(1, i) = P0(1, i); X_ref(2, i) = Yaws.data[i];
      X_ref(3, i) = cv.data[0] + (tv.data[0] - cv.data[0]) * (N > 1 ? (double)i / (double)(N - 1) : 0.0);   // DM::linspace
      X_ref(5, i) = X_ref(3, i) / Radii.data[i];
    }
    for (int i = 0; i < N - 1; i++) {
      const double v0 = X_ref(3, i), v1 = X_ref(3, i + 1);
      const double dd = std::hypot(P0(0, i) - P0(0, i + 1), P0(1, i) - P0(1, i + 1));
      const double a = (v1 * v1 - v0 * v0) / (2.0 * dd);
      const double f = model_->p.mass * a;
      const double ub[3] = {f > 0.0 ? f : 0.0, f > 0.0 ? 0.0 : f, std::atan(model_->p.wheel_base / Radii.data[i])};
      U_ref(0, i) = std::fabs(ub[0]) > std::fabs(ub[1]) ? ub[0] : ub[1];   // from_base_control
      U_ref(1, i) = ub[2];
      T_ref.data[i] = dd / v0;
    }
    out["X_ref"] = X_ref;
    out["U_ref"] = U_ref;
  }

  // SQP options of the full_dynamics variant (the counterparts of IPOPT's max_iter 1000 / tol, racing_mpc.cpp:67-84)
  int max_sqp_iter = 100;
  double sqp_tol = 1e-9;

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
  static double align_abscissa(double s1, double s2, double L) {   // lmpc_utils/utils.hpp:35-41
    const double k = std::fabs(s2 - s1) + 0.5 * L;
    return s1 + (k - std::fmod(k, L)) * ((s2 > s1) - (s2 < s1));
  }
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
