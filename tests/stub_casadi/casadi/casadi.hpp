// Minimal stand-in for <casadi/casadi.hpp> (test infrastructure): just enough of casadi::DM / DMDict / Dict for the
// compile test of the adapter's casadi::DMDict overload (racing_mpc_b200.hpp, -DLMPC_HAVE_CASADI) and of the
// INTEGRATION.md binding snippet.  CasADi itself is not installable in this environment.  Dense column-major storage,
// the member functions the adapter and the snippet call, nothing else.
#pragma once
#include <map>
#include <string>
#include <vector>
namespace casadi {
typedef long long casadi_int;
class DM {
 public:
  DM() : r_(0), c_(0) {}
  DM(double v) : r_(1), c_(1), d_(1, v) {}
  explicit DM(const std::vector<double>& v) : r_((casadi_int)v.size()), c_(1), d_(v) {}
  static DM zeros(casadi_int r, casadi_int c) { DM m; m.r_ = r; m.c_ = c; m.d_.assign((size_t)(r * c), 0.0); return m; }
  static DM densify(const DM& m) { return m; }
  static DM reshape(const DM& m, casadi_int r, casadi_int c) { DM o = m; o.r_ = r; o.c_ = c; return o; }
  casadi_int size1() const { return r_; }
  casadi_int size2() const { return c_; }
  std::vector<double> get_elements() const { return d_; }
  double nz(casadi_int k) const { return d_[(size_t)k]; }
  double operator()(casadi_int k) const { return d_[(size_t)k]; }
  explicit operator double() const { return d_.empty() ? 0.0 : d_[0]; }
  std::vector<double>& data() { return d_; }
 private:
  casadi_int r_, c_;
  std::vector<double> d_;
};
typedef std::map<std::string, DM> DMDict;
class GenericType {
 public:
  GenericType() : v_(0.0) {}
  GenericType(double v) : v_(v) {}
  operator double() const { return v_; }
 private:
  double v_;
};
typedef std::map<std::string, GenericType> Dict;
}  // namespace casadi
