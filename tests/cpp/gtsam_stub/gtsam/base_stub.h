// Signature-only stand-ins for the handful of GTSAM / Eigen types include/glim_b200/gtsam_points_compat.hpp touches in
// -DGLIM_B200_WITH_GTSAM mode.  They exist ONLY so that mode can be compile-checked in an environment without GTSAM
// (tests/test_cpp_shim.py); they implement no GTSAM behaviour.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <vector>
namespace gtsam {
using Key = std::uint64_t;
using KeyVector = std::vector<Key>;
template <int R, int C>
struct FixedMatrix {
  double v[R * C] = {};
  double& operator()(int r, int c) { return v[c * R + r]; }
  double operator()(int r, int c) const { return v[c * R + r]; }
  double& operator()(int r) { return v[r]; }
  double operator()(int r) const { return v[r]; }
};
using Matrix4 = FixedMatrix<4, 4>;
using Matrix6 = FixedMatrix<6, 6>;
using Vector6 = FixedMatrix<6, 1>;
class Pose3 {
public:
  Pose3() { for (int i = 0; i < 4; i++) m_(i, i) = 1.0; }
  explicit Pose3(const Matrix4& m) : m_(m) {}
  Matrix4 matrix() const { return m_; }
private:
  Matrix4 m_;
};
class Values {
public:
  template <typename T> const T& at(Key k) const { return poses_.at(k); }
  void insert(Key k, const Pose3& p) { poses_[k] = p; }
private:
  std::map<Key, Pose3> poses_;
};
class GaussianFactor { public: virtual ~GaussianFactor() = default; };
class HessianFactor : public GaussianFactor {
public:
  HessianFactor(Key j, const Matrix6& G, const Vector6& g, double f) : keys{j}, G11(G), g1(g), f(f) {}
  HessianFactor(Key j1, Key j2, const Matrix6& G11, const Matrix6& G12, const Vector6& g1, const Matrix6& G22, const Vector6& g2, double f) : keys{j1, j2}, G11(G11), G12(G12), G22(G22), g1(g1), g2(g2), f(f) {}
  KeyVector keys; Matrix6 G11, G12, G22; Vector6 g1, g2; double f;
};
class NonlinearFactor {
public:
  using shared_ptr = std::shared_ptr<NonlinearFactor>;
  NonlinearFactor() = default;
  explicit NonlinearFactor(const KeyVector& keys) : keys_(keys) {}
  virtual ~NonlinearFactor() = default;
  virtual std::size_t dim() const = 0;
  virtual double error(const Values& c) const = 0;
  virtual std::shared_ptr<GaussianFactor> linearize(const Values& c) const = 0;
  virtual shared_ptr clone() const = 0;
  const KeyVector& keys() const { return keys_; }
protected:
  KeyVector keys_;
};
class NonlinearFactorGraph {
public:
  void add(const NonlinearFactor::shared_ptr& f) { factors_.push_back(f); }
  std::vector<NonlinearFactor::shared_ptr>::const_iterator begin() const { return factors_.begin(); }
  std::vector<NonlinearFactor::shared_ptr>::const_iterator end() const { return factors_.end(); }
private:
  std::vector<NonlinearFactor::shared_ptr> factors_;
};
}  // namespace gtsam
