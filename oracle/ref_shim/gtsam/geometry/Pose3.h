// stand-in for gtsam::Pose3 (GTSAM is not installed here): cloud_deskewing.cpp uses Vector6 and Pose3::Expmap(xi).matrix()
// only.  xi = [omega; v]; Rot3::Expmap by Rodrigues' formula, translation as GTSAM documents it:
//   t = (omega x v - R (omega x v) + omega (omega . v)) / |omega|^2   (|omega|^2 > eps), else t = v.
#pragma once
#include <Eigen/Core>
namespace gtsam {
using Vector6 = Eigen::Mat<6, 1>;
class Pose3 {
public:
  static Pose3 Expmap(const Vector6& xi) {
    const Eigen::Vector3d w(xi[0], xi[1], xi[2]), v(xi[3], xi[4], xi[5]);
    const double theta2 = w.dot(w);
    Pose3 p;
    p.m_ = Eigen::Matrix4d::Identity();
    Eigen::Matrix3d R = Eigen::Matrix3d::Identity();
    Eigen::Vector3d t = v;
    if (theta2 > std::numeric_limits<double>::epsilon()) {  // below it GTSAM uses the first-order rotation I + hat(omega) and t = v
      const double theta = std::sqrt(theta2), s = std::sin(theta), c = std::cos(theta);
      Eigen::Matrix3d K = Eigen::Matrix3d::Zero();
      K(0, 1) = -w[2]; K(0, 2) = w[1]; K(1, 0) = w[2]; K(1, 2) = -w[0]; K(2, 0) = -w[1]; K(2, 1) = w[0];
      R = Eigen::Matrix3d::Identity() + K * (s / theta) + (K * K) * ((1.0 - c) / theta2);
      const Eigen::Vector3d wxv = w.cross(v);
      t = (wxv - R * wxv + w * w.dot(v)) / theta2;
    } else {
      R(0, 1) = -w[2]; R(0, 2) = w[1]; R(1, 0) = w[2]; R(1, 2) = -w[0]; R(2, 0) = -w[1]; R(2, 1) = w[0];
    }
    p.m_.block<3, 3>(0, 0) = R;
    p.m_.block<3, 1>(0, 3) = t;
    return p;
  }
  const Eigen::Matrix4d& matrix() const { return m_; }

private:
  Eigen::Matrix4d m_;
};
}  // namespace gtsam
