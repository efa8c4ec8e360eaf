// ref_glue.cpp -- C entry points around the reference's OWN translation units
//   /root/reference/src/glim/common/cloud_covariance_estimation.cpp   (glim::CloudCovarianceEstimation)
//   /root/reference/src/glim/common/cloud_deskewing.cpp               (glim::CloudDeskewing)
// which oracle/Makefile (target _ref) compiles UNMODIFIED, from where they lie, against the stand-in headers of
// oracle/ref_shim/ (Eigen, GTSAM, spdlog and gtsam_points are not installed here; the stand-ins are written in this repository).
// TEST INFRASTRUCTURE ONLY: tests/test_oracle_vs_reference_tu.py uses the resulting oracle/_ref/libglim_ref.so to check that
// oracle/glim_oracle.c restates those files faithfully.  Same argument conventions as go_covariance_estimate / go_deskew_*.
#include <glim/common/cloud_covariance_estimation.hpp>
#include <glim/common/cloud_deskewing.hpp>

#include <cstring>
#include <vector>

namespace {
std::vector<Eigen::Vector4d> to_points(int n, const double* p4) {
  std::vector<Eigen::Vector4d> v(n);
  for (int i = 0; i < n; i++) std::memcpy(v[i].a, p4 + 4 * (size_t)i, 4 * sizeof(double));
  return v;
}
Eigen::Isometry3d to_iso(const double* T16) {  // 16 doubles, column-major
  Eigen::Matrix4d m;
  std::memcpy(m.a, T16, 16 * sizeof(double));
  return Eigen::Isometry3d(m);
}
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

// CloudCovarianceEstimation(num_threads).estimate(points, neighbors, k_neighbors, normals, covs); covs column-major 4x4
REF_API void ref_covariance_estimate(int n, const double* pts4, const int* neighbors, int k_correspondences, int k_neighbors, int num_threads, double* normals4, double* cov16) {
  const std::vector<Eigen::Vector4d> points = to_points(n, pts4);
  const std::vector<int> nb(neighbors, neighbors + (size_t)n * k_correspondences);
  std::vector<Eigen::Vector4d> normals;
  std::vector<Eigen::Matrix4d> covs;
  glim::CloudCovarianceEstimation est(num_threads);
  est.estimate(points, nb, k_neighbors, normals, covs);
  for (int i = 0; i < n; i++) {
    std::memcpy(normals4 + 4 * (size_t)i, normals[i].a, 4 * sizeof(double));
    std::memcpy(cov16 + 16 * (size_t)i, covs[i].a, 16 * sizeof(double));
  }
}

// the (points, neighbors, k) overload: sample covariance (/(k - 1)), no normals
REF_API void ref_covariance_estimate_sample(int n, const double* pts4, const int* neighbors, int k_correspondences, int k_neighbors, double* cov16) {
  const std::vector<Eigen::Vector4d> points = to_points(n, pts4);
  const std::vector<int> nb(neighbors, neighbors + (size_t)n * k_correspondences);
  glim::CloudCovarianceEstimation est(1);
  const std::vector<Eigen::Matrix4d> covs = est.estimate(points, nb, k_neighbors);
  for (int i = 0; i < n; i++) std::memcpy(cov16 + 16 * (size_t)i, covs[i].a, 16 * sizeof(double));
}

// CloudDeskewing::deskew(T_imu_lidar, linear_vel, angular_vel, times, points)
REF_API void ref_deskew_const_vel(const double* T_imu_lidar, const double* linear_vel, const double* angular_vel, int n, const double* times, const double* pts4, double* out4) {
  glim::CloudDeskewing d;
  const std::vector<double> t(times, times + n);
  const std::vector<Eigen::Vector4d> out =
    d.deskew(to_iso(T_imu_lidar), Eigen::Vector3d(linear_vel[0], linear_vel[1], linear_vel[2]), Eigen::Vector3d(angular_vel[0], angular_vel[1], angular_vel[2]), t, to_points(n, pts4));
  for (size_t i = 0; i < out.size(); i++) std::memcpy(out4 + 4 * i, out[i].a, 4 * sizeof(double));
}

// CloudDeskewing::deskew(T_imu_lidar, imu_times, imu_poses, stamp, times, points); imu_poses n_imu x 16 column-major
REF_API void ref_deskew_imu(const double* T_imu_lidar, int n_imu, const double* imu_times, const double* imu_poses, double stamp, int n, const double* times, const double* pts4, double* out4) {
  glim::CloudDeskewing d;
  const std::vector<double> it(imu_times, imu_times + n_imu), t(times, times + n);
  std::vector<Eigen::Isometry3d> poses;
  for (int i = 0; i < n_imu; i++) poses.push_back(to_iso(imu_poses + 16 * (size_t)i));
  const std::vector<Eigen::Vector4d> out = d.deskew(to_iso(T_imu_lidar), it, poses, stamp, t, to_points(n, pts4));
  for (size_t i = 0; i < out.size(); i++) std::memcpy(out4 + 4 * i, out[i].a, 4 * sizeof(double));
}
