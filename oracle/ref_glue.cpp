// ref_glue.cpp -- C entry points around the reference's OWN translation units
//   /root/reference/src/glim/common/cloud_covariance_estimation.cpp   (glim::CloudCovarianceEstimation)
//   /root/reference/src/glim/common/cloud_deskewing.cpp               (glim::CloudDeskewing)
//   /root/reference/src/glim/preprocess/cloud_preprocessor.cpp        (glim::CloudPreprocessor; + callbacks.cpp)
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

// ---- glim::CloudPreprocessor::preprocess (src/glim/preprocess/cloud_preprocessor.cpp, compiled unmodified; its gtsam_points
// leaf calls resolve to the stand-ins of ref_shim/gtsam_points/, which forward to the oracle's restatements) ----
#include <glim/preprocess/cloud_preprocessor.hpp>
#include <gtsam_points/types/point_cloud_cpu.hpp>

namespace gtsam_points {
uint64_t g_ref_shim_seed = 0;
}

struct ref_preprocess_params {  // mirrors gb_preprocess_params where the reference has the field
  double distance_near_thresh, distance_far_thresh;
  int use_random_grid_downsampling;
  double downsample_resolution;
  int downsample_target;
  double downsample_rate;
  unsigned long long seed;
  int global_shutter;
  int crop_bbox_frame;  // 0 = off, 1 = "lidar", 2 = "imu"
  double crop_bbox_min[3], crop_bbox_max[3];
  double T_imu_lidar[16];
  int enable_outlier_removal, outlier_removal_k;
  double outlier_std_mul_factor;
  int k_correspondences, num_threads;
};

// returns the number of frame points; outputs sized for n raw points; neighbors row-major [i * k + j]
REF_API int ref_preprocess(const ref_preprocess_params* P, double stamp, int n, const double* pts4, const double* times, const double* intensities,
                           double* out_pts4, double* out_times, double* out_intensities, int* out_neighbors, double* scan_end_time, double* defaults_seen /* 6 */) {
  glim::CloudPreprocessorParams params;  // code defaults through the stand-in Config (an empty config file)
  if (defaults_seen) {
    defaults_seen[0] = params.distance_near_thresh; defaults_seen[1] = params.distance_far_thresh; defaults_seen[2] = params.downsample_resolution;
    defaults_seen[3] = params.downsample_rate; defaults_seen[4] = params.outlier_std_mul_factor; defaults_seen[5] = params.k_correspondences;
  }
  params.distance_near_thresh = P->distance_near_thresh; params.distance_far_thresh = P->distance_far_thresh;
  params.use_random_grid_downsampling = P->use_random_grid_downsampling != 0;
  params.downsample_resolution = P->downsample_resolution; params.downsample_target = P->downsample_target; params.downsample_rate = P->downsample_rate;
  params.global_shutter = P->global_shutter != 0;
  params.enable_cropbox_filter = P->crop_bbox_frame != 0;
  params.crop_bbox_frame = P->crop_bbox_frame == 2 ? "imu" : "lidar";
  params.crop_bbox_min = Eigen::Vector3d(P->crop_bbox_min[0], P->crop_bbox_min[1], P->crop_bbox_min[2]);
  params.crop_bbox_max = Eigen::Vector3d(P->crop_bbox_max[0], P->crop_bbox_max[1], P->crop_bbox_max[2]);
  params.T_imu_lidar = to_iso(P->T_imu_lidar);
  params.enable_outlier_removal = P->enable_outlier_removal != 0; params.outlier_removal_k = P->outlier_removal_k; params.outlier_std_mul_factor = P->outlier_std_mul_factor;
  params.k_correspondences = P->k_correspondences; params.num_threads = P->num_threads;
  gtsam_points::g_ref_shim_seed = P->seed;

  auto raw = std::make_shared<glim::RawPoints>();
  raw->stamp = stamp;
  raw->points = to_points(n, pts4);
  raw->times.assign(times, times + n);
  if (intensities) raw->intensities.assign(intensities, intensities + n);

  glim::CloudPreprocessor pre(params);
  const glim::PreprocessedFrame::Ptr fr = pre.preprocess(raw);
  const int m = fr->size();
  for (int i = 0; i < m; i++) std::memcpy(out_pts4 + 4 * (size_t)i, fr->points[i].a, 4 * sizeof(double));
  std::memcpy(out_times, fr->times.data(), sizeof(double) * m);
  if (out_intensities && fr->intensities.size()) std::memcpy(out_intensities, fr->intensities.data(), sizeof(double) * m);
  std::memcpy(out_neighbors, fr->neighbors.data(), sizeof(int) * fr->neighbors.size());
  *scan_end_time = fr->scan_end_time;
  return m;
}

// ---- the stand-ins themselves, exposed so that tests can check them against numpy / scipy (they are OUR restatements of Eigen's and
// GTSAM's published algorithms: tests/test_oracle_vs_reference_tu.py::test_stand_in_primitives_*) ----
REF_API void ref_shim_eigen_sym3(const double* A9_colmajor, double* vals3, double* vecs9_colmajor) {
  Eigen::Matrix3d m;
  std::memcpy(m.a, A9_colmajor, sizeof(m.a));
  Eigen::SelfAdjointEigenSolver<Eigen::Matrix3d> eig;
  eig.computeDirect(m);
  std::memcpy(vals3, eig.eigenvalues().a, 3 * sizeof(double));
  std::memcpy(vecs9_colmajor, eig.eigenvectors().a, 9 * sizeof(double));
}
REF_API void ref_shim_slerp(const double* R0_colmajor, const double* R1_colmajor, double t, double* R_colmajor) {
  Eigen::Matrix3d a, b;
  std::memcpy(a.a, R0_colmajor, sizeof(a.a));
  std::memcpy(b.a, R1_colmajor, sizeof(b.a));
  const Eigen::Matrix3d r = Eigen::Quaterniond(a).slerp(t, Eigen::Quaterniond(b)).toRotationMatrix();
  std::memcpy(R_colmajor, r.a, sizeof(r.a));
}
#include <gtsam/geometry/Pose3.h>
REF_API void ref_shim_pose3_expmap(const double* xi6 /* omega, v */, double* T16_colmajor) {
  gtsam::Vector6 xi;
  std::memcpy(xi.a, xi6, sizeof(xi.a));
  std::memcpy(T16_colmajor, gtsam::Pose3::Expmap(xi).matrix().a, 16 * sizeof(double));
}
