// glim_preprocess_compat.hpp -- header-only C++17 shims with the interface of GLIM's own per-frame preprocess classes, forwarding
// to the device-resident pipeline of libglim_b200.so (gb_preprocess / gb_covariances, include/glim_b200.h):
//
//   glim::CloudPreprocessorParams / glim::CloudPreprocessor::preprocess(raw_points) -> PreprocessedFrame
//        include/glim/preprocess/cloud_preprocessor.hpp:14-75, src/glim/preprocess/cloud_preprocessor.cpp:77-221
//   glim::PreprocessedFrame / glim::RawPoints        include/glim/preprocess/preprocessed_frame.hpp:14-38
//   glim::CloudCovarianceEstimation::estimate         include/glim/common/cloud_covariance_estimation.hpp:17-60,
//        src/glim/common/cloud_covariance_estimation.cpp:24-122  (called at src/glim/odometry/odometry_estimation_imu.cpp:322-328)
//   glim::CloudDeskewing::deskew (both overloads)     include/glim/common/cloud_deskewing.hpp:11-50, src/glim/common/cloud_deskewing.cpp:11-133
//        (called at src/glim/odometry/odometry_estimation_imu.cpp:313, src/glim/mapping/sub_mapping.cpp:365)
//
// Same member names and call signatures; Eigen::Vector4d / Matrix4d are the layout-compatible PODs of
// gtsam_points_compat.hpp (real Eigen types with -DGLIM_B200_WITH_GTSAM).  Namespace `glim_b200::glim` so that the header can
// sit next to GLIM's own in one translation unit; `using namespace glim_b200;` (or a namespace alias) makes the names line up.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "gtsam_points_compat.hpp"

namespace glim_b200 {
namespace glim {

using gtsam_points::Matrix4d;
using gtsam_points::Vector4d;

/// include/glim/preprocess/preprocessed_frame.hpp (RawPoints: stamp, times, intensities, points)
struct RawPoints {
  using Ptr = std::shared_ptr<RawPoints>;
  using ConstPtr = std::shared_ptr<const RawPoints>;
  int size() const { return static_cast<int>(points.size()); }
  double stamp = 0.0;
  std::vector<double> times;
  std::vector<double> intensities;
  std::vector<Vector4d> points;
};

struct PreprocessedFrame {
  using Ptr = std::shared_ptr<PreprocessedFrame>;
  using ConstPtr = std::shared_ptr<const PreprocessedFrame>;
  int size() const { return static_cast<int>(points.size()); }
  double stamp = 0.0;
  double scan_end_time = 0.0;
  std::vector<double> times;
  std::vector<double> intensities;
  std::vector<Vector4d> points;
  int k_neighbors = 0;
  std::vector<int> neighbors;
  RawPoints::ConstPtr raw_points;
};

/// glim::CloudPreprocessorParams with the CODE defaults of cloud_preprocessor.cpp:28-61 (GLIM reads them from its config)
struct CloudPreprocessorParams {
  double distance_near_thresh = 1.0;
  double distance_far_thresh = 100.0;
  bool global_shutter = false;
  bool use_random_grid_downsampling = false;
  double downsample_resolution = 0.15;
  int downsample_target = 0;
  double downsample_rate = 0.3;
  bool enable_outlier_removal = false;
  int outlier_removal_k = 10;
  double outlier_std_mul_factor = 2.0;
  bool enable_cropbox_filter = false;
  std::string crop_bbox_frame = "lidar";
  double crop_bbox_min[3] = {0.0, 0.0, 0.0};
  double crop_bbox_max[3] = {0.0, 0.0, 0.0};
  Pose T_imu_lidar;
  int k_correspondences = 8;
  int num_threads = 2;  // unused: the pipeline runs on the GPU
};

class CloudPreprocessor {
public:
  explicit CloudPreprocessor(const CloudPreprocessorParams& params = CloudPreprocessorParams(), CUstream_st* stream = nullptr, std::uint64_t seed = 0) : params(params), stream_(stream), seed_(seed) {
    if (params.enable_cropbox_filter && params.crop_bbox_frame != "lidar" && params.crop_bbox_frame != "imu") throw std::runtime_error("Unsupported crop bbox frame: " + params.crop_bbox_frame);  // cloud_preprocessor.cpp:51
  }
  virtual ~CloudPreprocessor() = default;

  /// cloud_preprocessor.cpp:77-188 in one device-resident call
  virtual PreprocessedFrame::Ptr preprocess(const RawPoints::ConstPtr& raw_points) {
    gb_preprocess_params cp = c_params();
    cp.estimate_covariances = 0;
    const std::size_t n = raw_points->points.size();
    auto out = std::make_shared<PreprocessedFrame>();
    out->stamp = raw_points->stamp;
    out->raw_points = raw_points;
    out->k_neighbors = params.k_correspondences;
    out->times.resize(n);
    out->points.resize(n);
    if (!raw_points->intensities.empty()) out->intensities.resize(n);
    out->neighbors.resize(n * static_cast<std::size_t>(params.k_correspondences));
    gb_preprocessed res{};
    res.times = out->times.data();
    res.xyzw = reinterpret_cast<double*>(out->points.data());
    res.intensities = out->intensities.empty() ? nullptr : out->intensities.data();
    res.neighbors = out->neighbors.data();
    check(gb_preprocess(Context::of_stream(stream_), n, reinterpret_cast<const double*>(raw_points->points.data()), raw_points->times.empty() ? nullptr : raw_points->times.data(),
                        raw_points->intensities.empty() ? nullptr : raw_points->intensities.data(), &cp, &res),
          "gb_preprocess");
    const std::size_t m = res.num_points;
    out->times.resize(m);
    out->points.resize(m);
    if (!out->intensities.empty()) out->intensities.resize(m);
    out->neighbors.resize(m * static_cast<std::size_t>(params.k_correspondences));
    out->scan_end_time = raw_points->stamp + res.last_time;  // cloud_preprocessor.cpp:174
    return out;
  }

  /// preprocess + CloudCovarianceEstimation::estimate + PointCloudGPU::clone in one call: the frame as GLIM's odometry wants it
  /// (odometry_estimation_imu.cpp:322-328 + odometry_estimation_gpu.cpp:96), covariances / normals on the host AND on the device
  gtsam_points::PointCloudGPU::Ptr preprocess_to_gpu_frame(const RawPoints::ConstPtr& raw_points, PreprocessedFrame::Ptr* preprocessed = nullptr);

  gb_preprocess_params c_params() const {
    gb_preprocess_params cp;
    gb_preprocess_default_params(&cp);
    cp.distance_near_thresh = params.distance_near_thresh;
    cp.distance_far_thresh = params.distance_far_thresh;
    cp.use_random_grid_downsampling = params.use_random_grid_downsampling;
    cp.downsample_resolution = params.downsample_resolution;
    cp.downsample_target = params.downsample_target;
    cp.downsample_rate = params.downsample_rate;
    cp.seed = seed_;
    cp.global_shutter = params.global_shutter;
    cp.crop_bbox_frame = params.enable_cropbox_filter ? (params.crop_bbox_frame == "imu" ? 2 : 1) : 0;
    for (int a = 0; a < 3; a++) { cp.crop_bbox_min[a] = params.crop_bbox_min[a]; cp.crop_bbox_max[a] = params.crop_bbox_max[a]; }
    for (int e = 0; e < 16; e++) cp.T_imu_lidar[e] = params.T_imu_lidar.m[static_cast<std::size_t>(e)];
    cp.enable_outlier_removal = params.enable_outlier_removal;
    cp.outlier_removal_k = params.outlier_removal_k;
    cp.outlier_std_mul_factor = params.outlier_std_mul_factor;
    cp.k_correspondences = params.k_correspondences;
    cp.estimate_covariances = 1;
    return cp;
  }

private:
  CloudPreprocessorParams params;
  CUstream_st* stream_;
  std::uint64_t seed_;
};

enum class RegularizationMethod { NONE, PLANE, NORMALIZED_MIN_EIG, FROBENIUS };

/// glim::CloudCovarianceEstimation (PLANE regularization, the method GLIM hard-wires at cloud_covariance_estimation.cpp:20)
class CloudCovarianceEstimation {
public:
  explicit CloudCovarianceEstimation(const int /*num_threads*/ = 1, CUstream_st* stream = nullptr) : stream_(stream) {}

  void estimate(const std::vector<Vector4d>& points, const std::vector<int>& neighbors, std::vector<Vector4d>& normals, std::vector<Matrix4d>& covs) const {
    if (points.empty()) { normals.clear(); covs.clear(); return; }  // cloud_covariance_estimation.cpp:30-32
    const int k = static_cast<int>(neighbors.size() / points.size());
    if (static_cast<std::size_t>(k) * points.size() != neighbors.size()) throw std::runtime_error("k * points.size() != neighbors.size()");  // :34-38 (spdlog::critical + abort there)
    estimate(points, neighbors, k, normals, covs);
  }
  void estimate(const std::vector<Vector4d>& points, const std::vector<int>& neighbors, const int k_neighbors, std::vector<Vector4d>& normals, std::vector<Matrix4d>& covs) const {
    normals.resize(points.size());
    covs.resize(points.size());
    if (points.empty()) return;
    const int kc = static_cast<int>(neighbors.size() / points.size());
    check(gb_covariances(Context::of_stream(stream_), points.size(), reinterpret_cast<const double*>(points.data()), neighbors.data(), kc, k_neighbors, reinterpret_cast<double*>(normals.data()), reinterpret_cast<double*>(covs.data())),
          "gb_covariances");
  }
  std::vector<Matrix4d> estimate(const std::vector<Vector4d>& points, const std::vector<int>& neighbors, const int k_neighbors) const {
    std::vector<Vector4d> normals;
    std::vector<Matrix4d> covs;
    estimate(points, neighbors, k_neighbors, normals, covs);
    return covs;
  }
  std::vector<Matrix4d> estimate(const std::vector<Vector4d>& points, const std::vector<int>& neighbors) const {
    std::vector<Vector4d> normals;
    std::vector<Matrix4d> covs;
    estimate(points, neighbors, normals, covs);
    return covs;
  }

private:
  CUstream_st* stream_;
};

/// Layout-compatible stand-in for Eigen::Vector3d (24 B); the real type with -DGLIM_B200_WITH_GTSAM
#if defined(GLIM_B200_WITH_GTSAM) && defined(EIGEN_WORLD_VERSION)
using Vector3d = Eigen::Vector3d;
#else
struct Vector3d {
  double v[3] = {0.0, 0.0, 0.0};
  Vector3d() = default;
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  static Vector3d Zero() { return Vector3d(); }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  const double* data() const { return v; }
};
#endif
static_assert(sizeof(Pose) == 16 * sizeof(double), "std::vector<Pose> is a dense array of column-major 4x4 matrices");

/// glim::CloudDeskewing: same two overloads, same argument order.  The per-point transform runs on the device (gb_deskew); the
/// time table and the pose per 0.1 ms slot are built on the host, as the reference does (cloud_deskewing.cpp:22-43, :75-121).
class CloudDeskewing {
public:
  explicit CloudDeskewing(CUstream_st* stream = nullptr) : stream_(stream) {}
  ~CloudDeskewing() = default;

  /// constant linear / angular velocity (cloud_deskewing.cpp:11-55)
  std::vector<Vector4d> deskew(const Pose& T_imu_lidar, const Vector3d& linear_vel, const Vector3d& angular_vel, const std::vector<double>& times, const std::vector<Vector4d>& points) {
    return run(T_imu_lidar, linear_vel.data(), angular_vel.data(), nullptr, nullptr, 0.0, times, points, nullptr);
  }
  /// predicted IMU poses (cloud_deskewing.cpp:57-133); no poses -> the zero-velocity model (:69-71)
  std::vector<Vector4d> deskew(const Pose& T_imu_lidar, const std::vector<double>& imu_times, const std::vector<Pose>& imu_poses, const double stamp, const std::vector<double>& times,
                               const std::vector<Vector4d>& points) {
    return run(T_imu_lidar, nullptr, nullptr, &imu_times, &imu_poses, stamp, times, points, nullptr);
  }
  /// the same with the container GLIM passes: std::vector<Eigen::Isometry3d> pred_imu_poses (odometry_estimation_imu.cpp:313)
  template <class Iso, class Alloc, class = std::enable_if_t<!std::is_same<Iso, Pose>::value && std::is_constructible<Pose, const Iso&>::value>>
  std::vector<Vector4d> deskew(const Pose& T_imu_lidar, const std::vector<double>& imu_times, const std::vector<Iso, Alloc>& imu_poses, const double stamp, const std::vector<double>& times,
                               const std::vector<Vector4d>& points) {
    return deskew(T_imu_lidar, imu_times, std::vector<Pose>(imu_poses.begin(), imu_poses.end()), stamp, times, points);
  }
  /// extension: the `pt = T_imu_lidar * pt` loop that follows the call (odometry_estimation_imu.cpp:314-316) fused into the kernel
  std::vector<Vector4d> deskew_and_transform(const Pose& T_imu_lidar, const std::vector<double>& imu_times, const std::vector<Pose>& imu_poses, const double stamp, const std::vector<double>& times,
                                             const std::vector<Vector4d>& points, const Pose& T_post) {
    return run(T_imu_lidar, nullptr, nullptr, &imu_times, &imu_poses, stamp, times, points, &T_post);
  }

  /// The host half on its own (no device needed): slot index per point and T_lidar0_lidar1 per slot -- what deskew() hands to the
  /// kernel.  velocities / imu arguments as in the two overloads (pass nullptr for the half that is not used).
  static void pose_table(const Pose& T_imu_lidar, const Vector3d* linear_vel, const Vector3d* angular_vel, const std::vector<double>* imu_times, const std::vector<Pose>* imu_poses, const double stamp,
                         const std::vector<double>& times, std::vector<int>& time_indices, std::vector<Pose>& T_lidar0_lidar1) {
    const std::size_t n = times.size();
    time_indices.assign(n, 0);
    T_lidar0_lidar1.assign(n, Pose());
    std::size_t slots = 0;
    const std::size_t n_imu = (imu_times && imu_poses) ? imu_poses->size() : 0;
    if (imu_times && imu_poses && imu_times->size() != imu_poses->size()) throw std::runtime_error("imu_times.size() != imu_poses.size()");
    check(gb_deskew_pose_table(T_imu_lidar.data(), linear_vel ? linear_vel->data() : nullptr, angular_vel ? angular_vel->data() : nullptr, n_imu, n_imu ? imu_times->data() : nullptr,
                               n_imu ? (*imu_poses)[0].data() : nullptr, stamp, n, times.data(), time_indices.data(), n ? const_cast<double*>(T_lidar0_lidar1[0].data()) : nullptr, &slots),
          "gb_deskew_pose_table");
    T_lidar0_lidar1.resize(slots);
  }

private:
  std::vector<Vector4d> run(const Pose& T_imu_lidar, const double* linear_vel, const double* angular_vel, const std::vector<double>* imu_times, const std::vector<Pose>* imu_poses, const double stamp,
                            const std::vector<double>& times, const std::vector<Vector4d>& points, const Pose* T_post) {
    if (times.empty()) return std::vector<Vector4d>();  // cloud_deskewing.cpp:17-19, :65-67
    if (times.size() != points.size()) throw std::runtime_error("times.size() != points.size()");
    const std::size_t n_imu = (imu_times && imu_poses) ? imu_poses->size() : 0;
    if (imu_times && imu_poses && imu_times->size() != imu_poses->size()) throw std::runtime_error("imu_times.size() != imu_poses.size()");
    std::vector<Vector4d> out(points.size());
    check(gb_deskew(Context::of_stream(stream_), T_imu_lidar.data(), linear_vel, angular_vel, n_imu, n_imu ? imu_times->data() : nullptr, n_imu ? (*imu_poses)[0].data() : nullptr, stamp, points.size(),
                    times.data(), reinterpret_cast<const double*>(points.data()), T_post ? T_post->data() : nullptr, reinterpret_cast<double*>(out.data())),
          "gb_deskew");
    return out;
  }
  CUstream_st* stream_;
};

inline gtsam_points::PointCloudGPU::Ptr CloudPreprocessor::preprocess_to_gpu_frame(const RawPoints::ConstPtr& raw_points, PreprocessedFrame::Ptr* preprocessed) {
  gb_preprocess_params cp = c_params();
  const std::size_t n = raw_points->points.size();
  auto fr = std::make_shared<PreprocessedFrame>();
  fr->stamp = raw_points->stamp;
  fr->raw_points = raw_points;
  fr->k_neighbors = params.k_correspondences;
  fr->times.resize(n);
  fr->points.resize(n);
  if (!raw_points->intensities.empty()) fr->intensities.resize(n);
  fr->neighbors.resize(n * static_cast<std::size_t>(params.k_correspondences));
  std::vector<Vector4d> normals(n);
  std::vector<Matrix4d> covs(n);
  gb_preprocessed res{};
  res.times = fr->times.data();
  res.xyzw = reinterpret_cast<double*>(fr->points.data());
  res.intensities = fr->intensities.empty() ? nullptr : fr->intensities.data();
  res.neighbors = fr->neighbors.data();
  res.normals4 = reinterpret_cast<double*>(normals.data());
  res.cov4x4 = reinterpret_cast<double*>(covs.data());
  check(gb_preprocess(Context::of_stream(stream_), n, reinterpret_cast<const double*>(raw_points->points.data()), raw_points->times.empty() ? nullptr : raw_points->times.data(),
                      raw_points->intensities.empty() ? nullptr : raw_points->intensities.data(), &cp, &res),
        "gb_preprocess");
  const std::size_t m = res.num_points;
  fr->times.resize(m);
  fr->points.resize(m);
  if (!fr->intensities.empty()) fr->intensities.resize(m);
  fr->neighbors.resize(m * static_cast<std::size_t>(params.k_correspondences));
  fr->scan_end_time = raw_points->stamp + res.last_time;
  if (preprocessed) *preprocessed = fr;
  // the host side of the frame (what PointCloudGPU::clone would have copied) + the device cloud gb_preprocess already built
  return gtsam_points::PointCloudGPU::adopt(reinterpret_cast<const double*>(fr->points.data()), reinterpret_cast<const double*>(covs.data()), reinterpret_cast<const double*>(normals.data()), fr->times.data(),
                                            fr->intensities.empty() ? nullptr : fr->intensities.data(), m, res.cloud);
}

}  // namespace glim
}  // namespace glim_b200
