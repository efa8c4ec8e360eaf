// gtsam_eigen_mode_callsites.cpp -- COMPILE-ONLY check of the -DGLIM_B200_WITH_GTSAM build of the shims with Eigen-TYPED frames:
// the statements of GLIM's GPU call sites (cited per block) against include/glim_b200/gtsam_points_compat.hpp, with
//   * Eigen::Vector4d / Matrix4d / Isometry3d from the stand-in of oracle/ref_shim/ (in this mode the shim's points / covs ARE
//     `Eigen::` types, so any place where the shim would rely on its own POD members fails to compile here), and
//   * gtsam::Key / Pose3 / Values / NonlinearFactorGraph / HessianFactor from the signature stubs of tests/cpp/gtsam_stub/.
// Neither Eigen nor GTSAM is installed in this environment; with them, these are the lines GLIM itself compiles.
#include <Eigen/Core>
#include <Eigen/Geometry>
#define GLIM_B200_WITH_GTSAM 1
#include "glim_b200/gtsam_points_compat.hpp"
#include "glim_b200/glim_preprocess_compat.hpp"

#include <any>
#include <cmath>

struct EstimationFrame {  // the members of glim::EstimationFrame the call sites touch (include/glim/odometry/estimation_frame.hpp)
  using ConstPtr = std::shared_ptr<const EstimationFrame>;
  gtsam_points::PointCloud::ConstPtr frame;
  std::vector<gtsam_points::GaussianVoxelMap::Ptr> voxelmaps;
  Eigen::Isometry3d T_world_imu;
};

double callsites(EstimationFrame& new_frame, const EstimationFrame::ConstPtr& target, const EstimationFrame::ConstPtr& source, const gtsam::Values& values) {
  // ---- OdometryEstimationGPU ctor (odometry_estimation_gpu.cpp:76-77)
  std::unique_ptr<gtsam_points::CUDAStream> stream(new gtsam_points::CUDAStream());
  std::unique_ptr<gtsam_points::StreamTempBufferRoundRobin> stream_buffer_roundrobin(new gtsam_points::StreamTempBufferRoundRobin());

  // ---- create_frame (odometry_estimation_gpu.cpp:89-106)
  const int max_scan_count = 256;
  const double dist_median = gtsam_points::median_distance(new_frame.frame, max_scan_count);
  const double p = std::max(0.0, std::min(1.0, (dist_median - 5.0) / (20.0 - 5.0)));
  const double base_resolution = 0.25 + p * (0.5 - 0.25);
  new_frame.frame = gtsam_points::PointCloudGPU::clone(*new_frame.frame);
  for (int i = 0; i < 2; i++) {
    if (!new_frame.frame->size()) break;
    const double resolution = base_resolution * std::pow(2.0, i);
    auto voxelmap = std::make_shared<gtsam_points::GaussianVoxelMapGPU>(resolution, 8192 * 2, 10, 1e-3, *stream);
    voxelmap->insert(*new_frame.frame);
    new_frame.voxelmaps.push_back(voxelmap);
  }
  // host reads of the Eigen-typed frame AFTER the clone replaced it (median_distance, deskewing, viewer)
  const Eigen::Vector4d first = new_frame.frame->points[0];
  const double range = first.head<3>().norm() + (new_frame.frame->covs ? new_frame.frame->covs[0](0, 0) : 0.0);

  // ---- create_factors (odometry_estimation_gpu.cpp:128-165): binary and fixed-target-pose forms, stream + buffer from the round robin
  gtsam::NonlinearFactorGraph factors;
  {
    auto stream_buffer = stream_buffer_roundrobin->get_stream_buffer();
    const auto& fstream = stream_buffer.first;
    const auto& buffer = stream_buffer.second;
    for (const auto& voxelmap : target->voxelmaps) {
      auto factor = std::make_shared<gtsam_points::IntegratedVGICPFactorGPU>(gtsam::Key(0), gtsam::Key(1), voxelmap, source->frame, fstream, buffer);
      factor->set_enable_surface_validation(true);
      factors.add(factor);
      const gtsam::Pose3 fixed_target_pose;
      auto unary = std::make_shared<gtsam_points::IntegratedVGICPFactorGPU>(fixed_target_pose, gtsam::Key(1), voxelmap, source->frame, fstream, buffer);
      unary->set_enable_surface_validation(true);
      factors.add(unary);
    }
  }

  // ---- keyframe bookkeeping (odometry_estimation_gpu.cpp:224-231, :247-248): Eigen::Isometry3d deltas
  std::vector<gtsam_points::GaussianVoxelMap::ConstPtr> keyframes_(1);
  std::vector<Eigen::Isometry3d> delta_from_keyframes(1);
  keyframes_[0] = target->voxelmaps.back();
  delta_from_keyframes[0] = target->T_world_imu.inverse() * source->T_world_imu;
  double overlap = gtsam_points::overlap_gpu(keyframes_, source->frame, delta_from_keyframes, *stream);
  const Eigen::Isometry3d delta = target->T_world_imu.inverse() * source->T_world_imu;
  overlap += gtsam_points::overlap_gpu(target->voxelmaps.back(), source->frame, delta, *stream);

  // ---- batch linearization (odometry_estimation_gpu.cpp:383-386)
  gtsam_points::NonlinearFactorSetGPU factor_set;
  factor_set.add(factors);
  factor_set.linearize(values);
  double err = 0.0;
  for (const auto& f : factors) {
    std::shared_ptr<gtsam::GaussianFactor> linearized = f->linearize(values);
    err += f->error(values) + (linearized ? 1.0 : 0.0) + (double)f->keys().size() + (f->clone() ? 1.0 : 0.0);
  }

  // ---- sub_mapping.cpp:165-169 / global_mapping.cpp:252-266: upload only if the frame is not on the device yet
  gtsam_points::PointCloud::ConstPtr frame = source->frame;
  if (!frame->points_gpu) frame = gtsam_points::PointCloudGPU::clone(*frame, *stream);
  auto submap_voxelmap = std::make_shared<gtsam_points::GaussianVoxelMapGPU>(0.5);
  submap_voxelmap->insert(*frame);
  overlap += gtsam_points::overlap_auto(submap_voxelmap, frame, delta);

  // ---- the preprocess shims with Eigen-typed containers (odometry_estimation_imu.cpp:313-328)
  glim_b200::glim::CloudDeskewing deskewing;
  glim_b200::glim::CloudCovarianceEstimation covariance_estimation(2);
  const std::vector<double> times(2, 0.0), pred_imu_times(2, 0.0);
  const std::vector<Eigen::Isometry3d> pred_imu_poses(2, delta);
  const std::vector<Eigen::Vector4d> raw_points(2, first);
  const std::vector<Eigen::Vector4d> deskewed = deskewing.deskew(delta, pred_imu_times, pred_imu_poses, 10.0, times, raw_points);
  const std::vector<Eigen::Vector4d> by_velocity = deskewing.deskew(delta, Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), times, raw_points);
  std::vector<Eigen::Vector4d> normals;
  std::vector<Eigen::Matrix4d> covs;
  covariance_estimation.estimate(deskewed, std::vector<int>(4, 0), normals, covs);
  return range + overlap + err + (double)by_velocity.size() + (double)covs.size();
}
