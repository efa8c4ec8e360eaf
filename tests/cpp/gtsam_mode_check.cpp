// Compile-only check of the -DGLIM_B200_WITH_GTSAM build of the shims against signature stubs (tests/cpp/gtsam_stub).
#define GLIM_B200_WITH_GTSAM 1
#include "glim_b200/gtsam_points_compat.hpp"
int gtsam_mode_check(const gtsam_points::GaussianVoxelMap::ConstPtr& vm, const gtsam_points::PointCloud::ConstPtr& pc) {
  gtsam::NonlinearFactorGraph graph;
  auto f = std::make_shared<gtsam_points::IntegratedVGICPFactorGPU>(gtsam::Key(0), gtsam::Key(1), vm, pc);
  f->set_enable_surface_validation(true);
  graph.add(f);
  graph.add(std::make_shared<gtsam_points::IntegratedVGICPFactorGPU>(gtsam::Pose3(), gtsam::Key(1), vm, pc));
  gtsam::Values values;
  values.insert(0, gtsam::Pose3());
  values.insert(1, gtsam::Pose3());
  gtsam_points::NonlinearFactorSetGPU set;
  set.add(graph);
  set.linearize(values);
  std::shared_ptr<gtsam::GaussianFactor> h = f->linearize(values);
  return (int)f->dim() + (int)f->keys().size() + (f->clone() != nullptr) + (h != nullptr) + (f->error(values) >= 0.0);
}
