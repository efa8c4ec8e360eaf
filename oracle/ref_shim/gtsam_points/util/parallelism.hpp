#pragma once
namespace gtsam_points {
inline bool is_omp_default() { return true; }
}  // namespace gtsam_points
