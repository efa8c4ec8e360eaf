#pragma once
namespace gtsam_points {
inline bool is_omp_default() { return true; }
inline bool is_tbb_default() { return false; }
}  // namespace gtsam_points
