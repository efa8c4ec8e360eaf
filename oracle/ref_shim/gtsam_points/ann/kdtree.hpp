// stand-in for gtsam_points::KdTree: exact k-NN by brute force with the oracle's un-contracted distance and (distance, index) tie rule
#pragma once
#include <Eigen/Core>
#include <algorithm>
#include <utility>
#include <vector>
namespace gtsam_points {
struct KdTree {
  const Eigen::Vector4d* points;
  int n;
  KdTree(const Eigen::Vector4d* p, int num_points) : points(p), n(num_points) {}
  size_t knn_search(const double* pt, size_t k, size_t* k_indices, double* k_sq_dists) const {
    std::vector<std::pair<double, size_t>> best;  // ascending (distance, index), at most k entries
    best.reserve(k + 1);
    for (int j = 0; j < n; j++) {
      const double dx = pt[0] - points[j][0], dy = pt[1] - points[j][1], dz = pt[2] - points[j][2];
      const std::pair<double, size_t> c((dx * dx + dy * dy) + dz * dz, (size_t)j);
      if (best.size() == k && !(c < best.back())) continue;
      best.insert(std::upper_bound(best.begin(), best.end(), c), c);
      if (best.size() > k) best.pop_back();
    }
    for (size_t i = 0; i < best.size(); i++) { k_indices[i] = best[i].second; k_sq_dists[i] = best[i].first; }
    return best.size();
  }
};
}  // namespace gtsam_points
