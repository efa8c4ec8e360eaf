// stand-in for gtsam_points::PointCloud / PointCloudCPU and the point-cloud utilities cloud_preprocessor.cpp calls
// (gtsam_points is not installed).  The containers are plain; the LEAF algorithms are [EXT] behaviour and forward to the
// oracle's restatements (oracle/glim_oracle.c) -- what compiling cloud_preprocessor.cpp against this pins is that file's OWN
// logic: which stage runs when, with which parameters, the gates, the sort key, the crop box, the k-NN layout.
#pragma once
#include <Eigen/Core>
#include <cmath>
#include <cstdint>
#include <memory>
#include <random>
#include <vector>

extern "C" {
int go_voxelgrid_sampling(int n, const double* pts4, const double* times, const double* intensities, double resolution, double* out_pts4, double* out_times, double* out_intensities);
int go_randomgrid_sampling(int n, const double* pts4, double resolution, double rate, uint64_t seed, int* keep);
void go_knn_bruteforce(int n, const double* pts4, int k, int num_threads, int* neighbors, double* sq_dists);
}

namespace gtsam_points {

extern uint64_t g_ref_shim_seed;  // stands for the std::mt19937 stream (not reproducible across implementations: SURVEY C.2)

struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud>;
  using ConstPtr = std::shared_ptr<const PointCloud>;
  size_t num_points = 0;
  double* times = nullptr;
  Eigen::Vector4d* points = nullptr;
  double* intensities = nullptr;
  size_t size() const { return num_points; }
  virtual ~PointCloud() {}
};

struct PointCloudCPU : public PointCloud {
  using Ptr = std::shared_ptr<PointCloudCPU>;
  using ConstPtr = std::shared_ptr<const PointCloudCPU>;
  std::vector<double> times_storage, intensities_storage;
  std::vector<Eigen::Vector4d> points_storage;
  void add_times(const std::vector<double>& t) { times_storage = t; times = times_storage.data(); num_points = t.size(); }
  void add_points(const std::vector<Eigen::Vector4d>& p) { points_storage = p; points = points_storage.data(); num_points = p.size(); }
  void add_intensities(const std::vector<double>& v) { intensities_storage = v; intensities = intensities_storage.data(); }
};

inline PointCloudCPU::Ptr sample(const PointCloud::ConstPtr& f, const std::vector<int>& indices) {
  auto out = std::make_shared<PointCloudCPU>();
  std::vector<Eigen::Vector4d> p(indices.size());
  std::vector<double> t(indices.size()), it(indices.size());
  for (size_t k = 0; k < indices.size(); k++) {
    p[k] = f->points[indices[k]];
    if (f->times) t[k] = f->times[indices[k]];
    if (f->intensities) it[k] = f->intensities[indices[k]];
  }
  if (f->times) out->add_times(t);
  out->add_points(p);
  if (f->intensities) out->add_intensities(it);
  return out;
}

template <typename Pred> PointCloudCPU::Ptr filter(const PointCloud::ConstPtr& f, const Pred& pred) {  // keeps the order
  std::vector<int> idx;
  for (size_t i = 0; i < f->size(); i++)
    if (pred(f->points[i])) idx.push_back((int)i);
  return sample(f, idx);
}

inline PointCloudCPU::Ptr voxelgrid_sampling(const PointCloud::ConstPtr& f, double resolution, int /*num_threads*/) {
  const int n = (int)f->size();
  std::vector<double> op(4 * (size_t)n), ot(n), oi(n);
  const int m = go_voxelgrid_sampling(n, n ? f->points[0].data() : nullptr, f->times, f->intensities, resolution, op.data(), f->times ? ot.data() : nullptr, f->intensities ? oi.data() : nullptr);
  auto out = std::make_shared<PointCloudCPU>();
  std::vector<Eigen::Vector4d> p(m);
  for (int i = 0; i < m; i++) for (int c = 0; c < 4; c++) p[i][c] = op[4 * (size_t)i + c];
  if (f->times) out->add_times(std::vector<double>(ot.begin(), ot.begin() + m));
  out->add_points(p);
  if (f->intensities) out->add_intensities(std::vector<double>(oi.begin(), oi.begin() + m));
  return out;
}

inline PointCloudCPU::Ptr randomgrid_sampling(const PointCloud::ConstPtr& f, double resolution, double rate, std::mt19937& /*mt*/, int /*num_threads*/) {
  const int n = (int)f->size();
  std::vector<int> keep(n), idx;
  go_randomgrid_sampling(n, n ? f->points[0].data() : nullptr, resolution, rate, g_ref_shim_seed, keep.data());
  for (int i = 0; i < n; i++) if (keep[i]) idx.push_back(i);
  return sample(f, idx);
}

// statistical outlier removal, the documented [EXT] rule: d_i = mean distance to the k nearest neighbours
// (query included); keep i iff d_i < mean(d) + std_thresh * stddev(d), population variance
inline PointCloudCPU::Ptr remove_outliers(const PointCloud::ConstPtr& f, int k, double std_thresh, int num_threads) {
  const int n = (int)f->size();
  std::vector<int> nb((size_t)n * k), idx;
  std::vector<double> sq((size_t)n * k), d(n);
  go_knn_bruteforce(n, n ? f->points[0].data() : nullptr, k, num_threads > 0 ? num_threads : 1, nb.data(), sq.data());
  double sum = 0.0, sum2 = 0.0;
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int j = 0; j < k; j++) {
      const Eigen::Vector4d& q = f->points[nb[(size_t)i * k + j]];
      const double ex = f->points[i][0] - q[0], ey = f->points[i][1] - q[1], ez = f->points[i][2] - q[2];
      s += std::sqrt((ex * ex + ey * ey) + ez * ez);
    }
    d[i] = s / k;
    sum += d[i];
    sum2 += d[i] * d[i];
  }
  const double mean = n ? sum / n : 0.0, var = n ? sum2 / n - mean * mean : 0.0;
  const double thresh = mean + std_thresh * std::sqrt(var > 0.0 ? var : 0.0);
  for (int i = 0; i < n; i++) if (d[i] < thresh) idx.push_back(i);
  return sample(f, idx);
}

}  // namespace gtsam_points
