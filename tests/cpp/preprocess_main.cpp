// preprocess_main.cpp -- the per-frame path as GLIM runs it, through the shims of include/glim_b200/glim_preprocess_compat.hpp:
// CloudPreprocessor::preprocess (cloud_preprocessor.cpp:77-188) -> CloudCovarianceEstimation::estimate
// (odometry_estimation_imu.cpp:322-328) -> PointCloudGPU::clone (odometry_estimation_gpu.cpp:96), and the fused
// preprocess_to_gpu_frame; then merge_frames_gpu of two such frames (sub_mapping.cpp:491).  Results go to tests/test_cpp_shim.py.
//   in : int32 n | n x 4 f64 pts | n f64 times | f64 resolution near far | int32 k
//   out: int32 m | m x 4 f64 pts | m f64 times | m x k int32 neighbors | m x 16 f64 covs (two-step) | m x 16 f64 covs (fused) | int32 merged | int32 fused_has_gpu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "glim_b200/glim_preprocess_compat.hpp"

using namespace glim_b200::glim;

static void read_all(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) return 2;
  int n = 0, k = 0;
  read_all(fi, &n, sizeof(n));
  auto raw = std::make_shared<RawPoints>();
  raw->stamp = 100.0;
  raw->points.resize((size_t)n);
  raw->times.resize((size_t)n);
  read_all(fi, raw->points.data(), sizeof(double) * 4 * (size_t)n);
  read_all(fi, raw->times.data(), sizeof(double) * (size_t)n);
  double cfg[3];
  read_all(fi, cfg, sizeof(cfg));
  read_all(fi, &k, sizeof(k));
  fclose(fi);
  try {
    gtsam_points::CUDAStream stream;
    CloudPreprocessorParams params;
    params.downsample_resolution = cfg[0];
    params.distance_near_thresh = cfg[1];
    params.distance_far_thresh = cfg[2];
    params.k_correspondences = k;
    CloudPreprocessor preprocessor(params, stream);
    PreprocessedFrame::Ptr frame = preprocessor.preprocess(raw);                       // async_odometry / preprocess module
    CloudCovarianceEstimation covariance_estimation(2, stream);
    std::vector<Vector4d> normals;
    std::vector<Matrix4d> covs;
    covariance_estimation.estimate(frame->points, frame->neighbors, normals, covs);    // odometry_estimation_imu.cpp:322-328
    auto host = std::make_shared<gtsam_points::PointCloudCPU>();
    host->add_points(reinterpret_cast<const double*>(frame->points.data()), frame->points.size());
    host->add_covs(reinterpret_cast<const double*>(covs.data()), covs.size());
    host->add_normals(reinterpret_cast<const double*>(normals.data()), normals.size());
    gtsam_points::PointCloud::ConstPtr f0 = gtsam_points::PointCloudGPU::clone(*host, stream);   // odometry_estimation_gpu.cpp:96
    // the fused path: the same frame in one device-resident call
    PreprocessedFrame::Ptr frame2;
    gtsam_points::PointCloudGPU::Ptr f1 = preprocessor.preprocess_to_gpu_frame(raw, &frame2);
    if (frame2->points.size() != frame->points.size() || memcmp(frame2->points.data(), frame->points.data(), sizeof(Vector4d) * frame->points.size()) != 0) throw std::runtime_error("fused frame differs");
    // sub-mapping: merge the two (identical) keyframes, one shifted by 10 cm
    glim_b200::Pose shift;
    shift(0, 3) = 0.1;
    auto merged = gtsam_points::merge_frames_gpu({glim_b200::Pose(), shift}, {f0, f1}, 0.25, 0, stream);
    const int m = frame->size(), mm = (int)merged->size(), has_gpu = f1->points_gpu != nullptr && f1->covs_gpu != nullptr && f1->has_normals();
    FILE* fo = fopen(argv[2], "wb");
    fwrite(&m, sizeof(int), 1, fo);
    fwrite(frame->points.data(), sizeof(Vector4d), (size_t)m, fo);
    fwrite(frame->times.data(), sizeof(double), (size_t)m, fo);
    fwrite(frame->neighbors.data(), sizeof(int), (size_t)m * (size_t)k, fo);
    fwrite(covs.data(), sizeof(Matrix4d), (size_t)m, fo);
    fwrite(f1->covs, sizeof(Matrix4d), (size_t)m, fo);
    fwrite(&mm, sizeof(int), 1, fo);
    fwrite(&has_gpu, sizeof(int), 1, fo);
    fclose(fo);
  } catch (const std::exception& e) {
    fprintf(stderr, "preprocess_main: %s\n", e.what());
    return 1;
  }
  return 0;
}
