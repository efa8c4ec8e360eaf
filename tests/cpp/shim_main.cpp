// shim_main.cpp -- drives the hot path exactly the way GLIM's modules do, through the C++ shims of
// include/glim_b200/gtsam_points_compat.hpp (create_frame: odometry_estimation_gpu.cpp:86-107; create_factors: :128-206;
// batched linearization: :383-386; keyframe overlap: :231).  Input / output are raw binary files exchanged with
// tests/test_cpp_shim.py, which checks the numbers against the oracle.
//   in : int32 n0, n1 | n0 x 4 f64 pts | n0 x 16 f64 covs | n1 x 4 f64 pts | n1 x 16 f64 covs | 16 f64 T_target | 16 f64 T_source
//   out: 4 x gb_linearized6 (binary L0, binary L1, unary L0, unary L1) | f64 error(unary L1) | f64 overlap | i32 num_voxels[2] num_buckets[2]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "glim_b200/gtsam_points_compat.hpp"

using namespace gtsam_points;

static void read_all(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: shim_main in.bin out.bin\n"); return 2; }
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) return 2;
  int n[2];
  read_all(fi, n, sizeof(n));
  std::vector<double> pts[2], covs[2];
  for (int k = 0; k < 2; k++) {
    pts[k].resize(4 * (size_t)n[k]); covs[k].resize(16 * (size_t)n[k]);
    read_all(fi, pts[k].data(), sizeof(double) * pts[k].size());
    read_all(fi, covs[k].data(), sizeof(double) * covs[k].size());
  }
  glim_b200::Pose Tt, Ts;
  read_all(fi, Tt.m.data(), sizeof(double) * 16);
  read_all(fi, Ts.m.data(), sizeof(double) * 16);
  fclose(fi);

  try {
    CUDAStream stream;                               // odometry_estimation_gpu.cpp:76
    StreamTempBufferRoundRobin roundrobin;           // :77
    PointCloud::ConstPtr frames[2];
    double md = 0.0;
    for (int k = 0; k < 2; k++) {
      auto host = std::make_shared<PointCloudCPU>();
      host->add_points(pts[k].data(), (size_t)n[k]);
      host->add_covs(covs[k].data(), (size_t)n[k]);
      frames[k] = host;
      md = median_distance(frames[k], 256);                          // :91
      frames[k] = PointCloudGPU::clone(*frames[k]);                  // :96  (the host frame dies here: clone owns a copy)
      if (md != median_distance(frames[k], 256)) throw std::runtime_error("clone changed the host data");
    }
    // create_frame: voxelmap_levels maps with resolution * scaling^level   (:97-106)
    std::vector<GaussianVoxelMap::Ptr> voxelmaps;
    for (int level = 0; level < 2; level++) {
      auto vm = std::make_shared<GaussianVoxelMapGPU>(0.25f * (float)(1 << level), 8192 * 2, 10, 1e-3, stream);
      vm->insert(*frames[0]);
      voxelmaps.push_back(vm);
    }
    // create_factors: one binary and one unary factor per voxel map   (:143-147, :160-164)
    auto sb = roundrobin.get_stream_buffer();
    std::vector<std::shared_ptr<IntegratedVGICPFactorGPU>> graph;
    for (auto& vm : voxelmaps) {
      auto f = std::make_shared<IntegratedVGICPFactorGPU>(Key(0), Key(1), vm, frames[1], sb.first, sb.second);
      f->set_enable_surface_validation(false);
      graph.push_back(f);
    }
    for (auto& vm : voxelmaps) graph.push_back(std::make_shared<IntegratedVGICPFactorGPU>(Tt, Key(1), vm, frames[1], sb.first, sb.second));
    Values values;
    values[0] = Tt;
    values[1] = Ts;
    NonlinearFactorSetGPU set;                       // :383-385
    set.add(graph);
    set.linearize(values);
    const double err = graph[3]->error(values);
    const double ov = overlap_gpu(std::vector<GaussianVoxelMap::ConstPtr>{voxelmaps[1]}, frames[1], std::vector<glim_b200::Pose>{Tt.inverse() * Ts}, stream);   // :231

    FILE* fo = fopen(argv[2], "wb");
    fwrite(set.results().data(), sizeof(gb_linearized6), 4, fo);
    fwrite(&err, sizeof(double), 1, fo);
    fwrite(&ov, sizeof(double), 1, fo);
    for (int l = 0; l < 2; l++) { auto vm = std::static_pointer_cast<GaussianVoxelMapGPU>(voxelmaps[l]); fwrite(&vm->voxelmap_info.num_voxels, sizeof(int), 1, fo); }
    for (int l = 0; l < 2; l++) { auto vm = std::static_pointer_cast<GaussianVoxelMapGPU>(voxelmaps[l]); fwrite(&vm->voxelmap_info.num_buckets, sizeof(int), 1, fo); }
    fclose(fo);
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_main: %s\n", e.what());
    return 1;
  }
  return 0;
}
