// host_helpers_main.cpp -- the host-side gtsam_points helpers GLIM's GPU modules call around the device classes
// (sub_mapping.cpp:385 / global_mapping.cpp:248 random_sampling before clone; offline_viewer.cpp:29 factor-set hook;
// standard_viewer_mem.cpp:77 VoxelBucket).  No device call: runs on the CPU-only box.
//   out: int32 n_sampled | sampled indices recovered from the x coordinate | 1 byte "all attributes consistent"
#include <cstdio>
#include <random>
#include <vector>

#include "glim_b200/gtsam_points_compat.hpp"

using namespace gtsam_points;

int main(int argc, char** argv) {
  if (argc != 4) { fprintf(stderr, "usage: host_helpers n rate out.bin\n"); return 2; }
  const int n = atoi(argv[1]);
  const double rate = atof(argv[2]);
  std::vector<double> pts(4 * (size_t)n), covs(16 * (size_t)n, 0.0), times(n), inten(n);
  for (int i = 0; i < n; i++) {
    pts[4 * i] = i; pts[4 * i + 1] = 2.0 * i; pts[4 * i + 2] = -1.0 * i; pts[4 * i + 3] = 1.0;
    covs[16 * (size_t)i] = 100.0 + i;
    times[i] = 0.001 * i;
    inten[i] = 7.0 + i;
  }
  auto frame = std::make_shared<PointCloudCPU>();
  frame->add_points(pts.data(), n);
  frame->add_covs(covs.data(), n);
  frame->add_times(times.data(), n);
  frame->add_intensities(inten.data(), n);

  std::mt19937 mt(8192);  // sub_mapping.cpp / global_mapping.cpp keep a std::mt19937 member
  PointCloud::ConstPtr sub = random_sampling(frame, rate, mt);
  bool ok = sub->has_points() && sub->has_covs() && sub->has_times() && sub->has_intensities() && !sub->has_normals() && !sub->points_gpu;
  std::vector<int> idx;
  for (size_t k = 0; k < sub->size(); k++) {
    const int i = (int)sub->points[k][0];
    idx.push_back(i);
    ok = ok && sub->points[k][1] == 2.0 * i && sub->covs[k](0, 0) == 100.0 + i && sub->times[k] == 0.001 * i && sub->intensities[k] == 7.0 + i;
    if (k) ok = ok && idx[k] > idx[k - 1];
  }
  // rate >= 1: the frame itself (deep copy), as the callers expect a frame of the same size
  std::mt19937 mt2(1);
  ok = ok && random_sampling(frame, 1.0, mt2)->size() == (size_t)n;
  // the same generator state gives the same draw (the reference's reproducibility contract)
  std::mt19937 a(8192), b(8192);
  auto s1 = random_sampling(frame, rate, a), s2 = random_sampling(frame, rate, b);
  ok = ok && s1->size() == s2->size();
  for (size_t k = 0; ok && k < s1->size(); k++) ok = s1->points[k][0] == s2->points[k][0];
  // factor-set hook (offline_viewer.cpp:29)
  LinearizationHook::register_hook([] { return create_nonlinear_factor_set_gpu(); });
  const auto sets = LinearizationHook::create_factor_sets();
  ok = ok && sets.size() == 1 && sets[0] && sets[0]->size() == 0;
  ok = ok && sizeof(VoxelBucket) == 16;
  // merge_frames (sub_mapping.cpp:496) refuses host-only frames instead of falling back to a CPU path
  try {
    merge_frames({glim_b200::Pose()}, {frame}, 0.5, 0);
    ok = false;
  } catch (const std::runtime_error&) {
  }

  FILE* fo = fopen(argv[3], "wb");
  const int m = (int)idx.size();
  fwrite(&m, sizeof(int), 1, fo);
  fwrite(idx.data(), sizeof(int), idx.size(), fo);
  const unsigned char okb = ok ? 1 : 0;
  fwrite(&okb, 1, 1, fo);
  fclose(fo);
  return ok ? 0 : 1;
}
