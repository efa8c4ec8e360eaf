// deskew_host_main.cpp -- the host half of the glim::CloudDeskewing shim (include/glim_b200/glim_preprocess_compat.hpp), run on the
// CPU-only box: CloudDeskewing::pose_table for both overloads of deskew(), applied to the points on the host, written out for
// tests/test_cpp_shim.py, which compares with the oracle and with the reference's own cloud_deskewing.cpp (oracle/_ref).  Also
// checks that deskew() itself fails loudly without a device (no CPU fallback) and returns an empty cloud for an empty scan.
//   in : int32 n | n f64 times | n x 4 f64 pts | 16 f64 T_imu_lidar | 3 f64 linear_vel | 3 f64 angular_vel | int32 n_imu | n_imu f64 | n_imu x 16 f64 | f64 stamp
//   out: n x 4 f64 (constant velocity) | n x 4 f64 (IMU poses) | int32 slots | int32 device_call_threw
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "glim_b200/glim_preprocess_compat.hpp"

using namespace glim_b200::glim;
using glim_b200::Pose;

static void read_all(FILE* f, void* p, size_t n) { if (n && fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

static std::vector<Vector4d> apply(const std::vector<int>& idx, const std::vector<Pose>& T, const std::vector<Vector4d>& pts) {
  std::vector<Vector4d> out(pts.size());
  for (size_t i = 0; i < pts.size(); i++) {
    const Pose& M = T[(size_t)idx[i]];
    for (int r = 0; r < 4; r++) {
      double s = 0.0;
      for (int c = 0; c < 4; c++) s += M(r, c) * pts[i][c];
      out[i][r] = s;
    }
  }
  return out;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) return 2;
  int n = 0, n_imu = 0;
  read_all(fi, &n, sizeof(n));
  std::vector<double> times((size_t)n);
  std::vector<Vector4d> pts((size_t)n);
  read_all(fi, times.data(), sizeof(double) * (size_t)n);
  read_all(fi, pts.data(), sizeof(Vector4d) * (size_t)n);
  Pose T_imu_lidar;
  Vector3d lv, av;
  read_all(fi, T_imu_lidar.m.data(), sizeof(double) * 16);
  read_all(fi, lv.v, sizeof(double) * 3);
  read_all(fi, av.v, sizeof(double) * 3);
  read_all(fi, &n_imu, sizeof(n_imu));
  std::vector<double> imu_times((size_t)n_imu);
  std::vector<Pose> imu_poses((size_t)n_imu);
  read_all(fi, imu_times.data(), sizeof(double) * (size_t)n_imu);
  for (auto& p : imu_poses) read_all(fi, p.m.data(), sizeof(double) * 16);
  double stamp = 0.0;
  read_all(fi, &stamp, sizeof(stamp));
  fclose(fi);

  int threw = 0, slots = 0;
  std::vector<Vector4d> a, b;
  try {
    std::vector<int> idx;
    std::vector<Pose> tab;
    CloudDeskewing::pose_table(T_imu_lidar, &lv, &av, nullptr, nullptr, 0.0, times, idx, tab);
    a = apply(idx, tab, pts);
    slots = (int)tab.size();
    CloudDeskewing::pose_table(T_imu_lidar, nullptr, nullptr, &imu_times, &imu_poses, stamp, times, idx, tab);
    b = apply(idx, tab, pts);
    CloudDeskewing deskewing;
    if (!deskewing.deskew(T_imu_lidar, lv, av, std::vector<double>(), std::vector<Vector4d>()).empty()) return 3;  // empty scan: no device call
    if (gb_device_count() == 0) {
      try {
        deskewing.deskew(T_imu_lidar, imu_times, imu_poses, stamp, times, pts);
      } catch (const std::runtime_error&) {
        threw = 1;  // GB_ERR_NO_DEVICE surfaced as an exception: there is no CPU fallback behind the shim
      }
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "deskew_host_main: %s\n", e.what());
    return 1;
  }
  FILE* fo = fopen(argv[2], "wb");
  fwrite(a.data(), sizeof(Vector4d), a.size(), fo);
  fwrite(b.data(), sizeof(Vector4d), b.size(), fo);
  fwrite(&slots, sizeof(int), 1, fo);
  fwrite(&threw, sizeof(int), 1, fo);
  fclose(fo);
  return 0;
}
