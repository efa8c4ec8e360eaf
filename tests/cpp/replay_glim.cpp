// replay_glim.cpp -- replays, statement for statement, what GLIM's three GPU modules do with the gtsam_points surface, each on
// its own thread with its own stream objects, handing frames from one module to the next as GLIM's async wrappers do
// (async_odometry_estimation.cpp:15 -> async_sub_mapping.cpp:8 -> async_global_mapping.cpp:24):
//
//   odometry thread     odometry_estimation_gpu.cpp:76-77 (CUDAStream + StreamTempBufferRoundRobin members), :91 median_distance,
//                       :96  new_frame->frame = PointCloudGPU::clone(*new_frame->frame)   <- the ONLY owner is replaced by its clone
//                       :103-104 GaussianVoxelMapGPU(res, 8192*2, 10, 1e-3, *stream)->insert(*frame)
//                       :139-147 binary factors with (stream, buffer) from the round robin, :383-385 NonlinearFactorSetGPU
//   sub-mapping thread  sub_mapping.cpp:165-169  if (!frame->points_gpu) frame = clone(*frame, *stream);  :296-307 factors on ITS stream
//                       between voxel maps / frames that the ODOMETRY thread uploaded; :252 overlap_auto
//   global-mapping thr. global_mapping.cpp:239 median_distance(submap->frame) AFTER :253 submap->frame = clone(*submap->frame)
//                       (the host data must have survived), :265 GaussianVoxelMapGPU(resolution) 1-argument form,
//                       :322 overlap_auto, :330 points_gpu test, :335 factors with its own round robin
//
// Built with -fsanitize=address by tests/test_cpp_shim.py: a clone that aliases the caller's arrays (round 1) is a
// heap-use-after-free here.  Results of the three modules are written out and compared (same pair, same pose => same
// blocks whichever thread / stream linearized it), and against the oracle by the Python side.
//   in : int32 n0, n1 | n0 x 4 f64 pts | n0 x 16 f64 covs | n1 x 4 f64 pts | n1 x 16 f64 covs | 16 f64 T_target | 16 f64 T_source
//   out: 3 modules x 2 levels gb_linearized6 | 3 x f64 overlap | f64 median_distance before / after clone
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>
#include <vector>

#include "glim_b200/gtsam_points_compat.hpp"

using namespace gtsam_points;

static void read_all(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

struct EstimationFrame {  // include/glim/odometry/estimation_frame.hpp:103-105
  PointCloud::ConstPtr frame;
  std::vector<GaussianVoxelMap::Ptr> voxelmaps;
};

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: replay_glim in.bin out.bin\n"); return 2; }
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) return 2;
  int n[2];
  read_all(fi, n, sizeof(n));
  glim_b200::Pose Tt, Ts;
  std::vector<std::shared_ptr<EstimationFrame>> frames(2);
  double md_before[2] = {0, 0};
  {
    // the preprocessed frames as GLIM's preprocessing hands them over: PointCloudCPU owning its arrays; the raw buffers we
    // read them into are freed right away
    for (int k = 0; k < 2; k++) {
      std::vector<double> pts(4 * (size_t)n[k]), covs(16 * (size_t)n[k]);
      read_all(fi, pts.data(), sizeof(double) * pts.size());
      read_all(fi, covs.data(), sizeof(double) * covs.size());
      auto host = std::make_shared<PointCloudCPU>();
      host->add_points(pts.data(), (size_t)n[k]);
      host->add_covs(covs.data(), (size_t)n[k]);
      frames[k] = std::make_shared<EstimationFrame>();
      frames[k]->frame = host;
    }
    read_all(fi, Tt.m.data(), sizeof(double) * 16);
    read_all(fi, Ts.m.data(), sizeof(double) * 16);
    fclose(fi);
  }
  Values values;
  values[0] = Tt;
  values[1] = Ts;
  const glim_b200::Pose delta = Tt.inverse() * Ts;
  std::vector<gb_linearized6> results(6);
  double overlaps[3] = {0, 0, 0};
  double md_after = 0.0;
  int rc = 0;

  try {
    // ------------------------------------------------------------------ odometry thread
    std::thread odometry([&] {
      try {
        auto stream = std::make_unique<CUDAStream>();                          // :76
        auto stream_buffer_roundrobin = std::make_unique<StreamTempBufferRoundRobin>(8);  // :77
        for (int k = 0; k < 2; k++) {
          auto& new_frame = frames[k];
          md_before[k] = median_distance(new_frame->frame, 256);               // :91
          new_frame->frame = PointCloudGPU::clone(*new_frame->frame);          // :96  -- the PointCloudCPU dies HERE
          for (int i = 0; i < 2; i++) {
            if (!new_frame->frame->size()) break;                              // :98
            auto voxelmap = std::make_shared<GaussianVoxelMapGPU>(0.25 * (1 << i), 8192 * 2, 10, 1e-3, *stream);  // :103
            voxelmap->insert(*new_frame->frame);                               // :104
            new_frame->voxelmaps.push_back(voxelmap);
          }
        }
        std::vector<std::shared_ptr<IntegratedVGICPFactorGPU>> graph;
        const auto sb = stream_buffer_roundrobin->get_stream_buffer();          // :139
        for (const auto& voxelmap : frames[0]->voxelmaps) {
          auto factor = std::make_shared<IntegratedVGICPFactorGPU>(Key(0), Key(1), voxelmap, frames[1]->frame, sb.first, sb.second);  // :144
          factor->set_enable_surface_validation(false);
          graph.push_back(factor);
        }
        NonlinearFactorSetGPU set;                                             // :383-385
        set.add(graph);
        set.linearize(values);
        results[0] = set.results()[0];
        results[1] = set.results()[1];
        overlaps[0] = overlap_gpu(frames[0]->voxelmaps.back(), frames[1]->frame, delta, *stream);  // :248
        // the module's stream objects die with the thread; the frames it uploaded live on in the next module
      } catch (const std::exception& e) { fprintf(stderr, "odometry: %s\n", e.what()); rc = 1; }
    });
    odometry.join();
    if (rc) return rc;

    // ------------------------------------------------------------------ sub-mapping thread
    std::thread sub_mapping([&] {
      try {
        auto stream = std::make_shared<CUDAStream>();                          // sub_mapping.cpp:86
        auto stream_buffer_roundrobin = std::make_shared<StreamTempBufferRoundRobin>(8);  // :87
        for (auto& odom_frame : frames) {
          if (!odom_frame->frame->points_gpu) {                                // :165 (false: the odometry module uploaded them)
            auto frame_gpu = PointCloudGPU::clone(*odom_frame->frame, *stream);  // :168
            odom_frame->frame = frame_gpu;
          }
        }
        // a keyframe re-uploaded on THIS module's stream (sub_mapping.cpp:393-399), used together with a frame of the other module
        auto keyframe = std::make_shared<EstimationFrame>();
        keyframe->frame = PointCloudGPU::clone(*frames[0]->frame, *stream);    // :393
        for (int i = 0; i < 2; i++) {
          auto voxelmap = std::make_shared<GaussianVoxelMapGPU>(0.25 * (1 << i), 8192 * 2, 10, 1e-3, *stream);  // :398
          voxelmap->insert(*keyframe->frame);                                  // :399
          keyframe->voxelmaps.push_back(voxelmap);
        }
        overlaps[1] = overlap_auto(keyframe->voxelmaps.back(), frames[1]->frame, delta);  // :252
        const auto sb = stream_buffer_roundrobin->get_stream_buffer();          // :296
        std::vector<std::shared_ptr<IntegratedVGICPFactorGPU>> graph;
        // one factor on this module's own voxel map, one on the voxel map the odometry thread built: a MIXED factor set
        graph.push_back(std::make_shared<IntegratedVGICPFactorGPU>(Key(0), Key(1), keyframe->voxelmaps[0], frames[1]->frame, sb.first, sb.second));  // :307
        graph.push_back(std::make_shared<IntegratedVGICPFactorGPU>(Key(0), Key(1), frames[0]->voxelmaps[1], frames[1]->frame, sb.first, sb.second));
        NonlinearFactorSetGPU set;
        set.add(graph);
        set.linearize(values);
        results[2] = set.results()[0];
        results[3] = set.results()[1];
      } catch (const std::exception& e) { fprintf(stderr, "sub_mapping: %s\n", e.what()); rc = 1; }
    });
    sub_mapping.join();
    if (rc) return rc;

    // ------------------------------------------------------------------ global-mapping thread
    std::thread global_mapping([&] {
      try {
        auto stream_buffer_roundrobin = std::make_shared<StreamTempBufferRoundRobin>(64);  // global_mapping.cpp:110
        // submaps arrive with CPU frames (merge_frames output): rebuild a CPU-only copy of frame 0 / 1 to replay :252-253
        std::vector<std::shared_ptr<EstimationFrame>> submaps(2);
        for (int k = 0; k < 2; k++) {
          submaps[k] = std::make_shared<EstimationFrame>();
          submaps[k]->frame = std::make_shared<PointCloudCPU>(*frames[k]->frame);   // host copy only: points_gpu == nullptr
          const double dist_median = median_distance(submaps[k]->frame, 256);  // :239
          (void)dist_median;
          if (!submaps[k]->frame->points_gpu) submaps[k]->frame = PointCloudGPU::clone(*submaps[k]->frame);  // :252-253 (no stream)
          for (int i = 0; i < 2; i++) {
            auto voxelmap = std::make_shared<GaussianVoxelMapGPU>(0.25 * (1 << i));  // :265 (1-argument form)
            voxelmap->insert(*submaps[k]->frame);                              // :266
            submaps[k]->voxelmaps.push_back(voxelmap);
          }
        }
        md_after = median_distance(submaps[1]->frame, 256);                     // host data of a clone of a clone
        overlaps[2] = overlap_auto(submaps[0]->voxelmaps.back(), submaps[1]->frame, delta);  // :322
        std::vector<std::shared_ptr<IntegratedVGICPFactorGPU>> graph;
        if (std::dynamic_pointer_cast<GaussianVoxelMapGPU>(submaps[0]->voxelmaps.back()) && submaps[1]->frame->points_gpu) {  // :330
          const auto stream_buffer = stream_buffer_roundrobin->get_stream_buffer();
          for (const auto& voxelmap : submaps[0]->voxelmaps) graph.push_back(std::make_shared<IntegratedVGICPFactorGPU>(Key(0), Key(1), voxelmap, submaps[1]->frame, stream_buffer.first, stream_buffer.second));  // :335
        }
        NonlinearFactorSetGPU set;
        set.add(graph);
        set.linearize(values);
        results[4] = set.results()[0];
        results[5] = set.results()[1];
        // frames of the first two modules are released here, on a thread that never uploaded them
        frames.clear();
      } catch (const std::exception& e) { fprintf(stderr, "global_mapping: %s\n", e.what()); rc = 1; }
    });
    global_mapping.join();
    if (rc) return rc;
  } catch (const std::exception& e) {
    fprintf(stderr, "replay_glim: %s\n", e.what());
    return 1;
  }

  FILE* fo = fopen(argv[2], "wb");
  fwrite(results.data(), sizeof(gb_linearized6), results.size(), fo);
  fwrite(overlaps, sizeof(double), 3, fo);
  fwrite(&md_before[1], sizeof(double), 1, fo);
  fwrite(&md_after, sizeof(double), 1, fo);
  fclose(fo);
  return 0;
}
