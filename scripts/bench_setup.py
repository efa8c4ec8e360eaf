#!/usr/bin/env python
"""Per-frame "setup" timings on the GPU box (SURVEY 8(f) row 1 and the preprocess rows a9/a10), next to the CPU oracle:
cloud upload, voxel-map build per level, overlap call, voxel-grid downsampling, k-NN, covariance estimation.
Wall-clock per call through the C-ABI with HOST buffers (these calls synchronise), median of `reps`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glim_b200 import gpu, preprocess, synth  # noqa: E402
from oracle import oracle  # noqa: E402


def med(fn, reps=7):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def main():
    ctx = gpu.Context(0)
    sc = synth.make_hall_scene()
    traj = synth.arc_trajectory(8)
    out = {}
    for sensor, res in (("hdl32", (0.25, 0.5)), ("os1_64", (0.25, 0.5)), ("mid360", (0.1, 0.2))):
        pts, tms = synth.scan(sc, sensor, traj[3], synth.rng_for(700), backend="torch")
        n = len(pts)
        r = {"points": n}
        nb = preprocess.find_neighbors(pts, 10, ctx=ctx)
        nrm, cov = preprocess.CloudCovarianceEstimation(ctx=ctx).estimate(pts, nb)
        r["gpu_find_neighbors_ms"] = med(lambda: preprocess.find_neighbors(pts, 10, ctx=ctx), 3)
        r["gpu_covariances_ms"] = med(lambda: preprocess.CloudCovarianceEstimation(ctx=ctx).estimate(pts, nb), 5)
        r["gpu_voxelgrid_ms"] = med(lambda: preprocess.voxelgrid_sampling(pts, 0.1, times=tms, ctx=ctx), 5)
        c16 = np.ascontiguousarray(np.swapaxes(cov, 1, 2)).reshape(n, 16)
        holder = {}
        # the fused device-resident frame pipeline (gb_preprocess): voxel grid 0.1 m -> gates -> time order -> k-NN -> covariances -> device cloud
        fp = preprocess.FramePreprocessorGPU(preprocess.CloudPreprocessorParams(distance_near_thresh=0.5, distance_far_thresh=100.0, downsample_resolution=0.1, k_correspondences=10), ctx)
        r["gpu_preprocess_fused_device_cloud_ms"] = med(lambda: fp.preprocess(0.0, tms, pts, host_outputs=False), 7)
        r["gpu_preprocess_fused_with_host_products_ms"] = med(lambda: fp.preprocess(0.0, tms, pts, host_outputs=True), 5)
        r["frame_points_after_preprocess"] = int(fp.preprocess(0.0, tms, pts, host_outputs=False)[3].size())

        def up():
            holder["c"] = gpu.PointCloudGPU.clone(pts, cov, ctx=ctx)

        r["gpu_upload_ms"] = med(up)
        cloud = holder["c"]
        for rr in res:
            r[f"gpu_voxelmap_build_{rr}_ms"] = med(lambda: gpu.GaussianVoxelMapGPU(rr, ctx=ctx).insert(cloud))
        m = gpu.GaussianVoxelMapGPU(res[1], ctx=ctx).insert(cloud)
        r["voxels_buckets"] = [m.num_voxels, m.num_buckets]
        T = np.eye(4)
        r["gpu_overlap_ms"] = med(lambda: gpu.overlap_gpu(m, cloud, T), 11)
        # CPU oracle twins (all host threads)
        thr = oracle.num_threads()
        r["cpu_threads"] = thr
        if n <= 140_000:
            r["cpu_covariances_ms"] = med(lambda: oracle.covariance_estimate(pts, nb.reshape(n, 10), num_threads=thr), 3)
        r["cpu_voxelgrid_ms"] = med(lambda: oracle.voxelgrid_sampling(pts, 0.1, times=tms), 3)
        xyz, cov6 = oracle.pack_cloud(pts, c16)
        r["cpu_gpumap_build_ms"] = med(lambda: oracle.GpuMap(xyz, cov6, res[0]), 3)
        cm_t = []
        for _ in range(3):
            cm = oracle.CpuMap(res[0])
            t0 = time.perf_counter()
            cm.insert(pts, c16)
            cm_t.append(time.perf_counter() - t0)
        r["cpu_voxelmap_cpu_insert_ms"] = float(np.median(cm_t)) * 1e3
        from scipy.spatial import cKDTree

        t0 = time.perf_counter()
        cKDTree(pts[:, :3]).query(pts[:, :3], k=10, workers=-1)
        r["cpu_ckdtree_knn_ms"] = (time.perf_counter() - t0) * 1e3
        out[sensor] = r
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
