"""Host-side mirror of GLIM's per-frame preprocess (the part of the hot path that is in-reference):

    CloudPreprocessor::preprocess_impl      src/glim/preprocess/cloud_preprocessor.cpp:92-188
    CloudPreprocessor::find_neighbors       src/glim/preprocess/cloud_preprocessor.cpp:190-221
    CloudCovarianceEstimation::estimate     src/glim/common/cloud_covariance_estimation.cpp:43-122

The heavy steps (voxel-grid downsampling, exact k-NN, covariance + normal estimation) run in the CUDA
kernels of libglim_b200.so; the cheap, order-defining steps (range gate, time sort) stay on the host
exactly where the reference has them.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from .capi import check, f64, lib, ptr
from .gpu import Context, default_context


@dataclasses.dataclass
class CloudPreprocessorParams:
    """Defaults of config/config_preprocess.json (read at cloud_preprocessor.cpp:28-71), except that the
    downsampling defaults to the deterministic voxel grid BASELINE.json names."""
    distance_near_thresh: float = 0.5
    distance_far_thresh: float = 100.0
    global_shutter: bool = False
    use_random_grid_downsampling: bool = False
    downsample_resolution: float = 0.15
    k_correspondences: int = 10
    num_threads: int = 2


@dataclasses.dataclass
class PreprocessedFrame:
    """include/glim/preprocess/preprocessed_frame.hpp:14-38"""
    stamp: float
    scan_end_time: float
    times: np.ndarray
    intensities: np.ndarray | None
    points: np.ndarray  # (N,4)
    k_neighbors: int
    neighbors: np.ndarray  # (N*k,) row-major neighbors[i*k + j]

    def size(self):
        return self.points.shape[0]


def voxelgrid_sampling(points, resolution, times=None, intensities=None, ctx: Context | None = None):
    """gtsam_points::voxelgrid_sampling (cloud_preprocessor.cpp:108)."""
    ctx = ctx or default_context()
    points = f64(points)
    n = points.shape[0]
    t = f64(times) if times is not None else None
    it = f64(intensities) if intensities is not None else None
    op = np.empty((n, 4))
    ot = np.empty((n,)) if t is not None else None
    oi = np.empty((n,)) if it is not None else None
    m = C.c_size_t()
    check(lib().gb_voxelgrid_sampling(ctx.h, n, ptr(points), ptr(t), ptr(it), float(resolution), ptr(op), ptr(ot), ptr(oi), C.byref(m)))
    m = m.value
    return op[:m].copy(), (ot[:m].copy() if ot is not None else None), (oi[:m].copy() if oi is not None else None)


def find_neighbors(points, k, ctx: Context | None = None) -> np.ndarray:
    """CloudPreprocessor::find_neighbors (cloud_preprocessor.cpp:190-221): (N*k,) int32, query included."""
    ctx = ctx or default_context()
    points = f64(points)
    n = points.shape[0]
    nb = np.empty((n * k,), np.int32)
    check(lib().gb_find_neighbors(ctx.h, n, ptr(points), k, ptr(nb)))
    return nb


class CloudCovarianceEstimation:
    """glim::CloudCovarianceEstimation with RegularizationMethod::PLANE (cloud_covariance_estimation.cpp:20)."""

    def __init__(self, num_threads: int = 1, ctx: Context | None = None):
        self.num_threads = num_threads
        self.ctx = ctx

    def estimate(self, points, neighbors, k_neighbors: int | None = None):
        """-> normals (N,4), covs (N,4,4) [i,row,col]   (cloud_covariance_estimation.cpp:24-41 / :43-122)."""
        ctx = self.ctx or default_context()
        points = f64(points)
        n = points.shape[0]
        if n == 0:  # :30-32
            return np.zeros((0, 4)), np.zeros((0, 4, 4))
        neighbors = np.ascontiguousarray(neighbors, dtype=np.int32).reshape(-1)
        kc = neighbors.size // n
        if kc * n != neighbors.size:  # :34-38: spdlog::critical + abort in the reference
            raise ValueError("k * points.size() != neighbors.size()")
        k = kc if k_neighbors is None else k_neighbors
        normals = np.empty((n, 4))
        covs = np.empty((n, 16))
        check(lib().gb_covariances(ctx.h, n, ptr(points), ptr(neighbors), kc, k, ptr(normals), ptr(covs)))
        return normals, covs.reshape(n, 4, 4).transpose(0, 2, 1).copy()


class CloudPreprocessor:
    """glim::CloudPreprocessor (voxel-grid path)."""

    def __init__(self, params: CloudPreprocessorParams | None = None, ctx: Context | None = None):
        self.params = params or CloudPreprocessorParams()
        self.ctx = ctx

    def preprocess(self, stamp: float, times, points, intensities=None) -> PreprocessedFrame:
        p = self.params
        ctx = self.ctx or default_context()
        if p.use_random_grid_downsampling:
            raise NotImplementedError("randomgrid_sampling draws from std::mt19937 and is not reproducible across implementations (SURVEY C.2); use the voxel grid")
        # downsampling (:104-109)
        pts, tms, ints = voxelgrid_sampling(points, p.downsample_resolution, times, intensities, ctx)
        # distance filter (:116-128)
        sq = np.einsum("ij,ij->i", pts[:, :3], pts[:, :3])
        keep = (sq > p.distance_near_thresh**2) & (sq < p.distance_far_thresh**2) & np.isfinite(pts).all(axis=1)
        idx = np.nonzero(keep)[0]
        # sort by time (:135-136; std::sort is not stable in the reference, ties are unspecified -- we use a stable sort)
        idx = idx[np.argsort(tms[idx], kind="stable")]
        pts, tms = pts[idx], tms[idx]
        ints = ints[idx] if ints is not None else None
        if p.global_shutter:  # :138-140
            tms = np.zeros_like(tms)
        scan_end = stamp + (tms[-1] if len(tms) else 0.0)  # :174
        nb = find_neighbors(pts, p.k_correspondences, ctx)  # :182-183
        return PreprocessedFrame(stamp, scan_end, tms, ints, np.ascontiguousarray(pts), p.k_correspondences, nb)


def _poses16(poses):
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    return np.ascontiguousarray(np.swapaxes(poses, 1, 2)).reshape(-1, 16)


def deskew_pose_table(T_imu_lidar, times, linear_vel=None, angular_vel=None, imu_times=None, imu_poses=None, stamp=0.0):
    """Host half of CloudDeskewing::deskew: (time_indices (N,), table poses (M,4,4)).  Needs no GPU."""
    from .capi import pose16

    times = f64(times)
    n = times.shape[0]
    idx = np.empty((n,), np.int32)
    table = np.empty((max(n, 1), 16))
    m = C.c_size_t()
    it = f64(imu_times) if imu_times is not None else None
    ip = _poses16(imu_poses) if imu_poses is not None else None
    lv = f64(linear_vel) if linear_vel is not None else None
    av = f64(angular_vel) if angular_vel is not None else None
    check(lib().gb_deskew_pose_table(ptr(pose16(T_imu_lidar)), ptr(lv), ptr(av), 0 if it is None else len(it), ptr(it), ptr(ip), float(stamp), n, ptr(times), ptr(idx), ptr(table), C.byref(m)))
    return idx, np.swapaxes(table[: m.value].reshape(-1, 4, 4), 1, 2).copy()


class CloudDeskewing:
    """glim::CloudDeskewing (src/glim/common/cloud_deskewing.cpp): both overloads of deskew(), plus the optional fused second
    transform of the call site (odometry_estimation_imu.cpp:313-316)."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx

    def deskew(self, T_imu_lidar, times, points, linear_vel=None, angular_vel=None, imu_times=None, imu_poses=None, stamp=0.0, T_post=None):
        from .capi import pose16

        ctx = self.ctx or default_context()
        times, points = f64(times), f64(points)
        n = points.shape[0]
        if n == 0:
            return np.zeros((0, 4))
        out = np.empty_like(points)
        it = f64(imu_times) if imu_times is not None else None
        ip = _poses16(imu_poses) if imu_poses is not None else None
        lv = f64(linear_vel) if linear_vel is not None else None
        av = f64(angular_vel) if angular_vel is not None else None
        tp = pose16(T_post) if T_post is not None else None
        check(lib().gb_deskew(ctx.h, ptr(pose16(T_imu_lidar)), ptr(lv), ptr(av), 0 if it is None else len(it), ptr(it), ptr(ip), float(stamp), n, ptr(times), ptr(points), ptr(tp), ptr(out)))
        return out
