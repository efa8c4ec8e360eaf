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
    """glim::CloudPreprocessorParams.  Field defaults are the CODE defaults of cloud_preprocessor.cpp:28-61 (what the reference
    uses when a key is missing); `from_shipped_config()` gives the values of config/config_preprocess.json:19-34 (random grid,
    1.0 m, target 10 000, k = 10, near 0.5 m), which is what GLIM runs with out of the box."""
    distance_near_thresh: float = 1.0        # :28
    distance_far_thresh: float = 100.0       # :29
    global_shutter: bool = False
    use_random_grid_downsampling: bool = False  # :30
    downsample_resolution: float = 0.15      # :31
    downsample_target: int = 0               # :32
    downsample_rate: float = 0.3             # :33
    enable_outlier_removal: bool = False     # :34
    outlier_removal_k: int = 10              # :35
    outlier_std_mul_factor: float = 2.0      # :36
    enable_cropbox_filter: bool = False      # :38
    crop_bbox_frame: str = "lidar"           # :39
    crop_bbox_min: tuple = (0.0, 0.0, 0.0)
    crop_bbox_max: tuple = (0.0, 0.0, 0.0)
    T_imu_lidar: tuple = tuple(np.eye(4).reshape(-1))
    k_correspondences: int = 8               # :59
    num_threads: int = 2                     # :61

    @staticmethod
    def from_shipped_config() -> "CloudPreprocessorParams":
        return CloudPreprocessorParams(distance_near_thresh=0.5, distance_far_thresh=100.0, use_random_grid_downsampling=True, downsample_resolution=1.0, downsample_target=10000,
                                       downsample_rate=0.1, outlier_std_mul_factor=1.0, crop_bbox_min=(-1.0, -1.0, -1.0), crop_bbox_max=(1.0, 1.0, 1.0), k_correspondences=10)


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
    """glim::CloudPreprocessor::preprocess (cloud_preprocessor.cpp:77-188) -> PreprocessedFrame.  One gb_preprocess call: the
    whole pipeline runs on the device; only the PreprocessedFrame fields come back."""

    def __init__(self, params: CloudPreprocessorParams | None = None, ctx: Context | None = None, seed: int = 0):
        self.params = params or CloudPreprocessorParams()
        self.ctx = ctx
        self.seed = seed

    def preprocess(self, stamp: float, times, points, intensities=None) -> PreprocessedFrame:
        from .capi import Preprocessed

        ctx = self.ctx or default_context()
        g = FramePreprocessorGPU(self.params, ctx, self.seed)
        cp = g.c_params()
        cp.estimate_covariances = 0
        points = f64(points)
        n = points.shape[0]
        t = f64(times) if times is not None else None
        it = f64(intensities) if intensities is not None else None
        k = cp.k_correspondences
        out = Preprocessed()
        o_t, o_p, o_i, o_n = np.empty(n), np.empty((n, 4)), (np.empty(n) if it is not None else None), np.empty((n, k), np.int32)
        out.times, out.xyzw, out.intensities, out.neighbors = o_t.ctypes.data, o_p.ctypes.data, (o_i.ctypes.data if o_i is not None else None), o_n.ctypes.data
        check(lib().gb_preprocess(ctx.h, n, ptr(points), ptr(t), ptr(it), C.byref(cp), C.byref(out)))
        m = out.num_points
        return PreprocessedFrame(stamp, stamp + out.last_time, o_t[:m].copy(), o_i[:m].copy() if o_i is not None else None, np.ascontiguousarray(o_p[:m]), k, o_n[:m].reshape(-1).copy())


class FramePreprocessorGPU:
    """gb_preprocess: CloudPreprocessor::preprocess_impl + CloudCovarianceEstimation::estimate + PointCloudGPU::clone in one
    device-resident call (include/glim_b200.h).  `params` is the CloudPreprocessorParams of this module; `host_outputs`
    selects whether the PreprocessedFrame fields / fp64 covariances are copied back (the device cloud is always produced)."""

    def __init__(self, params: "CloudPreprocessorParams | None" = None, ctx: Context | None = None, seed: int = 0, knn_cell_size: float = 0.0):
        self.params = params or CloudPreprocessorParams()
        self.ctx = ctx
        self.seed = seed
        self.knn_cell_size = knn_cell_size

    def c_params(self):
        from .capi import PreprocessParams

        p = self.params
        cp = PreprocessParams()
        check(lib().gb_preprocess_default_params(C.byref(cp)))
        cp.distance_near_thresh, cp.distance_far_thresh = p.distance_near_thresh, p.distance_far_thresh
        cp.use_random_grid_downsampling = int(p.use_random_grid_downsampling)
        cp.downsample_resolution, cp.downsample_target, cp.downsample_rate = p.downsample_resolution, p.downsample_target, p.downsample_rate
        cp.seed = self.seed
        cp.global_shutter = int(p.global_shutter)
        cp.crop_bbox_frame = {"": 0, "lidar": 1, "imu": 2}[p.crop_bbox_frame] if p.enable_cropbox_filter else 0
        for a in range(3):
            cp.crop_bbox_min[a], cp.crop_bbox_max[a] = p.crop_bbox_min[a], p.crop_bbox_max[a]
        T = np.asarray(p.T_imu_lidar, dtype=np.float64).reshape(4, 4)
        for c in range(4):
            for r in range(4):
                cp.T_imu_lidar[c * 4 + r] = T[r, c]
        cp.enable_outlier_removal = int(p.enable_outlier_removal)
        cp.outlier_removal_k, cp.outlier_std_mul_factor = p.outlier_removal_k, p.outlier_std_mul_factor
        cp.k_correspondences = p.k_correspondences
        cp.estimate_covariances = 1
        cp.knn_cell_size = self.knn_cell_size
        return cp

    def preprocess(self, stamp: float, times, points, intensities=None, host_outputs: bool = True):
        """-> (PreprocessedFrame or None, normals (M,4), covs (M,4,4), gpu.PointCloudGPU)"""
        from . import gpu
        from .capi import Preprocessed

        ctx = self.ctx or default_context()
        points = f64(points)
        n = points.shape[0]
        t = f64(times) if times is not None else None
        it = f64(intensities) if intensities is not None else None
        cp = self.c_params()
        k = cp.k_correspondences
        out = Preprocessed()
        bufs = {}
        if host_outputs:
            bufs = {"times": np.empty(n), "xyzw": np.empty((n, 4)), "intensities": np.empty(n) if it is not None else None, "neighbors": np.empty((n, k), np.int32), "normals4": np.empty((n, 4)), "cov4x4": np.empty((n, 16))}
            for name, b in bufs.items():
                setattr(out, name, b.ctypes.data if b is not None else None)
        check(lib().gb_preprocess(ctx.h, n, ptr(points), ptr(t), ptr(it), C.byref(cp), C.byref(out)))
        m = out.num_points
        cloud = gpu.PointCloudGPU(ctx, C.c_void_p(out.cloud), m) if out.cloud else None
        if not host_outputs:
            return None, None, None, cloud
        fr = PreprocessedFrame(stamp, stamp + out.last_time, bufs["times"][:m].copy(), bufs["intensities"][:m].copy() if it is not None else None, bufs["xyzw"][:m].copy(), k, bufs["neighbors"][:m].reshape(-1).copy())
        covs = bufs["cov4x4"][:m].reshape(m, 4, 4).transpose(0, 2, 1).copy()
        return fr, bufs["normals4"][:m].copy(), covs, cloud


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
