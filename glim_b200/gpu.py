"""Host-side mirror (Python) of the gtsam_points GPU class surface GLIM constructs (SURVEY.md 8(b) 'inner'
boundary), forwarding to the C-ABI of libglim_b200.so.  Same names, argument meaning and defaults as
the reference call sites; used by the parity tests and the bench.  The C++ twin of this file is
include/glim_b200/gtsam_points_compat.hpp.

    PointCloudGPU.clone(points, covs)                       odometry_estimation_gpu.cpp:96
    GaussianVoxelMapGPU(resolution, 8192*2, 10, 1e-3).insert(cloud)   odometry_estimation_gpu.cpp:103-104
    IntegratedVGICPFactorGPU(target_key | fixed_target_pose, source_key, voxelmap, source)   :144, :161
    NonlinearFactorSetGPU.add(...) / .linearize(values)     odometry_estimation_gpu.cpp:383-386
    overlap_gpu(voxelmap(s), source, delta(s))              odometry_estimation_gpu.cpp:231, :248
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, f64, lib, pose16, ptr

LIN_DTYPE = np.dtype([("H_tt", "f8", (36,)), ("H_ss", "f8", (36,)), ("H_ts", "f8", (36,)), ("b_t", "f8", (6,)), ("b_s", "f8", (6,)), ("error", "f8"), ("num_inliers", "f8")])
assert LIN_DTYPE.itemsize == 122 * 8


class Context:
    """gtsam_points::CUDAStream + StreamTempBufferRoundRobin (odometry_estimation_gpu.cpp:76-77)."""

    def __init__(self, device: int = 0, cuda_stream: int | None = None):
        h = C.c_void_p()
        if cuda_stream is None:
            check(lib().gb_ctx_create(device, C.byref(h)))
        else:
            check(lib().gb_ctx_create_on_stream(device, C.c_void_p(cuda_stream), C.byref(h)))
        self.h = h
        self.device = device

    def synchronize(self):
        check(lib().gb_ctx_synchronize(self.h))

    @property
    def stream(self) -> int:
        return lib().gb_ctx_stream(self.h) or 0

    @property
    def kernel_launches(self) -> int:
        return lib().gb_ctx_kernel_launches(self.h)

    def close(self):
        if self.h:
            lib().gb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class PointCloudGPU:
    """gtsam_points::PointCloudGPU (device copy of points / covs / normals in fp32)."""

    def __init__(self, ctx: Context, handle, n: int):
        self.ctx, self.h, self.n = ctx, handle, n

    @staticmethod
    def clone(points, covs=None, normals=None, ctx: Context | None = None) -> "PointCloudGPU":
        """points (N,4) fp64 with w = 1; covs (N,4,4) fp64 [i,row,col] (symmetric); normals (N,4) or None."""
        ctx = ctx or default_context()
        points = f64(points)
        n = points.shape[0]
        c16 = None
        if covs is not None:
            covs = np.asarray(covs, dtype=np.float64)
            # [i,row,col] -> column-major 4x4 per point
            c16 = np.ascontiguousarray(np.swapaxes(covs.reshape(n, 4, 4), 1, 2)).reshape(n, 16)
        nr = f64(normals) if normals is not None else None
        h = C.c_void_p()
        check(lib().gb_cloud_upload(ctx.h, n, ptr(points), ptr(c16), ptr(nr), C.byref(h)))
        return PointCloudGPU(ctx, h, n)

    def size(self) -> int:
        return self.n

    def download(self):
        xyz = np.empty((self.n, 3), np.float32)
        cov6 = np.empty((self.n, 6), np.float32)
        check(lib().gb_cloud_download(self.h, ptr(xyz), ptr(cov6)))
        return xyz, cov6

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().gb_cloud_destroy(self.h)
            self.h = None


class GaussianVoxelMapGPU:
    """gtsam_points::GaussianVoxelMapGPU(resolution, init_num_buckets, max_bucket_scan_count, target_points_drop_rate)."""

    def __init__(self, resolution: float, init_num_buckets: int = 8192 * 2, max_bucket_scan_count: int = 10, target_points_drop_rate: float = 1e-3, ctx: Context | None = None):
        self.resolution = float(resolution)
        self.init_num_buckets = init_num_buckets
        self.max_bucket_scan_count = max_bucket_scan_count
        self.target_points_drop_rate = target_points_drop_rate
        self.ctx = ctx
        self.h = None
        self.num_voxels = 0
        self.num_buckets = 0
        self._cloud = None

    def insert(self, cloud: PointCloudGPU):
        if self.h is not None:
            raise capi.GlimB200Error("GaussianVoxelMapGPU::insert may be called once (as every GLIM call site does)")
        self.ctx = self.ctx or cloud.ctx
        h = C.c_void_p()
        check(lib().gb_voxelmap_build(self.ctx.h, cloud.h, self.resolution, self.init_num_buckets, self.max_bucket_scan_count, self.target_points_drop_rate, C.byref(h)))
        self.h = h
        nv, nb, res = C.c_int(), C.c_int(), C.c_float()
        check(lib().gb_voxelmap_info(h, C.byref(nv), C.byref(nb), C.byref(res)))
        self.num_voxels, self.num_buckets = nv.value, nb.value
        return self

    def voxel_resolution(self) -> float:
        return self.resolution

    def download(self):
        buckets = np.empty((self.num_buckets, 4), np.int32)
        vnum = np.empty((self.num_voxels,), np.int32)
        vmean = np.empty((self.num_voxels, 3), np.float32)
        vcov = np.empty((self.num_voxels, 6), np.float32)
        check(lib().gb_voxelmap_download(self.h, ptr(buckets), ptr(vnum), ptr(vmean), ptr(vcov)))
        return buckets, vnum, vmean, vcov

    def __del__(self):
        if getattr(self, "h", None) and self.ctx and self.ctx.h:
            lib().gb_voxelmap_destroy(self.h)
            self.h = None


def unpack_linearized(rec) -> dict:
    """gb_linearized6 record -> numpy blocks indexed [row, col]."""
    return {
        "H_tt": np.asarray(rec["H_tt"]).reshape(6, 6).T.copy(),
        "H_ss": np.asarray(rec["H_ss"]).reshape(6, 6).T.copy(),
        "H_ts": np.asarray(rec["H_ts"]).reshape(6, 6).T.copy(),
        "b_t": np.asarray(rec["b_t"]).copy(),
        "b_s": np.asarray(rec["b_s"]).copy(),
        "error": float(rec["error"]),
        "num_inliers": float(rec["num_inliers"]),
    }


class IntegratedVGICPFactorGPU:
    """gtsam_points::IntegratedVGICPFactorGPU.

    Binary form: (target_key, source_key, voxelmap, source); unary form: (fixed_target_pose 4x4, source_key, ...).
    `values` is a dict key -> 4x4 pose.  delta = T_target^-1 T_source.
    """

    def __init__(self, target, source_key, voxelmap: GaussianVoxelMapGPU, source: PointCloudGPU, ctx: Context | None = None):
        self.ctx = ctx or source.ctx
        if isinstance(target, np.ndarray):
            self.fixed_target_pose = np.asarray(target, dtype=np.float64).reshape(4, 4)
            self.target_key = None
        else:
            self.fixed_target_pose = None
            self.target_key = target
        self.source_key = source_key
        self.voxelmap, self.source = voxelmap, source  # keep alive, as the reference factor's shared_ptrs do
        self.flags = 0
        self.h = None
        self._lin_point = None

    def set_enable_surface_validation(self, enable: bool):
        if self.h is not None:
            raise capi.GlimB200Error("set_enable_surface_validation must precede the first linearize")
        self.flags = capi.GB_FACTOR_SURFACE_VALIDATION if enable else 0

    def _handle(self):
        if self.h is None:
            h = C.c_void_p()
            check(lib().gb_vgicp_factor_create(self.ctx.h, self.voxelmap.h, self.source.h, self.flags, C.byref(h)))
            self.h = h
        return self.h

    def keys(self):
        return [self.source_key] if self.target_key is None else [self.target_key, self.source_key]

    def dim(self):
        return 6

    def is_binary(self):
        return self.target_key is not None

    def get_fixed_target_pose(self):
        return self.fixed_target_pose

    def delta(self, values) -> np.ndarray:
        Tt = self.fixed_target_pose if self.target_key is None else np.asarray(values[self.target_key], dtype=np.float64)
        Ts = np.asarray(values[self.source_key], dtype=np.float64)
        Tti = np.eye(4)
        Tti[:3, :3] = Tt[:3, :3].T
        Tti[:3, 3] = -Tt[:3, :3].T @ Tt[:3, 3]
        return Tti @ Ts

    def linearize(self, values) -> dict:
        d = self.delta(values)
        self._lin_point = d
        out = np.zeros(1, LIN_DTYPE)
        check(lib().gb_vgicp_linearize(self._handle(), ptr(pose16(d)), ptr(out)))
        return unpack_linearized(out[0])

    def error(self, values) -> float:
        """error at `values` with the inlier set of the last linearization point."""
        d = self.delta(values)
        lin = self._lin_point if self._lin_point is not None else d
        e = C.c_double()
        check(lib().gb_vgicp_error(self._handle(), ptr(pose16(lin)), ptr(pose16(d)), C.byref(e)))
        return e.value

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().gb_vgicp_factor_destroy(self.h)
            self.h = None


class NonlinearFactorSetGPU:
    """gtsam_points::NonlinearFactorSetGPU: batch linearization of every GPU factor of a graph."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx
        self.factors: list[IntegratedVGICPFactorGPU] = []

    def add(self, factors):
        for f in (factors if isinstance(factors, (list, tuple)) else [factors]):
            if isinstance(f, IntegratedVGICPFactorGPU):
                self.factors.append(f)
                self.ctx = self.ctx or f.ctx
        return self

    def size(self):
        return len(self.factors)

    def _handles(self):
        return (C.c_void_p * len(self.factors))(*[f._handle() for f in self.factors])

    def linearize(self, values) -> list[dict]:
        F = len(self.factors)
        if F == 0:
            return []
        deltas = np.stack([f.delta(values) for f in self.factors])
        for f, d in zip(self.factors, deltas):
            f._lin_point = d
        return [unpack_linearized(r) for r in self.linearize_deltas(deltas)]

    def linearize_deltas(self, deltas) -> np.ndarray:
        F = len(self.factors)
        out = np.zeros(F, LIN_DTYPE)
        if F:
            check(lib().gb_factor_set_linearize(self.ctx.h, F, C.cast(self._handles(), C.c_void_p), ptr(pose16(deltas)), ptr(out)))
        return out

    def error_deltas(self, deltas_lin, deltas_eval) -> np.ndarray:
        F = len(self.factors)
        out = np.zeros(F)
        if F:
            check(lib().gb_factor_set_error(self.ctx.h, F, C.cast(self._handles(), C.c_void_p), ptr(pose16(deltas_lin)), ptr(pose16(deltas_eval)), ptr(out)))
        return out


class Sweep:
    """A prepared, device-resident batch (gb_sweep): upload poses / launch / fetch are separate steps."""

    def __init__(self, ctx: Context, factors: list[IntegratedVGICPFactorGPU], pair_index=None):
        self.ctx = ctx
        self.factors = list(factors)
        F = len(factors)
        arr = (C.c_void_p * F)(*[f._handle() for f in factors])
        pi = np.ascontiguousarray(pair_index, dtype=np.int32) if pair_index is not None else None
        h = C.c_void_p()
        check(lib().gb_sweep_create(ctx.h, F, C.cast(arr, C.c_void_p), ptr(pi), C.byref(h)))
        self.h = h
        self.F = F
        pf, ab, nt, gs = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_uint32()
        check(lib().gb_sweep_stats(h, C.byref(pf), C.byref(ab), C.byref(nt), C.byref(gs)))
        self.point_factors, self.algorithmic_bytes, self.num_tiles, self.grid = pf.value, ab.value, nt.value, gs.value

    def attach_peer_slab(self, peer_slab: "PeerSlab | None"):
        self._peer = peer_slab  # keep alive
        check(lib().gb_sweep_attach_peer_slab(self.h, peer_slab.h if peer_slab is not None else None))

    def attach_slab(self, device_ptr: int, num_pairs: int):
        check(lib().gb_sweep_attach_slab(self.h, C.c_void_p(device_ptr), num_pairs))

    def set_poses(self, deltas):
        self._poses = pose16(deltas)  # keep the host array alive until the copy was staged
        check(lib().gb_sweep_set_poses(self.h, ptr(self._poses)))

    def launch(self):
        check(lib().gb_sweep_launch(self.h))

    def fetch(self) -> np.ndarray:
        out = np.zeros(self.F, LIN_DTYPE)
        check(lib().gb_sweep_fetch(self.h, ptr(out)))
        return out

    def linearize(self, deltas) -> np.ndarray:
        """gb_sweep_linearize: poses up, one launch (a CUDA graph for small sweeps), records down."""
        out = np.zeros(self.F, LIN_DTYPE)
        self._poses = pose16(deltas)
        check(lib().gb_sweep_linearize(self.h, ptr(self._poses), ptr(out)))
        return out

    def linearize_raw(self, poses16: np.ndarray, out: np.ndarray) -> np.ndarray:
        """The same call with caller-owned buffers (poses16: (F,16) float64 column-major, out: (F,) LIN_DTYPE): no Python-side
        array work around the C entry point."""
        check(lib().gb_sweep_linearize(self.h, poses16.ctypes.data, out.ctypes.data))
        return out

    def results_device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().gb_sweep_results_device(self.h, C.byref(p)))
        return p.value or 0

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().gb_sweep_destroy(self.h)
            self.h = None


class PeerSlab:
    """gb_peer_slab: ping-pong fp32 [num_pairs][96] result buffers shared with the other ranks of the box through CUDA
    IPC; the sweep's epilogue stores finished pair rows straight into every rank's buffer (multi-GPU exchange fused into
    the kernel).  `exchange(handle_bytes) -> list[bytes]` must all-gather the 64-byte handles in rank order
    (torch.distributed in the bench); with world == 1 nothing is exchanged."""

    def __init__(self, ctx: Context, num_pairs: int, world: int = 1, rank: int = 0, exchange=None):
        self.ctx, self.num_pairs, self.world, self.rank = ctx, num_pairs, world, rank
        h = C.c_void_p()
        check(lib().gb_peer_slab_create(ctx.h, num_pairs, world, rank, C.byref(h)))
        self.h = h
        if world > 1:
            buf = (C.c_ubyte * capi.GB_IPC_HANDLE_BYTES)()
            check(lib().gb_peer_slab_export(self.h, C.cast(buf, C.c_void_p)))
            handles = exchange(bytes(buf))
            assert len(handles) == world and all(len(x) == capi.GB_IPC_HANDLE_BYTES for x in handles)
            blob = (C.c_ubyte * (capi.GB_IPC_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
            check(lib().gb_peer_slab_connect(self.h, C.cast(blob, C.c_void_p)))

    def signal_wait(self):
        check(lib().gb_peer_slab_signal_wait(self.h))

    def fetch(self) -> np.ndarray:
        out = np.empty((self.num_pairs, capi.GB_SLAB_STRIDE), np.float32)
        check(lib().gb_peer_slab_fetch(self.h, ptr(out)))
        return out

    def fetch_async(self) -> np.ndarray:
        """Enqueue the D2H into the slab's pinned buffer; the returned view is valid after the next stream sync."""
        p = C.POINTER(C.c_float)()
        check(lib().gb_peer_slab_fetch_async(self.h, C.byref(p)))
        return np.ctypeslib.as_array(p, shape=(self.num_pairs, capi.GB_SLAB_STRIDE))

    def device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().gb_peer_slab_device_ptr(self.h, C.byref(p)))
        return p.value or 0

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().gb_peer_slab_destroy(self.h)
            self.h = None


def overlap_gpu(voxelmaps, source: PointCloudGPU, deltas, ctx: Context | None = None) -> float:
    """gtsam_points::overlap_gpu: single (voxelmap, delta) or lists (odometry_estimation_gpu.cpp:231, :248)."""
    if isinstance(voxelmaps, GaussianVoxelMapGPU):
        voxelmaps, deltas = [voxelmaps], [deltas]
    ctx = ctx or source.ctx
    T = len(voxelmaps)
    arr = (C.c_void_p * T)(*[m.h for m in voxelmaps])
    d = pose16(np.stack([np.asarray(x, dtype=np.float64) for x in deltas])) if T else np.zeros((0, 16))
    out = C.c_double()
    check(lib().gb_overlap(ctx.h, T, C.cast(arr, C.c_void_p), source.h, ptr(d), C.byref(out)))
    return out.value


def merge_frames_gpu(poses, frames, downsample_resolution: float, target_num_points: int = 0, seed: int = 0, ctx: Context | None = None, host_outputs: bool = True):
    """gtsam_points::merge_frames(poses, frames, downsample_resolution, target_num_points) on the device (sub_mapping.cpp:481-497).
    poses: K x (4,4) T_origin_frame; frames: K PointCloudGPU.  -> (points (M,4), covs (M,4,4) [i,row,col], PointCloudGPU)"""
    ctx = ctx or frames[0].ctx
    K = len(frames)
    arr = (C.c_void_p * K)(*[f.h for f in frames])
    T = pose16(np.stack([np.asarray(p, dtype=np.float64) for p in poses]))
    cap = sum(f.n for f in frames)
    pts = np.empty((cap, 4)) if host_outputs else None
    cov = np.empty((cap, 16)) if host_outputs else None
    m = C.c_size_t()
    h = C.c_void_p()
    check(lib().gb_merge_frames(ctx.h, K, C.cast(arr, C.c_void_p), ptr(T), float(downsample_resolution), int(target_num_points), int(seed), ptr(pts), ptr(cov), C.byref(m), C.byref(h)))
    cloud = PointCloudGPU(ctx, h, m.value) if h.value else None
    if not host_outputs:
        return None, None, cloud
    return pts[: m.value].copy(), cov[: m.value].reshape(-1, 4, 4).transpose(0, 2, 1).copy(), cloud


overlap_auto = overlap_gpu  # gtsam_points::overlap_auto dispatches to the GPU version for GPU voxel maps (sub_mapping.cpp:252)


def median_distance(points, max_scan_count: int = 256) -> float:
    """gtsam_points::median_distance(frame, 256) (odometry_estimation_gpu.cpp:91): strided sample of <= max_scan_count
    points, median of their norms.  256 points: stays on the host."""
    points = np.asarray(points)
    n = points.shape[0]
    if n == 0:
        return 0.0
    step = max(1, n // max_scan_count)
    d = np.linalg.norm(points[::step, :3], axis=1)
    d = np.sort(d)
    return float(d[len(d) // 2])
