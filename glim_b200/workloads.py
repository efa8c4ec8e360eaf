"""The five BASELINE.json configurations as factor-set workloads, built the way GLIM's modules build them.

Each builder reproduces the factor-creation rule of the module it names (file:line below) on synthetic scans
(glim_b200.synth), so the bench and the parity tests sweep exactly the factor sets the reference would hand to
NonlinearFactorSetGPU::linearize:

    single_pair        OdometryEstimationCPU::create_factors      src/glim/odometry/odometry_estimation_cpu.cpp:105-110
    odometry_gpu       OdometryEstimationGPU::create_frame / create_factors / update_keyframes_overlap
                                                                   src/glim/odometry/odometry_estimation_gpu.cpp:86-107, :128-206, :212-295
    sub_mapping_gpu    SubMapping::insert_keyframe / insert_frame  src/glim/mapping/sub_mapping.cpp:275-316, :391-401
    global_mapping_gpu GlobalMapping::insert_submap / create_matching_cost_factors
                                                                   src/glim/mapping/global_mapping.cpp:234-283, :430-484
    livox_stress       odometry_gpu's factor rule on one 500 k-point MID-360-shaped frame

This module (like glim_b200.synth) builds bench / test INPUTS; it is not part of the drop-in surface.  Host-side logic only;
every GPU operation goes through glim_b200.gpu (the C-ABI).  When no context is given (CPU-only construction of the sample the
reference arm times, CPU tests) the overlap GATE is evaluated by a host numpy twin of its definition -- that affects only which
pairs the synthetic workload contains, never a product result.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

from . import gpu, synth


@dataclasses.dataclass
class Factor:
    target: int  # index of the cloud whose voxel map is the target
    level: int  # voxel-map level of the target
    source: int  # index of the source cloud
    pair: int  # (target, source) pair id: levels of one pair share it


@dataclasses.dataclass
class FactorSet:
    """One NonlinearFactorSetGPU::linearize call: factors + the relative poses to linearize at."""
    factors: list
    deltas: np.ndarray  # (F,4,4) T_target^-1 T_source


class Workload:
    def __init__(self, name: str, ctx: gpu.Context | None):
        self.name = name
        self.ctx = ctx
        self.host_clouds: list = []  # (points (N,4), covs (N,4,4))
        self.host_normals: list = []  # (N,4) per cloud (only the workloads whose factors use surface validation fill it)
        self.surface_validation = False  # factor->set_enable_surface_validation(true): odometry only (odometry_estimation_gpu.cpp:145, :162)
        self.poses: list = []  # ground-truth T_world_sensor per cloud
        self.resolutions: list = []  # voxel resolution per level
        self.clouds: list = []  # gpu.PointCloudGPU
        self.maps: list = []  # [cloud][level] gpu.GaussianVoxelMapGPU
        self.sets: list = []  # FactorSet
        self.notes: dict = {}
        self._cache: dict = {}

    # ---- device side -------------------------------------------------------------------------
    def upload(self):
        if self.ctx is None:  # host-only construction (CPU tests, reference arm without a GPU)
            return
        for k, (pts, cov) in enumerate(self.host_clouds):
            nrm = self.host_normals[k] if k < len(self.host_normals) else None
            self.clouds.append(gpu.PointCloudGPU.clone(pts, cov, nrm, ctx=self.ctx))

    def overlap(self, targets, level, source, deltas) -> float:
        """overlap_gpu(voxelmaps, source, deltas): GPU kernel when a context exists, else a host numpy twin of the
        same definition (fraction of transformed source points whose voxel is occupied in any target)."""
        if self.ctx is not None:
            return gpu.overlap_gpu([self.maps[t][level] for t in targets], self.clouds[source], deltas, ctx=self.ctx)
        res = self.resolutions[level]
        inv = np.float32(1.0) / np.float32(res)
        src = self.host_clouds[source][0][:, :3]
        hit = np.zeros(len(src), bool)
        for t, d in zip(targets, deltas):
            key = ("keys", t, level)
            if key not in self._cache:
                c = np.floor(self.host_clouds[t][0][:, :3].astype(np.float32) * inv).astype(np.int64) + (1 << 20)
                self._cache[key] = np.unique((c[:, 0] << 42) | (c[:, 1] << 21) | c[:, 2])
            q = (src @ d[:3, :3].T + d[:3, 3]).astype(np.float32)
            c = np.floor(q * inv).astype(np.int64) + (1 << 20)
            hit |= np.isin((c[:, 0] << 42) | (c[:, 1] << 21) | c[:, 2], self._cache[key])
        return float(hit.mean()) if len(src) else 0.0

    def build_maps(self, which=None):
        """GaussianVoxelMapGPU(resolution, 8192*2, 10, 1e-3).insert(frame) per level (odometry_estimation_gpu.cpp:97-106)."""
        if self.ctx is None:
            return
        self.maps = [None] * len(self.clouds)
        for i, c in enumerate(self.clouds):
            if which is not None and i not in which:
                continue
            self.maps[i] = [gpu.GaussianVoxelMapGPU(r, 8192 * 2, 10, 1e-3, ctx=self.ctx).insert(c) for r in self.resolutions]

    def gpu_factors(self, fset: FactorSet):
        out = []
        for f in fset.factors:
            g = gpu.IntegratedVGICPFactorGPU(f.target, f.source, self.maps[f.target][f.level], self.clouds[f.source], ctx=self.ctx)
            if self.surface_validation:
                g.set_enable_surface_validation(True)
            out.append(g)
        return out

    def gt_delta(self, target: int, source: int):
        return synth.inv_pose(self.poses[target]) @ self.poses[source]

    def noisy_deltas(self, factors, rng, sigma_rot, sigma_trans):
        """GT relative pose of each PAIR composed with one noise draw per pair (levels of a pair share the pose)."""
        cache = {}
        out = []
        for f in factors:
            if f.pair not in cache:
                cache[f.pair] = synth.perturb(self.gt_delta(f.target, f.source), rng, sigma_rot, sigma_trans)
            out.append(cache[f.pair])
        return np.stack(out) if out else np.zeros((0, 4, 4))

    @property
    def point_factors(self):
        return sum(len(self.host_clouds[f.source][0]) for s in self.sets for f in s.factors)


# ------------------------------------------------------------------------------------------------------------------
# scan generation (numpy by default; torch on the GPU when asked -- synthetic input plumbing, not the product)
# ------------------------------------------------------------------------------------------------------------------
def _covariances(points, k, ctx, use_gpu):
    if use_gpu and ctx is not None:
        from . import preprocess

        nb = preprocess.find_neighbors(points, k, ctx=ctx)
        return preprocess.CloudCovarianceEstimation(ctx=ctx).estimate(points, nb)
    return synth.with_covariances(points, k)


def make_scan(scene, sensor, T, rng, n_rays=None, max_points=None, ctx=None, use_gpu=False, with_normals=False):
    pts, tms = synth.scan(scene, sensor, T, rng, n_rays=n_rays, backend="torch" if (use_gpu and ctx is not None) else "numpy")
    if max_points is not None and len(pts) > max_points:
        sel = np.sort(rng.choice(len(pts), max_points, replace=False))  # random_sampling keeps the order (sub_mapping.cpp:385)
        pts, tms = pts[sel], tms[sel]
    nrm, cov = _covariances(pts, 10, ctx, use_gpu)
    if with_normals:
        return np.ascontiguousarray(pts), np.ascontiguousarray(cov), np.ascontiguousarray(nrm)
    return np.ascontiguousarray(pts), np.ascontiguousarray(cov)


# ------------------------------------------------------------------------------------------------------------------
# M1  single pair (CPU odometry config with registration_type = VGICP)
# ------------------------------------------------------------------------------------------------------------------
def single_pair(ctx, n_rays=None, sensor="generic64", num_draws=16, use_gpu=False) -> Workload:
    w = Workload("single_pair", ctx)
    sc = synth.make_hall_scene()
    traj = synth.arc_trajectory(8)
    for i in (3, 4):
        w.host_clouds.append(make_scan(sc, sensor, traj[i], synth.rng_for(101, i), n_rays=n_rays, ctx=ctx, use_gpu=use_gpu))
        w.poses.append(traj[i])
    w.resolutions = [0.5]  # vgicp_resolution 0.5, vgicp_voxelmap_levels 1 (config_odometry_cpu.json:30-31)
    w.upload()
    w.build_maps(which={0})
    rng = synth.rng_for(102)
    # unary form, fixed target pose (odometry_estimation_cpu.cpp:107: IntegratedVGICPFactor(gtsam::Pose3(), X(current), ...))
    for _ in range(num_draws):
        f = [Factor(0, 0, 1, 0)]
        w.sets.append(FactorSet(f, w.noisy_deltas(f, rng, 0.01, 0.05)))
    return w


# ------------------------------------------------------------------------------------------------------------------
# M2 / M5  GPU odometry stream
# ------------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class OdometryParams:
    """config/config_odometry_gpu.json:54-68 (voxel_resolution_max == voxel_resolution: BASELINE fixes '0.25 m voxels')."""
    voxel_resolution: float = 0.25
    voxelmap_levels: int = 2
    voxelmap_scaling_factor: float = 2.0
    full_connection_window_size: int = 2
    max_num_keyframes: int = 15
    keyframe_min_overlap: float = 0.01
    keyframe_max_overlap: float = 0.7
    smoother_lag_frames: int = 50  # smoother_lag 5.0 s at 10 Hz


def odometry_stream(ctx, n_frames=64, first_bench_frame=16, sensor="hdl32", n_rays=None, params: OdometryParams | None = None, name="odometry_gpu", use_gpu=False, step=1.0) -> Workload:
    """Runs the reference's keyframe bookkeeping over the whole stream (needs the GPU for overlap_gpu) and records,
    for every frame >= first_bench_frame, the factor set OdometryEstimationGPU::create_factors would create."""
    p = params or OdometryParams()
    w = Workload(name, ctx)
    sc = synth.make_hall_scene()
    traj = synth.arc_trajectory(n_frames, step=step)
    rng_pose = synth.rng_for(202)
    w.surface_validation = True  # odometry_estimation_gpu.cpp:145, :162
    w.notes["surface_validation"] = True
    for i in range(n_frames):
        pts, cov, nrm = make_scan(sc, sensor, traj[i], synth.rng_for(201, i), n_rays=n_rays, ctx=ctx, use_gpu=use_gpu, with_normals=True)
        w.host_clouds.append((pts, cov))
        w.host_normals.append(nrm)
        w.poses.append(traj[i])
    w.resolutions = [p.voxel_resolution * p.voxelmap_scaling_factor**l for l in range(p.voxelmap_levels)]
    w.upload()
    w.build_maps()
    keyframes: list[int] = []
    for cur in range(n_frames):
        # ---- create_factors (odometry_estimation_gpu.cpp:128-206) ----
        factors = []
        if cur > 0:
            pair = 0
            for tgt in range(cur - p.full_connection_window_size, cur):  # :175-181
                if tgt < 0:
                    continue
                for l in range(p.voxelmap_levels):
                    factors.append(Factor(tgt, l, cur, pair))
                pair += 1
            for kf in keyframes:  # :183-203 (binary inside the smoother window, unary outside: same device work)
                if kf >= cur - p.full_connection_window_size:
                    continue
                for l in range(p.voxelmap_levels):
                    factors.append(Factor(kf, l, cur, pair))
                pair += 1
        if cur >= first_bench_frame and factors:
            w.sets.append(FactorSet(factors, w.noisy_deltas(factors, rng_pose, 0.002, 0.01)))
        # ---- update_keyframes_overlap (:212-295) ----
        _update_keyframes_overlap(w, keyframes, cur, p)
    w.notes["keyframes_final"] = list(keyframes)
    return w


def _update_keyframes_overlap(w: Workload, keyframes: list, cur: int, p: OdometryParams):
    if not keyframes:  # :219-222
        keyframes.append(cur)
        return
    last = len(w.resolutions) - 1  # voxelmaps.back()
    deltas = [w.gt_delta(k, cur) for k in keyframes]
    overlap = w.overlap(keyframes, last, cur, deltas)  # :231
    if overlap > p.keyframe_max_overlap:
        return
    keyframes.append(cur)  # :236-237
    if len(keyframes) <= p.max_num_keyframes:
        return
    # remove keyframes without overlap to the new keyframe (:245-254)
    i = 0
    while i < len(keyframes):
        k = keyframes[i]
        ov = w.overlap([k], last, cur, [w.gt_delta(k, cur)])
        if ov < p.keyframe_min_overlap:
            keyframes.pop(i)
        else:
            i += 1
    if len(keyframes) <= p.max_num_keyframes:
        return
    # remove the keyframe with the minimum score (:261-294)
    scores = []
    for i in range(len(keyframes) - 1):
        k = keyframes[i]
        ov_latest = w.overlap([k], last, cur, [w.gt_delta(k, cur)])
        others = [o for j, o in enumerate(keyframes[:-1]) if j != i]
        ov_others = w.overlap(others, last, k, [w.gt_delta(o, k) for o in others])
        scores.append(ov_latest * (1.0 - ov_others))
    keyframes.pop(int(np.argmin(scores)))


def livox_stress(ctx, n_rays=500_000, use_gpu=False, n_targets=17) -> Workload:
    """M5: one dense MID-360-shaped frame against (2 window frames + 15 keyframes) x 2 levels at 0.1 / 0.2 m."""
    w = Workload("livox_stress", ctx)
    sc = synth.make_hall_scene()
    n_tgt = n_targets
    traj = synth.arc_trajectory(n_tgt + 1, step=0.5)
    w.surface_validation = True  # odometry's factor rule (odometry_estimation_gpu.cpp:145, :162)
    w.notes["surface_validation"] = True
    for i in range(n_tgt + 1):
        pts, cov, nrm = make_scan(sc, "mid360", traj[i], synth.rng_for(501, i), n_rays=n_rays, ctx=ctx, use_gpu=use_gpu, with_normals=True)
        w.host_clouds.append((pts, cov))
        w.host_normals.append(nrm)
        w.poses.append(traj[i])
    w.resolutions = [0.1, 0.2]
    w.upload()
    w.build_maps(which=set(range(n_tgt)))
    factors = [Factor(t, l, n_tgt, t) for t in range(n_tgt) for l in range(2)]
    w.sets.append(FactorSet(factors, w.noisy_deltas(factors, synth.rng_for(502), 0.002, 0.01)))
    return w


# ------------------------------------------------------------------------------------------------------------------
# M3  sub-mapping bundle
# ------------------------------------------------------------------------------------------------------------------
def sub_mapping_bundle(ctx, n_keyframes=15, sensor="os1_64", n_rays=None, num_iterations=1, use_gpu=False) -> Workload:
    """15 keyframes, fully connected: for every earlier keyframe i and level, a factor (X(i), X(current)) with target =
    voxelmap_i,level and source = the new keyframe's cloud (sub_mapping.cpp:276-310): 105 pairs x 2 levels."""
    w = Workload("sub_mapping_gpu", ctx)
    sc = synth.make_hall_scene()
    traj = synth.arc_trajectory(n_keyframes, step=1.5)
    for i in range(n_keyframes):
        w.host_clouds.append(make_scan(sc, sensor, traj[i], synth.rng_for(301, i), n_rays=n_rays, ctx=ctx, use_gpu=use_gpu))
        w.poses.append(traj[i])
    w.resolutions = [0.25, 0.5]  # keyframe_voxel_resolution 0.25, keyframe_voxelmap_levels 2 (config_sub_mapping_gpu.json)
    w.upload()
    w.build_maps()
    factors, pair = [], 0
    for cur in range(1, n_keyframes):
        for i in range(cur):
            for l in range(len(w.resolutions)):
                factors.append(Factor(i, l, cur, pair))
            pair += 1
    rng = synth.rng_for(302)
    for it in range(num_iterations):  # LM-like iterations with shrinking noise
        s = 0.5**it
        w.sets.append(FactorSet(factors, w.noisy_deltas(factors, rng, 0.01 * s, 0.05 * s)))
    return w


# ------------------------------------------------------------------------------------------------------------------
# M4  global mapping
# ------------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class GlobalMappingParams:
    """config/config_global_mapping_gpu.json (submap_voxel_resolution_max == submap_voxel_resolution: BASELINE fixes 0.5 / 1.0 m)."""
    submap_voxel_resolution: float = 0.5
    submap_voxelmap_levels: int = 2
    submap_voxelmap_scaling_factor: float = 2.0
    max_implicit_loop_distance: float = 100.0
    min_implicit_loop_overlap: float = 0.2
    submap_target_num_points: int = 50000  # config_sub_mapping_gpu.json:54


def global_mapping(ctx, n_submaps=256, laps=4, sensor="os1_64", n_rays=None, params: GlobalMappingParams | None = None, use_gpu=False, side=180.0) -> Workload:
    """256 submaps on a street loop driven `laps` times; for every new submap, factors to ALL previous submaps within
    max_implicit_loop_distance whose overlap_auto(submaps[i].voxelmaps.back(), current.frame, delta) >= min_implicit_loop_overlap,
    one per voxel-map level (global_mapping.cpp:441-470).  The whole graph is one relinearization sweep."""
    p = params or GlobalMappingParams()
    w = Workload("global_mapping_gpu", ctx)
    sc = synth.make_blocks_scene()
    traj = synth.loop_trajectory(n_submaps // laps, laps, side=side)
    for i in range(n_submaps):
        w.host_clouds.append(make_scan(sc, sensor, traj[i], synth.rng_for(401, i), n_rays=n_rays, max_points=p.submap_target_num_points, ctx=ctx, use_gpu=use_gpu))
        w.poses.append(traj[i])
    w.resolutions = [p.submap_voxel_resolution * p.submap_voxelmap_scaling_factor**l for l in range(p.submap_voxelmap_levels)]
    w.upload()
    w.build_maps()
    factors, pair = [], 0
    pair_overlap = {}
    d2max = p.max_implicit_loop_distance**2
    for cur in range(1, n_submaps):
        for i in range(cur):
            if np.sum((w.poses[i][:3, 3] - w.poses[cur][:3, 3]) ** 2) > d2max:  # :442-445
                continue
            ov = w.overlap([i], len(w.resolutions) - 1, cur, [w.gt_delta(i, cur)])  # :447-448
            if ov < p.min_implicit_loop_overlap:
                continue
            for l in range(p.submap_voxelmap_levels):  # :462-467
                factors.append(Factor(i, l, cur, pair))
            pair_overlap[pair] = ov
            pair += 1
    w.notes["num_pairs"] = pair
    w.notes["_pair_overlap"] = pair_overlap  # used to balance the multi-GPU partition
    w.sets.append(FactorSet(factors, w.noisy_deltas(factors, synth.rng_for(402), 0.02, 0.2)))
    return w


BUILDERS = {
    "single_pair": single_pair,
    "odometry_gpu": odometry_stream,
    "sub_mapping_gpu": sub_mapping_bundle,
    "global_mapping_gpu": global_mapping,
    "livox_stress": livox_stress,
}
