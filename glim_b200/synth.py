"""Deterministic synthetic LiDAR-shaped inputs for the VGICP hot path.

There is no network and the reference vendors no data (docs/quickstart.md:6-10 only links
datasets), so every test and bench input is generated here: analytic ray casting of a seeded
scene with sensor patterns shaped like the ones BASELINE.json names (Velodyne HDL-32e, Ouster
OS1-64, a generic 64-beam 100k-ray pattern, Livox MID-360), range noise N(0, (2 cm)^2), the
range gate of config/config_preprocess.json:20-21 (0.5 m .. 100 m) and points ordered by time as
CloudPreprocessor does (src/glim/preprocess/cloud_preprocessor.cpp:135-136).

Host-side numpy only; no oracle and no CUDA dependency.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

SEED = 20240710


def rng_for(*key) -> np.random.Generator:
    """Philox stream keyed by (SEED, *key)."""
    k = [SEED] + [int(x) & 0xFFFFFFFF for x in key]
    while len(k) < 4:
        k.append(0)
    return np.random.Generator(np.random.Philox(key=np.array(k[:2], dtype=np.uint64), counter=np.array(k[2:4] + [0, 0], dtype=np.uint64)))


# ----------------------------------------------------------------------------------------------
# SE(3) helpers (GTSAM Pose3 conventions: tangent = [rotation(3); translation(3)])
# ----------------------------------------------------------------------------------------------
def hat(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def so3_exp(w):
    th = float(np.linalg.norm(w))
    K = hat(w)
    if th < 1e-10:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + (math.sin(th) / th) * K + ((1.0 - math.cos(th)) / (th * th)) * K @ K


def se3_exp(xi):
    """Pose3::Expmap([w; v]) as a 4x4 matrix."""
    w, v = np.asarray(xi[:3], float), np.asarray(xi[3:], float)
    th = float(np.linalg.norm(w))
    K = hat(w)
    R = so3_exp(w)
    if th < 1e-10:
        V = np.eye(3) + 0.5 * K
    else:
        V = np.eye(3) + ((1.0 - math.cos(th)) / th**2) * K + ((th - math.sin(th)) / th**3) * K @ K
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ v
    return T


def pose(x, y, z, yaw, pitch=0.0, roll=0.0):
    cy, sy, cp, sp, cr, sr = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch), math.cos(roll), math.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = [x, y, z]
    return T


def inv_pose(T):
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return Ti


def perturb(T, rng, sigma_rot, sigma_trans):
    """T * Exp(xi), xi ~ N(0, diag(sigma_rot^2 x3, sigma_trans^2 x3))."""
    xi = np.concatenate([rng.normal(0.0, sigma_rot, 3), rng.normal(0.0, sigma_trans, 3)])
    return T @ se3_exp(xi)


# ----------------------------------------------------------------------------------------------
# Scene: ground plane + axis-aligned boxes (seen from outside) + an optional enclosing room (seen
# from inside) + vertical cylinders.
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Scene:
    ground_z: float
    room: np.ndarray | None  # (2,3) min,max of the enclosing hall, or None (outdoor)
    boxes: np.ndarray  # (B,2,3)
    cylinders: np.ndarray  # (C,5): cx, cy, radius, zmin, zmax


def make_hall_scene(seed_key=0) -> Scene:
    """80 x 50 x 12 m hall, 24 boxes, 12 vertical cylinders (SURVEY 8(d) 'Scene')."""
    r = rng_for(1000, seed_key)
    gz = -1.8
    room = np.array([[-40.0, -25.0, gz], [40.0, 25.0, gz + 12.0]])
    def clear_of_path(cx, cy, half):
        # keep a 3 m corridor around the sensor path of arc_trajectory (radius 40 m about (0, -45)) free of obstacles
        return abs(math.hypot(cx, cy + 45.0) - 40.0) > half + 3.0

    boxes = []
    while len(boxes) < 24:
        c = np.array([r.uniform(-36, 36), r.uniform(-22, 22)])
        s = r.uniform(0.8, 4.0, 2)
        h = r.uniform(0.8, 6.0)
        if clear_of_path(c[0], c[1], float(np.hypot(s[0], s[1]))):
            boxes.append([[c[0] - s[0], c[1] - s[1], gz], [c[0] + s[0], c[1] + s[1], gz + h]])
    cyl = []
    while len(cyl) < 12:
        cx, cy, rad, top = r.uniform(-36, 36), r.uniform(-22, 22), r.uniform(0.2, 0.8), gz + r.uniform(3.0, 12.0)
        if clear_of_path(cx, cy, rad):
            cyl.append([cx, cy, rad, gz, top])
    return Scene(gz, room, np.array(boxes), np.array(cyl))


def make_blocks_scene(seed_key=0, extent=360.0, block=44.0, street=16.0) -> Scene:
    """Outdoor street grid: buildings on a regular lattice separated by streets (global-mapping scale)."""
    r = rng_for(2000, seed_key)
    gz = -1.8
    pitch = block + street
    n = int(extent // pitch)
    boxes = []
    for i in range(-n, n + 1):
        for j in range(-n, n + 1):
            cx, cy = i * pitch, j * pitch
            # split every block into 2x2 buildings of random height / setback for texture
            for sx in (-1, 1):
                for sy in (-1, 1):
                    hw = block / 4.0 - r.uniform(0.2, 1.5)
                    bx, by = cx + sx * block / 4.0, cy + sy * block / 4.0
                    boxes.append([[bx - hw, by - hw, gz], [bx + hw, by + hw, gz + r.uniform(6.0, 30.0)]])
    cyl = []
    for _ in range(200):
        # street furniture: poles / trunks near the street edges
        i, j = r.integers(-n, n + 1, 2)
        e = r.uniform(-0.5, 0.5) * pitch
        off = (block / 2.0 + r.uniform(1.0, 3.0)) * (1 if r.random() < 0.5 else -1)
        if r.random() < 0.5:
            cx, cy = i * pitch + off, j * pitch + e
        else:
            cx, cy = i * pitch + e, j * pitch + off
        cyl.append([cx, cy, r.uniform(0.1, 0.4), gz, gz + r.uniform(3.0, 9.0)])
    return Scene(gz, None, np.array(boxes), np.array(cyl))


def _raycast(scene: Scene, o: np.ndarray, d: np.ndarray, max_range: float):
    """Nearest hit distance for rays o + t d (d unit, (N,3) world frame). inf where nothing is hit."""
    n = d.shape[0]
    t_best = np.full(n, np.inf)
    eps = 1e-12
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / np.where(np.abs(d) < eps, eps, d)
        # ground
        tg = (scene.ground_z - o[2]) * inv[:, 2]
        tg = np.where(tg > 1e-6, tg, np.inf)
        t_best = np.minimum(t_best, tg)
        # enclosing room: exit distance of the slab intersection
        if scene.room is not None:
            t1 = (scene.room[0] - o) * inv
            t2 = (scene.room[1] - o) * inv
            texit = np.min(np.maximum(t1, t2), axis=1)
            t_best = np.minimum(t_best, np.where(texit > 1e-6, texit, np.inf))
        # boxes (cull those that cannot be reached)
        if len(scene.boxes):
            ctr = 0.5 * (scene.boxes[:, 0] + scene.boxes[:, 1])
            rad = 0.5 * np.linalg.norm(scene.boxes[:, 1] - scene.boxes[:, 0], axis=1)
            near = np.linalg.norm(ctr - o, axis=1) - rad < max_range
            for b in scene.boxes[near]:
                t1 = (b[0] - o) * inv
                t2 = (b[1] - o) * inv
                tn = np.max(np.minimum(t1, t2), axis=1)
                tf = np.min(np.maximum(t1, t2), axis=1)
                hit = (tn <= tf) & (tn > 1e-6)
                t_best = np.where(hit & (tn < t_best), tn, t_best)
        # vertical cylinders
        if len(scene.cylinders):
            cc = scene.cylinders
            near = np.hypot(cc[:, 0] - o[0], cc[:, 1] - o[1]) - cc[:, 2] < max_range
            for c in cc[near]:
                ox, oy = o[0] - c[0], o[1] - c[1]
                a = d[:, 0] ** 2 + d[:, 1] ** 2
                bq = 2.0 * (ox * d[:, 0] + oy * d[:, 1])
                cq = ox * ox + oy * oy - c[2] ** 2
                disc = bq * bq - 4.0 * a * cq
                ok = (disc > 0) & (a > eps)
                tt = (-bq - np.sqrt(np.where(ok, disc, 0.0))) / (2.0 * np.where(ok, a, 1.0))
                z = o[2] + tt * d[:, 2]
                hit = ok & (tt > 1e-6) & (z >= c[3]) & (z <= c[4])
                t_best = np.where(hit & (tt < t_best), tt, t_best)
    return t_best


# ----------------------------------------------------------------------------------------------
# Sensor patterns: unit directions in the sensor frame + per-ray time offset in [0, 0.1) s
# ----------------------------------------------------------------------------------------------
def _spinning(elev_deg: np.ndarray, n_az: int):
    az = np.arange(n_az) * (2.0 * math.pi / n_az)
    el = np.deg2rad(elev_deg)
    A, E = np.meshgrid(az, el, indexing="ij")  # azimuth-major == firing order
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    t = np.repeat(np.arange(n_az) * (0.1 / n_az), len(el))
    return d, t


def sensor_pattern(name: str, rng: np.random.Generator | None = None, n_rays: int | None = None):
    if name == "hdl32":  # 32 rings, -30.67..+10.67 deg, 1875 azimuth steps = 60 000 rays
        return _spinning(np.linspace(-30.67, 10.67, 32), n_az=(n_rays // 32) if n_rays else 1875)
    if name == "os1_64":  # 64 rings, +-16.6 deg, 2048 azimuth steps = 131 072 rays
        return _spinning(np.linspace(-16.6, 16.6, 64), n_az=(n_rays // 64) if n_rays else 2048)
    if name == "generic64":  # 64 rings, -25..+15 deg, 1563 azimuth steps ~ 100 k rays
        return _spinning(np.linspace(-25.0, 15.0, 64), n_az=(n_rays // 64) if n_rays else 1563)
    if name == "mid360":  # non-repetitive: uniform random directions, 360 x [-7, +52] deg
        n = n_rays or 500_000
        az = rng.uniform(0.0, 2.0 * math.pi, n)
        sz = rng.uniform(math.sin(math.radians(-7.0)), math.sin(math.radians(52.0)), n)
        cz = np.sqrt(1.0 - sz * sz)
        d = np.stack([cz * np.cos(az), cz * np.sin(az), sz], axis=-1)
        t = np.sort(rng.uniform(0.0, 0.1, n))
        return d, t
    raise ValueError(name)


def _raycast_torch(scene: Scene, o: np.ndarray, d: np.ndarray, max_range: float):
    """Same ray caster on the GPU with torch (fp64) -- only used to build synthetic bench inputs quickly."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device())
    D = torch.from_numpy(np.ascontiguousarray(d)).to(dev)
    O = torch.from_numpy(np.asarray(o, dtype=np.float64)).to(dev)
    eps = 1e-12
    inv = 1.0 / torch.where(D.abs() < eps, torch.full_like(D, eps), D)
    inf = torch.full((D.shape[0],), float("inf"), dtype=torch.float64, device=dev)
    tg = (scene.ground_z - O[2]) * inv[:, 2]
    t_best = torch.where(tg > 1e-6, tg, inf)
    if scene.room is not None:
        room = torch.from_numpy(scene.room).to(dev)
        t1 = (room[0] - O) * inv
        t2 = (room[1] - O) * inv
        texit = torch.maximum(t1, t2).min(dim=1).values
        t_best = torch.minimum(t_best, torch.where(texit > 1e-6, texit, inf))
    if len(scene.boxes):
        ctr = 0.5 * (scene.boxes[:, 0] + scene.boxes[:, 1])
        rad = 0.5 * np.linalg.norm(scene.boxes[:, 1] - scene.boxes[:, 0], axis=1)
        near = scene.boxes[np.linalg.norm(ctr - o, axis=1) - rad < max_range]
        for c0 in range(0, len(near), 32):
            B = torch.from_numpy(near[c0 : c0 + 32]).to(dev)  # (b,2,3)
            t1 = (B[None, :, 0, :] - O) * inv[:, None, :]
            t2 = (B[None, :, 1, :] - O) * inv[:, None, :]
            tn = torch.minimum(t1, t2).max(dim=2).values
            tf = torch.maximum(t1, t2).min(dim=2).values
            tn = torch.where((tn <= tf) & (tn > 1e-6), tn, torch.full_like(tn, float("inf")))
            t_best = torch.minimum(t_best, tn.min(dim=1).values)
    if len(scene.cylinders):
        cc = scene.cylinders
        near = cc[np.hypot(cc[:, 0] - o[0], cc[:, 1] - o[1]) - cc[:, 2] < max_range]
        if len(near):
            Cy = torch.from_numpy(near).to(dev)  # (c,5)
            ox = (O[0] - Cy[:, 0])[None, :]
            oy = (O[1] - Cy[:, 1])[None, :]
            a = (D[:, 0] ** 2 + D[:, 1] ** 2)[:, None]
            bq = 2.0 * (ox * D[:, 0:1] + oy * D[:, 1:2])
            cq = ox * ox + oy * oy - (Cy[:, 2] ** 2)[None, :]
            disc = bq * bq - 4.0 * a * cq
            ok = (disc > 0) & (a > eps)
            tt = (-bq - torch.sqrt(torch.where(ok, disc, torch.zeros_like(disc)))) / (2.0 * torch.where(ok, a.expand_as(disc), torch.ones_like(disc)))
            z = O[2] + tt * D[:, 2:3]
            hit = ok & (tt > 1e-6) & (z >= Cy[None, :, 3]) & (z <= Cy[None, :, 4])
            tt = torch.where(hit, tt, torch.full_like(tt, float("inf")))
            t_best = torch.minimum(t_best, tt.min(dim=1).values)
    return t_best.cpu().numpy()


def scan(scene: Scene, sensor: str, T_world_sensor: np.ndarray, rng: np.random.Generator, n_rays: int | None = None, min_range=0.5, max_range=100.0, noise=0.02, backend="numpy"):
    """One scan in the SENSOR frame: (points (N,4) float64 with w=1, times (N,) float64), time ordered."""
    d, t = sensor_pattern(sensor, rng, n_rays)
    R, o = T_world_sensor[:3, :3], T_world_sensor[:3, 3]
    dw = d @ R.T
    rngs = _raycast_torch(scene, o, dw, max_range) if backend == "torch" else _raycast(scene, o, dw, max_range)
    rngs = rngs + rng.normal(0.0, noise, rngs.shape)
    ok = np.isfinite(rngs) & (rngs > min_range) & (rngs < max_range)
    p = d[ok] * rngs[ok, None]
    pts = np.concatenate([p, np.ones((p.shape[0], 1))], axis=1)
    return np.ascontiguousarray(pts), np.ascontiguousarray(t[ok])


# ----------------------------------------------------------------------------------------------
# Covariances the way GLIM would attach them (k-NN with k = 10, PLANE regularization):
# an independent numpy/scipy implementation of CloudPreprocessor::find_neighbors
# (cloud_preprocessor.cpp:190-221) + CloudCovarianceEstimation::estimate
# (cloud_covariance_estimation.cpp:43-122), used to build inputs (and cross-check the oracle).
# ----------------------------------------------------------------------------------------------
def knn(points4: np.ndarray, k: int = 10) -> np.ndarray:
    from scipy.spatial import cKDTree

    tree = cKDTree(points4[:, :3])
    _, idx = tree.query(points4[:, :3], k=k, workers=-1)
    return np.ascontiguousarray(idx.reshape(points4.shape[0], k).astype(np.int32))


def plane_covariances(points4: np.ndarray, neighbors: np.ndarray):
    """-> normals (N,4), covs (N,4,4) [i,row,col]; cov = V diag(1e-3,1,1) V^T, normal = V[:,0] facing the sensor."""
    n, k = neighbors.shape
    P = points4[neighbors]  # (N,k,4)
    S = P.sum(axis=1)
    X = np.einsum("nki,nkj->nij", P, P)
    mean = S / k
    cov = (X - mean[:, :, None] * S[:, None, :]) / k
    w, V = np.linalg.eigh(cov[:, :3, :3])
    nrm = V[:, :, 0]
    covs = np.zeros((n, 4, 4))
    covs[:, :3, :3] = np.einsum("nik,k,njk->nij", V, np.array([1e-3, 1.0, 1.0]), V)
    flip = np.einsum("ni,ni->n", points4[:, :3], nrm) > 0.0
    nrm = np.where(flip[:, None], -nrm, nrm)
    normals = np.concatenate([nrm, np.zeros((n, 1))], axis=1)
    return normals, covs


def with_covariances(points4: np.ndarray, k: int = 10):
    nb = knn(points4, k)
    normals, covs = plane_covariances(points4, nb)
    return normals, covs


# ----------------------------------------------------------------------------------------------
# Trajectories
# ----------------------------------------------------------------------------------------------
def arc_trajectory(n_frames: int, step=1.0, radius=40.0, center=(0.0, -45.0), z=0.0):
    """Poses along a `radius` m arc, `step` m apart, heading tangent (M2: 10 Hz, 1 m / frame)."""
    poses = []
    span = (n_frames - 1) * step / radius
    for i in range(n_frames):
        phi = -0.5 * span + i * step / radius  # angle from the top of the circle
        x = center[0] + radius * math.sin(phi)
        y = center[1] + radius * math.cos(phi)
        poses.append(pose(x, y, z, yaw=-phi))
    return poses


def loop_trajectory(n_per_lap: int, laps: int, side=180.0, z=0.0, seed_key=0):
    """Square street loop (side metres, centred on street centre-lines of make_blocks_scene) driven
    `laps` times with a small lateral offset per lap -> every place is revisited `laps` times."""
    r = rng_for(3000, seed_key)
    poses = []
    per = 4.0 * side
    for lap in range(laps):
        for i in range(n_per_lap):
            s = (i + 0.37 * lap) * per / n_per_lap % per
            e, u = int(s // side), s % side
            h = side / 2.0
            if e == 0:
                x, y, yaw = -h + u, -h, 0.0
            elif e == 1:
                x, y, yaw = h, -h + u, math.pi / 2
            elif e == 2:
                x, y, yaw = h - u, h, math.pi
            else:
                x, y, yaw = -h, h - u, -math.pi / 2
            poses.append(pose(x + r.normal(0, 0.3), y + r.normal(0, 0.3), z, yaw + r.normal(0, 0.02)))
    return poses
