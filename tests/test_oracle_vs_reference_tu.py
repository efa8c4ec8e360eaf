"""The oracle (and the product's host halves) against the reference's OWN translation units.

oracle/_ref/libglim_ref.so = /root/reference/src/glim/common/cloud_covariance_estimation.cpp and cloud_deskewing.cpp compiled
UNMODIFIED, from where they lie, against the stand-in headers of oracle/ref_shim/ (recipe: `make -C oracle _ref`; Eigen, GTSAM,
spdlog and gtsam_points are not installed here, the stand-ins are written in this repository).  What that pins: the control flow
and formulas of those two files -- neighbour indexing, population covariance, the PLANE regularization constants, (3,3) = 0, the
normal's flip towards the sensor, the 0.1 ms time table, the IMU cursor / clamp / held-last-pose rules, the composition order of the
transforms.  What it does not: Eigen's and GTSAM's own arithmetic (direct 3x3 eigen-solver, quaternion slerp, Pose3::Expmap), which
the stand-ins restate from the published algorithms.  The VGICP factor itself lives in gtsam_points and stays unpinned.

The .so is a build product (git-ignored); it is built in the container that has /root/reference and travels to the GPU box.
Without it (fresh checkout elsewhere) these tests skip."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from glim_b200 import synth
from oracle import oracle
from tests.test_deskew import IMU_P, IMU_T, T_IL, V, W, scan_like
from tests.util import scan_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libglim_ref.so")


@pytest.fixture(scope="module")
def ref():
    if os.path.isdir("/root/reference/src/glim/common"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "_ref"])
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libglim_ref.so is not built (needs /root/reference: make -C oracle _ref)")
    L = C.CDLL(SO)
    vp, i32 = C.c_void_p, C.c_int
    L.ref_covariance_estimate.argtypes = [i32, vp, vp, i32, i32, i32, vp, vp]
    L.ref_covariance_estimate_sample.argtypes = [i32, vp, vp, i32, i32, vp]
    L.ref_deskew_const_vel.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    L.ref_deskew_imu.argtypes = [vp, i32, vp, vp, C.c_double, i32, vp, vp, vp]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_cov(L, pts, nb, k=None, threads=2):
    n, kc = nb.shape
    k = kc if k is None else k
    pts = np.ascontiguousarray(pts, np.float64)
    nb = np.ascontiguousarray(nb, np.int32)
    normals, covs = np.zeros((n, 4)), np.zeros((n, 16))
    L.ref_covariance_estimate(n, _p(pts), _p(nb), kc, k, threads, _p(normals), _p(covs))
    return normals, covs.reshape(n, 4, 4).transpose(0, 2, 1).copy()


def well_conditioned(pts, nb, k, rel_gap=1e-3):
    """neighbourhoods whose smallest eigenvalue is separated: the eigenvector is then unique up to sign, and two
    solvers agree to ~1e-13 / rel_gap (the covariance itself carries ~1e-13 of cancellation error at 40 m range)"""
    P = pts[nb[:, :k], :3]
    c = np.einsum("nki,nkj->nij", P - P.mean(1, keepdims=True), P - P.mean(1, keepdims=True)) / k
    w = np.linalg.eigvalsh(c)
    return (w[:, 1] - w[:, 0]) > rel_gap * np.maximum(w[:, 2], 1e-300)


@pytest.mark.parametrize("k", [10, 5])
def test_covariance_oracle_equals_reference_translation_unit(ref, k):
    sp = scan_pair()
    pts = sp["points"][0]
    nb = synth.knn(pts, 10).astype(np.int32)
    n_ref, c_ref = ref_cov(ref, pts, nb, k)
    n_orc, c_orc = oracle.covariance_estimate(pts, nb, k_neighbors=k, num_threads=2)
    ok = well_conditioned(pts, nb, k)
    assert ok.mean() > 0.9
    assert np.abs(c_ref[ok] - c_orc[ok]).max() < 1e-10
    assert np.abs(n_ref[ok] - n_orc[ok]).max() < 1e-10  # same sign: both flip towards the sensor origin
    loose = well_conditioned(pts, nb, k, 1e-6)           # nearly isotropic in the plane's normal direction: 1e-13 / gap
    assert np.abs(c_ref[loose] - c_orc[loose]).max() < 1e-6
    # invariants the reference file enforces on EVERY point, conditioned or not
    for c, nr in ((c_ref, n_ref), (c_orc, n_orc)):
        assert np.all(c[:, 3, :] == 0) and np.all(c[:, :, 3] == 0) and np.all(nr[:, 3] == 0)
        assert np.all(np.einsum("ni,ni->n", pts, nr) <= 1e-12)
        assert np.allclose(np.linalg.eigvalsh(c[:, :3, :3]), [1e-3, 1.0, 1.0], atol=1e-9)


def test_plane_regularization_makes_the_divisor_irrelevant(ref):
    """the (points, neighbors, k) overload divides by k - 1 (cloud_covariance_estimation.cpp:153) -- irrelevant under
    PLANE because only the eigenvectors survive.  Checked on the reference code itself."""
    sp = scan_pair()
    pts = sp["points"][1]
    nb = synth.knn(pts, 10).astype(np.int32)
    _, c_pop = ref_cov(ref, pts, nb)
    c_smp = np.zeros((len(pts), 16))
    ref.ref_covariance_estimate_sample(len(pts), _p(np.ascontiguousarray(pts)), _p(nb), 10, 10, _p(c_smp))
    c_smp = c_smp.reshape(-1, 4, 4).transpose(0, 2, 1)
    ok = well_conditioned(pts, nb, 10)
    assert np.abs(c_pop[ok] - c_smp[ok]).max() < 1e-9


def ref_deskew_cv(L, T, v, w, times, pts):
    out = np.empty_like(pts)
    L.ref_deskew_const_vel(_p(oracle.pose_colmajor(T)), _p(np.asarray(v, np.float64)), _p(np.asarray(w, np.float64)), len(times), _p(times), _p(pts), _p(out))
    return out


def ref_deskew_imu(L, T, imu_t, imu_p, stamp, times, pts):
    out = np.empty_like(pts)
    poses = np.ascontiguousarray(np.swapaxes(np.asarray(imu_p, np.float64).reshape(-1, 4, 4), 1, 2)).reshape(-1, 16)
    imu_t = np.ascontiguousarray(imu_t, np.float64)
    L.ref_deskew_imu(_p(oracle.pose_colmajor(T)), len(imu_t), _p(imu_t), _p(poses), float(stamp), len(times), _p(times), _p(pts), _p(out))
    return out


def test_deskew_oracle_equals_reference_translation_unit(ref):
    times, pts = scan_like(seed=3)
    for v, w in ((V, W), (V, np.zeros(3)), (np.zeros(3), W), (np.zeros(3), np.zeros(3))):
        assert np.abs(ref_deskew_cv(ref, T_IL, v, w, times, pts) - oracle.deskew_const_vel(T_IL, v, w, times, pts)).max() < 1e-11
    assert np.abs(ref_deskew_imu(ref, T_IL, IMU_T, IMU_P, 100.0, times, pts) - oracle.deskew_imu(T_IL, IMU_T, IMU_P, 100.0, times, pts)).max() < 1e-11
    # IMU poses that end before the scan does: the last pose is held (cloud_deskewing.cpp:104-105)
    assert np.abs(ref_deskew_imu(ref, T_IL, IMU_T[:6], IMU_P[:6], 100.0, times, pts) - oracle.deskew_imu(T_IL, IMU_T[:6], IMU_P[:6], 100.0, times, pts)).max() < 1e-11
    # a scan that starts before the first IMU pose: p is clamped to 0 (:111)
    assert np.abs(ref_deskew_imu(ref, T_IL, IMU_T + 0.05, IMU_P, 100.0, times, pts) - oracle.deskew_imu(T_IL, IMU_T + 0.05, IMU_P, 100.0, times, pts)).max() < 1e-11


def test_product_host_pose_table_equals_reference_translation_unit(ref):
    """gb_deskew_pose_table (the product's host half, no device needed) applied in numpy == the reference's deskewed points"""
    from glim_b200 import preprocess

    times, pts = scan_like(seed=4)
    idx, Ts = preprocess.deskew_pose_table(T_IL, times, linear_vel=V, angular_vel=W)
    assert np.abs(np.einsum("nij,nj->ni", Ts[idx], pts) - ref_deskew_cv(ref, T_IL, V, W, times, pts)).max() < 1e-11
    idx, Ts = preprocess.deskew_pose_table(T_IL, times, imu_times=IMU_T, imu_poses=IMU_P, stamp=100.0)
    assert np.abs(np.einsum("nij,nj->ni", Ts[idx], pts) - ref_deskew_imu(ref, T_IL, IMU_T, IMU_P, 100.0, times, pts)).max() < 1e-11


# ------------------------------------------------------------------------------------------------------------------------
# CloudPreprocessor::preprocess: the reference's cloud_preprocessor.cpp, compiled unmodified.  Its gtsam_points leaf calls
# (voxelgrid_sampling, randomgrid_sampling, sample, filter, remove_outliers, KdTree) are [EXT] and resolve to stand-ins that forward to
# the oracle's leaves -- so what this checks is the reference file's OWN logic against the composition the GPU parity tests use as
# their oracle (`_cpu_frame` in tests/test_gpu_parity.py): which stage runs when, the finite / range gates, the sort key, global
# shutter, the crop box in either frame, where outlier removal sits, scan_end_time, the k-NN layout and its self-fill.
# ------------------------------------------------------------------------------------------------------------------------
class RefParams(C.Structure):
    _fields_ = [("distance_near_thresh", C.c_double), ("distance_far_thresh", C.c_double), ("use_random_grid_downsampling", C.c_int), ("downsample_resolution", C.c_double),
                ("downsample_target", C.c_int), ("downsample_rate", C.c_double), ("seed", C.c_ulonglong), ("global_shutter", C.c_int), ("crop_bbox_frame", C.c_int),
                ("crop_bbox_min", C.c_double * 3), ("crop_bbox_max", C.c_double * 3), ("T_imu_lidar", C.c_double * 16), ("enable_outlier_removal", C.c_int),
                ("outlier_removal_k", C.c_int), ("outlier_std_mul_factor", C.c_double), ("k_correspondences", C.c_int), ("num_threads", C.c_int)]


def ref_preprocess(L, P, T, stamp=10.0, intensities=None, **kw):
    par = RefParams(distance_near_thresh=1.0, distance_far_thresh=60.0, use_random_grid_downsampling=0, downsample_resolution=0.2, downsample_target=0, downsample_rate=0.3,
                    seed=5, global_shutter=0, crop_bbox_frame=0, enable_outlier_removal=0, outlier_removal_k=10, outlier_std_mul_factor=1.0, k_correspondences=10, num_threads=2)
    par.T_imu_lidar = (C.c_double * 16)(*oracle.pose_colmajor(np.eye(4)))
    for k, v in kw.items():
        if k in ("crop_bbox_min", "crop_bbox_max"):
            v = (C.c_double * 3)(*v)
        elif k == "T_imu_lidar":
            v = (C.c_double * 16)(*oracle.pose_colmajor(v))
        setattr(par, k, v)
    n = len(P)
    P, T = np.ascontiguousarray(P, np.float64), np.ascontiguousarray(T, np.float64)
    op, ot, oi, nb = np.zeros((n, 4)), np.zeros(n), np.zeros(n), np.zeros((n, par.k_correspondences), np.int32)
    end, seen = C.c_double(), (C.c_double * 6)()
    L.ref_preprocess.restype = C.c_int
    L.ref_preprocess.argtypes = [C.c_void_p, C.c_double, C.c_int] + [C.c_void_p] * 9
    it = np.ascontiguousarray(intensities, np.float64) if intensities is not None else None
    m = L.ref_preprocess(C.byref(par), stamp, n, _p(P), _p(T), _p(it) if it is not None else None, _p(op), _p(ot), _p(oi), _p(nb), C.byref(end), seen)
    return op[:m], ot[:m], oi[:m], nb[:m], end.value, list(seen)


@pytest.fixture(scope="module")
def frame():
    sc = synth.make_hall_scene()
    P, T = synth.scan(sc, "hdl32", synth.arc_trajectory(8)[2], synth.rng_for(41), n_rays=32 * 250)
    P = P.copy()
    P[17, 1] = np.nan  # dropped by the finite gate (cloud_preprocessor.cpp:123)
    return P, T


def canonical(pts, tms, nb, extra=None):
    """The reference orders by time with std::sort (NOT stable: the order of equal times is unspecified) and the 32
    rings of a firing share a timestamp, so frames are compared modulo the order inside a group of equal times: rows sorted by
    (time, x, y, z), neighbour indices renumbered accordingly, every neighbour row as a sorted set."""
    perm = np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0], tms))
    inv = np.empty(len(perm), np.int64)
    inv[perm] = np.arange(len(perm))
    out = [pts[perm], tms[perm], np.sort(inv[nb[perm]], axis=1)]
    if extra is not None:
        out.append(extra[perm])
    return out


def test_preprocess_composition_equals_reference_translation_unit(ref, frame):
    from tests.test_gpu_parity import _cpu_frame

    P, T = frame
    lo, hi = np.array([-3.0, -2.0, -5.0]), np.array([4.0, 2.5, 5.0])
    cases = {
        "voxelgrid": (dict(), dict()),
        "cropbox lidar": (dict(crop_bbox_frame=1, crop_bbox_min=lo, crop_bbox_max=hi), dict(crop=(lo, hi))),
        "outliers": (dict(enable_outlier_removal=1, outlier_removal_k=8, outlier_std_mul_factor=1.0), dict(sor=(8, 1.0))),
    }
    for name, (rkw, okw) in cases.items():
        pts, tms, _, nb, end, seen = ref_preprocess(ref, P, T, **rkw)
        o_pts, o_tms, o_nb, _, _ = _cpu_frame(P, T, 0.2, 1.0, 60.0, 10, **okw)
        assert len(pts) == len(o_pts) > 1000, name
        assert np.array_equal(tms, o_tms), name  # the sequence of times is order-free
        for a, b in zip(canonical(pts, tms, nb), canonical(o_pts, o_tms, o_nb)):
            assert np.array_equal(a, b), name
        assert end == 10.0 + o_tms[-1], name
    # the code defaults of CloudPreprocessorParams (cloud_preprocessor.cpp:27-36,58) as the product's gb_preprocess_default_params
    # documents them next to the shipped config
    assert seen == [1.0, 100.0, 0.15, 0.3, 2.0, 8.0]
    # random grid: rate = target / N (cloud_preprocessor.cpp:105), survivors keep their order, then the same gates
    target = 2000
    pts, tms, _, nb, _, _ = ref_preprocess(ref, P, T, use_random_grid_downsampling=1, downsample_resolution=1.0, downsample_target=target, seed=5)
    mask = oracle.randomgrid_sampling(P, 1.0, target / len(P), seed=5)
    o_pts, o_tms, o_nb, _, _ = _cpu_frame(P, T, None, 1.0, 60.0, 10, mask=mask)
    assert len(pts) == len(o_pts) > 500
    for a, b in zip(canonical(pts, tms, nb), canonical(o_pts, o_tms, o_nb)):
        assert np.array_equal(a, b)


def test_preprocess_global_shutter_intensities_and_imu_crop_box(ref, frame):
    P, T = frame
    inten = np.linspace(0.0, 1.0, len(P))
    pts, tms, it, nb, end, _ = ref_preprocess(ref, P, T, intensities=inten, global_shutter=1)
    # global shutter: the points are still ordered by their ORIGINAL times, then every time becomes 0 (:135-140) -> scan_end_time = stamp
    o_pts, o_tms, o_it = oracle.voxelgrid_sampling(P, 0.2, times=T, intensities=inten)
    sq = (o_pts[:, :3] ** 2).sum(1)
    keep = np.nonzero((sq > 1.0) & (sq < 3600.0) & np.isfinite(o_pts).all(1))[0]
    keep = keep[np.argsort(o_tms[keep], kind="stable")]
    assert np.all(tms == 0.0) and end == 10.0
    zero_nb = np.zeros((len(pts), 1), np.int64)
    got, want = canonical(pts, o_tms[keep], zero_nb, it), canonical(o_pts[keep], o_tms[keep], zero_nb, o_it[keep])
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[3], want[3])  # intensities travel with the points
    # crop box given in the IMU frame: p_imu = T_imu_lidar * p_lidar (:152-157); points INSIDE the box are removed
    T_il = synth.pose(0.5, -0.2, 0.1, 0.4, 0.0, 0.0)
    lo, hi = np.array([-3.0, -2.0, -5.0]), np.array([4.0, 2.5, 5.0])
    pts2, _, _, _, _, _ = ref_preprocess(ref, P, T, crop_bbox_frame=2, crop_bbox_min=lo, crop_bbox_max=hi, T_imu_lidar=T_il)
    base = o_pts[keep]
    p_imu = base[:, :3] @ T_il[:3, :3].T + T_il[:3, 3]
    inside = (p_imu >= lo).all(1) & (p_imu <= hi).all(1)
    assert 0 < inside.sum() < len(base)
    assert np.array_equal(pts2[np.lexsort(pts2[:, :3].T)], base[~inside][np.lexsort(base[~inside][:, :3].T)])


# ------------------------------------------------------------------------------------------------------------------------
# The stand-ins are restatements too: check the three non-trivial ones against independent implementations, so that "the reference
# file compiled against a stand-in" cannot hide a wrong stand-in.
# ------------------------------------------------------------------------------------------------------------------------
def test_stand_in_primitives_eigen_solver(ref):
    rng = np.random.default_rng(11)
    worst = 0.0
    for trial in range(400):
        Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        w = np.sort(rng.uniform(1e-4, 5.0, 3))
        if trial % 5 == 0:
            w[0] = w[1] * 1e-6  # a flat neighbourhood: smallest eigenvalue far below the others
        A = (Q * w) @ Q.T
        A = 0.5 * (A + A.T) * rng.choice([1e-3, 1.0, 1e3])
        vals, vecs = np.zeros(3), np.zeros(9)
        ref.ref_shim_eigen_sym3(_p(np.ascontiguousarray(A.T.reshape(9))), _p(vals), _p(vecs))
        V = vecs.reshape(3, 3).T
        ew, eV = np.linalg.eigh(A)
        scale = np.abs(ew).max()
        assert np.all(np.diff(vals) >= -1e-12 * scale)  # ascending, like Eigen
        assert np.abs(vals - ew).max() < 1e-9 * scale
        assert np.abs(V.T @ V - np.eye(3)).max() < 1e-9 and np.linalg.det(V) > 0  # orthonormal, right-handed (col1 = col2 x col0)
        gap = np.min(np.diff(ew)) / scale
        if gap > 1e-3:
            err = max(min(np.abs(V[:, k] - eV[:, k]).max(), np.abs(V[:, k] + eV[:, k]).max()) for k in range(3))
            worst = max(worst, err)
    assert worst < 1e-9


def test_stand_in_primitives_slerp_and_expmap(ref):
    from scipy.spatial.transform import Rotation as Rot, Slerp

    rng = np.random.default_rng(12)
    for _ in range(200):
        r0, r1 = Rot.from_rotvec(rng.normal(size=3) * rng.uniform(0, 2.5)), Rot.from_rotvec(rng.normal(size=3) * rng.uniform(0, 2.5))
        t = float(rng.uniform(0, 1))
        out = np.zeros(9)
        ref.ref_shim_slerp.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        ref.ref_shim_slerp(_p(np.ascontiguousarray(r0.as_matrix().T.reshape(9))), _p(np.ascontiguousarray(r1.as_matrix().T.reshape(9))), t, _p(out))
        want = Slerp([0, 1], Rot.from_matrix([r0.as_matrix(), r1.as_matrix()]))([t])[0].as_matrix()
        assert np.abs(out.reshape(3, 3).T - want).max() < 1e-12
    for _ in range(200):
        xi = np.concatenate([rng.normal(size=3) * rng.choice([0.0, 1e-9, 0.3, 2.0]), rng.normal(size=3) * 3.0])
        T = np.zeros(16)
        ref.ref_shim_pose3_expmap(_p(xi), _p(T))
        assert np.abs(T.reshape(4, 4).T - synth.se3_exp(xi)).max() < 1e-12  # synth.se3_exp: the closed form with the V(omega) matrix
