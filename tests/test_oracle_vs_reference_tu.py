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
