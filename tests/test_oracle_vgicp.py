"""Pins the oracle's VGICP factor math (oracle/glim_oracle.c) without the reference binary:
an independent vectorised numpy restatement of SURVEY A.2, finite-difference gradients, the adjoint
identities of A.4, and consistency between the CPU-faithful path (IntegratedVGICPFactor, fp64
unordered_map voxel map) and the device-layout path (IntegratedVGICPFactorGPU, fp32 open-addressing map).
"""
import numpy as np
import pytest

from oracle import oracle
from glim_b200 import synth
from tests import util


def numpy_linearize(points, covs, mu_B, C_B, mask, T):
    """A.2 in plain numpy for points with a correspondence (mask), everything fp64."""
    R, t = T[:3, :3], T[:3, 3]
    a = points[mask, :3]
    CA = covs[mask][:, :3, :3]
    q = a @ R.T + t
    r = mu_B[mask] - q
    S = C_B[mask] + R @ CA @ R.T
    M = np.linalg.inv(S)

    def hat(v):
        H = np.zeros((v.shape[0], 3, 3))
        H[:, 0, 1], H[:, 0, 2] = -v[:, 2], v[:, 1]
        H[:, 1, 0], H[:, 1, 2] = v[:, 2], -v[:, 0]
        H[:, 2, 0], H[:, 2, 1] = -v[:, 1], v[:, 0]
        return H

    n = a.shape[0]
    Jt = np.concatenate([-hat(q), np.tile(np.eye(3), (n, 1, 1))], axis=2)
    Js = np.concatenate([R @ hat(a), np.tile(-R, (n, 1, 1))], axis=2)
    Mr = np.einsum("nij,nj->ni", M, r)
    return {
        "H_tt": np.einsum("nki,nkl,nlj->ij", Jt, M, Jt),
        "H_ss": np.einsum("nki,nkl,nlj->ij", Js, M, Js),
        "H_ts": np.einsum("nki,nkl,nlj->ij", Jt, M, Js),
        "b_t": np.einsum("nki,nk->i", Jt, Mr),
        "b_s": np.einsum("nki,nk->i", Js, Mr),
        "error": float(np.einsum("ni,ni->", r, Mr)),
        "num_inliers": float(n),
    }


@pytest.fixture(scope="module")
def pair():
    return util.scan_pair()


@pytest.fixture(scope="module")
def cpu_setup(pair):
    m = oracle.CpuMap(0.5)
    m.insert(pair["points"][0], util.cov_colmajor16(pair["covs"][0]))
    T_gt = synth.inv_pose(pair["poses"][0]) @ pair["poses"][1]
    return m, T_gt


def test_cpu_factor_matches_numpy(pair, cpu_setup):
    m, T_gt = cpu_setup
    P, Cv = pair["points"][1], pair["covs"][1]
    fac = oracle.CpuFactor(m, P, util.cov_colmajor16(Cv), num_threads=1)
    for T in util.test_poses(T_gt, 3):
        out = oracle.split122(fac.linearize(T))
        # gather correspondences through the public lookup
        q = P[:, :3] @ T[:3, :3].T + T[:3, 3]
        mask = np.zeros(len(P), bool)
        mu = np.zeros((len(P), 3))
        CB = np.zeros((len(P), 3, 3))
        for i in range(len(P)):
            j, mean, cov = m.lookup(q[i])
            if j >= 0:
                mask[i], mu[i], CB[i] = True, mean[:3], cov[:3, :3]
        ref = numpy_linearize(P, Cv, mu, CB, mask, T)
        assert out["num_inliers"] == ref["num_inliers"] > 0.5 * len(P)
        for k in ("H_tt", "H_ss", "H_ts", "b_t", "b_s"):
            assert util.rel_err(out[k], ref[k]) < 1e-10, k
        assert abs(out["error"] - ref["error"]) < 1e-9 * abs(ref["error"])


def test_cpu_factor_threads_agree(pair, cpu_setup):
    m, T_gt = cpu_setup
    P, Cv = pair["points"][1], util.cov_colmajor16(pair["covs"][1])
    a = oracle.CpuFactor(m, P, Cv, num_threads=1).linearize(T_gt)
    b = oracle.CpuFactor(m, P, Cv, num_threads=4).linearize(T_gt)
    assert util.rel_err(a, b) < 1e-12


def test_finite_difference_gradient(pair, cpu_setup):
    """With correspondences and Mahalanobis matrices frozen at the linearization point (what the CPU factor's error()
    does), d error / d xi = 2 b for both the source and the target perturbation (right-multiplicative, [rot; trans])."""
    m, T_gt = cpu_setup
    P, Cv = pair["points"][1], util.cov_colmajor16(pair["covs"][1])
    fac = oracle.CpuFactor(m, P, Cv)
    T0 = synth.perturb(T_gt, synth.rng_for(5), 0.01, 0.05)
    out = oracle.split122(fac.linearize(T0))
    h = 1e-6
    g_s, g_t = np.zeros(6), np.zeros(6)
    for k in range(6):
        xi = np.zeros(6)
        xi[k] = h
        # source: T_s <- T_s Exp(xi)  => delta <- delta Exp(xi)
        g_s[k] = (fac.error(T0 @ synth.se3_exp(xi)) - fac.error(T0 @ synth.se3_exp(-xi))) / (2 * h)
        # target: T_t <- T_t Exp(xi)  => delta <- Exp(-xi) delta
        g_t[k] = (fac.error(synth.se3_exp(-xi) @ T0) - fac.error(synth.se3_exp(xi) @ T0)) / (2 * h)
    assert util.rel_err(g_s, 2 * out["b_s"]) < 1e-6
    assert util.rel_err(g_t, 2 * out["b_t"]) < 1e-6


def test_adjoint_identities_and_psd(pair, cpu_setup):
    """J_s = -J_t Ad  =>  H_ss = Ad^T H_tt Ad, H_ts = -H_tt Ad, b_s = -Ad^T b_t."""
    m, T_gt = cpu_setup
    P, Cv = pair["points"][1], util.cov_colmajor16(pair["covs"][1])
    fac = oracle.CpuFactor(m, P, Cv)
    for T in util.test_poses(T_gt, 2):
        o = oracle.split122(fac.linearize(T))
        R, t = T[:3, :3], T[:3, 3]
        Ad = np.zeros((6, 6))
        Ad[:3, :3] = R
        Ad[3:, 3:] = R
        Ad[3:, :3] = synth.hat(t) @ R
        assert util.rel_err(Ad.T @ o["H_tt"] @ Ad, o["H_ss"]) < 1e-10
        assert util.rel_err(-o["H_tt"] @ Ad, o["H_ts"]) < 1e-10
        assert util.rel_err(-Ad.T @ o["b_t"], o["b_s"]) < 1e-10
        for k in ("H_tt", "H_ss"):
            assert np.allclose(o[k], o[k].T, rtol=0, atol=1e-9 * np.abs(o[k]).max())
            assert np.linalg.eigvalsh(o[k]).min() > -1e-9 * np.abs(o[k]).max()


def test_gpumap_path_matches_numpy(pair):
    """Device-layout oracle path: fp32 inputs, fp64 arithmetic, correspondences through the fp32 hash map."""
    xyz0, cov0 = oracle.pack_cloud(pair["points"][0], util.cov_colmajor16(pair["covs"][0]))
    xyz1, cov1 = oracle.pack_cloud(pair["points"][1], util.cov_colmajor16(pair["covs"][1]))
    m = oracle.GpuMap(xyz0, cov0, 0.5)
    T_gt = synth.inv_pose(pair["poses"][0]) @ pair["poses"][1]
    for T in util.test_poses(T_gt, 2):
        out, corr = oracle.linearize_gpumap(m, xyz1, cov1, T)
        o = oracle.split122(out)
        Tf = T.astype(np.float32).astype(np.float64)  # the factor sees the Isometry3f cast (A.1)
        mask = corr >= 0

        def full(c6):
            C = np.zeros((len(c6), 3, 3))
            C[:, 0, 0], C[:, 0, 1], C[:, 0, 2], C[:, 1, 1], C[:, 1, 2], C[:, 2, 2] = c6.T
            C[:, 1, 0], C[:, 2, 0], C[:, 2, 1] = c6[:, 1], c6[:, 2], c6[:, 4]
            return C

        mu = np.zeros((len(xyz1), 3))
        CB = np.zeros((len(xyz1), 3, 3))
        mu[mask] = m.vmean[corr[mask]]
        CB[mask] = full(m.vcov[corr[mask]].astype(np.float64))
        P4 = np.concatenate([xyz1.astype(np.float64), np.ones((len(xyz1), 1))], axis=1)
        C4 = np.zeros((len(xyz1), 4, 4))
        C4[:, :3, :3] = full(cov1.astype(np.float64))
        ref = numpy_linearize(P4, C4, mu, CB, mask, Tf)
        assert o["num_inliers"] == ref["num_inliers"]
        for k in ("H_tt", "H_ss", "H_ts", "b_t", "b_s"):
            assert util.rel_err(o[k], ref[k]) < 1e-10, k
        assert np.array_equal(corr, m.correspondences(xyz1, T))


def test_cpu_and_gpu_paths_agree_loosely(pair, cpu_setup):
    """fp64 unordered_map path vs fp32 device-layout path: same algorithm, different rounding of inputs and
    a few voxel-boundary flips -> agreement to ~1e-3, inlier counts within 0.5 %."""
    m_cpu, T_gt = cpu_setup
    P0, C0 = pair["points"][0], util.cov_colmajor16(pair["covs"][0])
    P1, C1 = pair["points"][1], util.cov_colmajor16(pair["covs"][1])
    xyz0, cov0 = oracle.pack_cloud(P0, C0)
    xyz1, cov1 = oracle.pack_cloud(P1, C1)
    m_gpu = oracle.GpuMap(xyz0, cov0, 0.5)
    assert m_gpu.num_voxels == m_cpu.num_voxels
    a = oracle.split122(oracle.CpuFactor(m_cpu, P1, C1).linearize(T_gt))
    b = oracle.split122(oracle.linearize_gpumap(m_gpu, xyz1, cov1, T_gt)[0])
    assert abs(a["num_inliers"] - b["num_inliers"]) <= 0.005 * a["num_inliers"] + m_gpu.num_dropped_points
    for k in ("H_tt", "H_ss"):
        assert util.rel_err(a[k], b[k]) < 5e-3, k
    # b is a sum of residuals that nearly cancels at the ground-truth pose: a handful of boundary flips moves it
    assert util.rel_err(a["b_s"], b["b_s"]) < 0.1


def test_error_semantics(pair):
    """error(values): inliers from T_lin; at T_eval == T_lin it equals linearize's error; overlap counts hits."""
    xyz0, cov0 = oracle.pack_cloud(pair["points"][0], util.cov_colmajor16(pair["covs"][0]))
    xyz1, cov1 = oracle.pack_cloud(pair["points"][1], util.cov_colmajor16(pair["covs"][1]))
    m = oracle.GpuMap(xyz0, cov0, 0.5)
    T_gt = synth.inv_pose(pair["poses"][0]) @ pair["poses"][1]
    out, corr = oracle.linearize_gpumap(m, xyz1, cov1, T_gt)
    assert abs(oracle.error_gpumap(m, xyz1, cov1, T_gt, T_gt) - out[120]) < 1e-9 * out[120]
    T2 = synth.perturb(T_gt, synth.rng_for(3), 0.02, 0.1)
    assert oracle.error_gpumap(m, xyz1, cov1, T_gt, T2) > out[120]  # moving away from the optimum raises the cost
    ov = oracle.overlap_gpumap([m], xyz1, [T_gt])
    assert ov == pytest.approx((corr >= 0).mean(), abs=1e-12)
    far = T_gt.copy()
    far[:3, 3] += 1000.0
    assert oracle.overlap_gpumap([m], xyz1, [far]) == 0.0
    assert oracle.overlap_gpumap([m, m], xyz1, [far, T_gt]) == pytest.approx(ov)


def test_cpu_voxelmap_lru_horizon():
    """GaussianVoxelMapCPU::set_lru_horizon as OdometryEstimationCPU uses it (odometry_estimation_cpu.cpp:67, lru_thresh): voxels
    that no insert touched for `horizon` inserts are erased at the next clear cycle; touched ones survive with their statistics."""
    import numpy as np

    from oracle import oracle

    rng = np.random.default_rng(1)

    def blob(cx, n=400):
        p = np.concatenate([rng.uniform(-2, 2, (n, 3)) + [cx, 0, 0], np.ones((n, 1))], 1)
        c = np.tile(np.diag([1.0, 1.0, 1.0, 0.0]), (n, 1, 1)).reshape(n, 16)
        return p, c

    m = oracle.CpuMap(0.5)
    m.set_lru_horizon(6, clear_cycle=2)
    keep_p, keep_c = blob(0.0)
    m.insert(keep_p, keep_c)
    n0 = m.num_voxels
    j0, mean0, cov0 = m.lookup(keep_p[0, :3])
    assert j0 >= 0
    sizes = []
    for k in range(1, 14):  # a sensor moving away: blob k is never touched again
        p, c = blob(10.0 * k)
        m.insert(p, c)
        if k % 3 == 0:
            m.insert(keep_p[:50], keep_c[:50])  # ... while the first region keeps being observed
        sizes.append(m.num_voxels)
    assert max(sizes) < n0 * 9  # without eviction it would hold all 14 blobs
    j1, mean1, _ = m.lookup(keep_p[0, :3])
    assert j1 >= 0  # the re-observed region survived
    assert m.lookup(np.array([10.0, 0.0, 0.0]))[0] < 0 or m.lookup(blob(10.0)[0][0, :3])[0] < 0  # the oldest blob is gone
    # eviction off (default): nothing disappears
    m2 = oracle.CpuMap(0.5)
    first = blob(0.0)
    m2.insert(*first)
    for k in range(1, 30):
        m2.insert(*blob(10.0 * k))
    assert m2.lookup(first[0][0, :3])[0] >= 0
