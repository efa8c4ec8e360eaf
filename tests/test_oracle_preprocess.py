"""Pins the oracle's voxel map and preprocess restatements against independent numpy / scipy implementations
(the reference ships no fixtures; see oracle/glim_oracle.c header)."""
import numpy as np
import pytest

from oracle import oracle
from glim_b200 import synth
from tests import util


@pytest.fixture(scope="module")
def pair():
    return util.scan_pair()


# ---------------------------------------------------------------------------------------------- voxel map (GPU layout)
def test_voxel_hash_widths_agree():
    """u64 XOR-of-primes hash modulo a power-of-two table == u32 evaluation (what the CUDA kernels use), incl. negatives."""
    rng = synth.rng_for(11)
    c = rng.integers(-(1 << 20), 1 << 20, size=(2000, 3))
    for x, y, z in c[:500]:
        h64 = oracle.lib().go_voxel_hash(int(x), int(y), int(z))
        h32 = ((int(x) * 73856093) & 0xFFFFFFFF) ^ ((int(y) * 19349669) & 0xFFFFFFFF) ^ ((int(z) * 83492791) & 0xFFFFFFFF)
        for bits in (14, 20, 28):
            assert (h64 & ((1 << bits) - 1)) == (h32 & ((1 << bits) - 1))


@pytest.mark.parametrize("res", [0.1, 0.25, 0.5, 1.0])
def test_gpumap_build_matches_numpy(pair, res):
    xyz, cov6 = oracle.pack_cloud(pair["points"][0], util.cov_colmajor16(pair["covs"][0]))
    m = oracle.GpuMap(xyz, cov6, res)
    inv = np.float32(1.0) / np.float32(res)
    coords = np.floor(xyz * inv).astype(np.int64)
    uniq, inverse, counts = np.unique(coords, axis=0, return_inverse=True, return_counts=True)  # lexicographic == packed-key order
    inverse = inverse.reshape(-1)
    assert m.num_voxels == len(uniq)
    assert np.array_equal(m.vcoord, uniq.astype(np.int32))
    assert np.array_equal(m.vnum, counts.astype(np.int32))
    # means / covs: fp32 sequential sums in point order -> compare against fp64 sums with fp32 tolerance
    mean64 = np.zeros((len(uniq), 3))
    np.add.at(mean64, inverse, xyz.astype(np.float64))
    mean64 /= counts[:, None]
    cov64 = np.zeros((len(uniq), 6))
    np.add.at(cov64, inverse, cov6.astype(np.float64))
    cov64 /= counts[:, None]
    assert np.allclose(m.vmean, mean64, rtol=0, atol=2e-5)
    assert np.allclose(m.vcov, cov64, rtol=0, atol=1e-5)
    # table invariants: power-of-two size >= init, every stored voxel is found again by the lookup rule (B.4)
    assert m.num_buckets >= 16384 and (m.num_buckets & (m.num_buckets - 1)) == 0
    stored = m.buckets[m.buckets[:, 3] >= 0]
    assert len(np.unique(stored[:, 3])) == len(stored)
    dropped_voxels = m.num_voxels - len(stored)
    assert m.num_dropped_points <= 1e-3 * len(xyz) + 1e-9
    assert np.array_equal(stored[:, :3], m.vcoord[stored[:, 3]])
    corr = m.correspondences(xyz, np.eye(4))
    assert (corr >= 0).sum() == len(xyz) - m.num_dropped_points
    hit = corr >= 0
    assert np.array_equal(corr[hit], inverse[hit])
    assert dropped_voxels >= 0


def test_gpumap_boundary_points_and_negative_coords():
    """Points exactly on voxel faces, on both sides of zero, with a power-of-two and a non-power-of-two resolution."""
    res = 0.5
    g = np.arange(-6, 7) * res
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
    pts = np.concatenate([pts, np.nextafter(pts, np.float32(-np.inf)), np.nextafter(pts, np.float32(np.inf))])
    cov = np.tile(np.array([1, 0, 0, 1, 0, 1], np.float32), (len(pts), 1))
    for r in (0.5, 0.1):
        m = oracle.GpuMap(pts, cov, r)
        inv = np.float32(1.0) / np.float32(r)
        coords = np.floor(pts * inv).astype(np.int32)
        corr = m.correspondences(pts, np.eye(4))
        assert (corr >= 0).all()
        assert np.array_equal(m.vcoord[corr], coords)


def test_gpumap_grows_when_points_are_dropped():
    """More voxels than init buckets forces the doubling loop (drop-rate rule, B.3)."""
    rng = synth.rng_for(12)
    pts = rng.uniform(-200, 200, size=(60000, 3)).astype(np.float32)
    cov = np.tile(np.array([1, 0, 0, 1, 0, 1], np.float32), (len(pts), 1))
    m = oracle.GpuMap(pts, cov, 0.5, init_buckets=1024)
    assert m.num_buckets > 1024 and m.num_buckets >= m.num_voxels
    assert m.num_dropped_points <= 1e-3 * len(pts)
    m2 = oracle.GpuMap(pts, cov, 0.5, init_buckets=1024, drop_rate=1.0)  # tolerate everything: no growth beyond >= V
    assert m2.num_buckets <= m.num_buckets


def test_empty_and_degenerate_clouds():
    z3, z6 = np.zeros((0, 3), np.float32), np.zeros((0, 6), np.float32)
    m = oracle.GpuMap(z3, z6, 0.5)
    assert m.num_voxels == 0 and m.num_buckets == 16384
    out, corr = oracle.linearize_gpumap(m, z3, z6, np.eye(4))
    assert not out.any() and corr.size == 0
    one = np.array([[1.0, 2.0, 3.0]], np.float32)
    c1 = np.array([[1, 0, 0, 1, 0, 1]], np.float32)
    m = oracle.GpuMap(one, c1, 0.5)
    out, corr = oracle.linearize_gpumap(m, one, c1, np.eye(4))
    assert corr[0] == 0 and out[121] == 1.0 and out[120] == 0.0  # residual is exactly zero
    nan = np.array([[np.nan, 0, 0], [1, 1, 1]], np.float32)
    m = oracle.GpuMap(nan, np.tile(c1, (2, 1)), 0.5)
    assert m.num_voxels == 1


# ---------------------------------------------------------------------------------------------- CPU map
def test_cpumap_matches_numpy_and_incremental_insert(pair):
    P, C16 = pair["points"][0], util.cov_colmajor16(pair["covs"][0])
    m = oracle.CpuMap(0.5)
    m.insert(P, C16)
    coords = np.floor(P[:, :3] * (1.0 / 0.5)).astype(np.int64)
    uniq, inverse, counts = np.unique(coords, axis=0, return_inverse=True, return_counts=True)
    inverse = inverse.reshape(-1)
    assert m.num_voxels == len(uniq)
    mean = np.zeros((len(uniq), 4))
    np.add.at(mean, inverse, P)
    mean /= counts[:, None]
    for i in (0, 17, len(P) - 1):
        j, mu, cov = m.lookup(P[i, :3])
        assert j >= 0 and np.allclose(mu, mean[inverse[i]], atol=1e-12)
    # inserting in two halves gives the same voxels (re-open / re-finalize, B.5)
    m2 = oracle.CpuMap(0.5)
    h = len(P) // 2
    m2.insert(P[:h], C16[:h])
    m2.insert(P[h:], C16[h:])
    assert m2.num_voxels == m.num_voxels
    for i in (3, h, len(P) - 2):
        _, mu1, c1 = m.lookup(P[i, :3])
        _, mu2, c2 = m2.lookup(P[i, :3])
        assert np.allclose(mu1, mu2, atol=1e-12) and np.allclose(c1, c2, atol=1e-12)


# ---------------------------------------------------------------------------------------------- eigen / covariance
def test_eigen_sym3_direct_matches_numpy():
    rng = synth.rng_for(13)
    for trial in range(200):
        A = rng.normal(size=(3, 3))
        A = A @ A.T * 10 ** rng.uniform(-6, 3)
        ev, V = oracle.eigen_sym3(A)
        w, _ = np.linalg.eigh(A)
        scale = np.abs(w).max()
        assert np.allclose(ev, w, rtol=0, atol=1e-9 * scale)
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-9)
        assert np.allclose(A @ V, V * ev, atol=1e-7 * scale)
    # degenerate: isotropic and rank-1
    ev, V = oracle.eigen_sym3(np.eye(3) * 2.5)
    assert np.allclose(ev, 2.5) and np.allclose(V.T @ V, np.eye(3))
    n = np.array([1.0, 2.0, -1.0]) / np.sqrt(6.0)
    ev, V = oracle.eigen_sym3(np.outer(n, n))
    assert np.allclose(ev, [0, 0, 1], atol=1e-12) and abs(abs(V[:, 2] @ n) - 1) < 1e-9


def test_covariance_estimate_matches_numpy(pair):
    """Line-by-line restatement (cloud_covariance_estimation.cpp:43-122) vs the numpy/eigh implementation in synth."""
    P = pair["points"][0]
    nb = synth.knn(P, 10)
    normals, covs = oracle.covariance_estimate(P, nb)
    n_ref, c_ref = synth.plane_covariances(P, nb)
    # cov == I - 0.999 n n^T (SURVEY C.4 [DERIVED]); sign-free comparison through n n^T
    nn = np.einsum("ni,nj->nij", normals[:, :3], normals[:, :3])
    assert np.allclose(covs[:, :3, :3], np.eye(3) - 0.999 * nn, atol=1e-9)
    assert (covs[:, 3, :] == 0).all() and (covs[:, :, 3] == 0).all()
    # compare with numpy where the smallest eigenvalue is well separated
    P_nb = P[nb]
    S = P_nb.sum(1)
    X = np.einsum("nki,nkj->nij", P_nb, P_nb)
    cov_raw = (X - (S / 10)[:, :, None] * S[:, None, :]) / 10
    w = np.linalg.eigvalsh(cov_raw[:, :3, :3])
    good = (w[:, 1] - w[:, 0]) > 1e-6 * w[:, 2]
    assert good.mean() > 0.9
    assert np.allclose(covs[good], c_ref[good], atol=1e-6)
    # normals point towards the sensor origin (p . n <= 0, :98-101) and are unit length
    assert (np.einsum("ni,ni->n", P[:, :3], normals[:, :3]) <= 1e-12).all()
    assert np.allclose(np.linalg.norm(normals[:, :3], axis=1), 1.0, atol=1e-9)
    assert np.allclose(np.abs(np.einsum("ni,ni->n", normals[good, :3], n_ref[good, :3])), 1.0, atol=1e-6)
    # k_neighbors < k_correspondences uses the first k columns (:84-88)
    n5, c5 = oracle.covariance_estimate(P, nb, k_neighbors=5)
    n5b, c5b = oracle.covariance_estimate(P, np.ascontiguousarray(nb[:, :5]))
    assert np.array_equal(c5, c5b) and np.array_equal(n5, n5b)
    # the invariants SubMap::load checks (src/glim/mapping/sub_map.cpp:150-167)
    ev = np.linalg.eigvalsh(covs[:, :3, :3])
    assert (ev > 1e-6).all() and (ev < 1e6).all() and np.isfinite(covs).all()


def test_covariance_empty_input():
    n, c = oracle.covariance_estimate(np.zeros((0, 4)), np.zeros((0, 10), np.int32))
    assert n.shape == (0, 4) and c.shape == (0, 4, 4)


# ---------------------------------------------------------------------------------------------- kNN / voxel grid
def test_knn_matches_ckdtree():
    P = util.scan_pair(n_rays=32 * 100)["points"][0]
    nb, d = oracle.knn_bruteforce(P, 10)
    ref = synth.knn(P, 10)
    assert (nb[:, 0] == np.arange(len(P))).all() and (d[:, 0] == 0).all()  # query itself first (:196)
    assert (np.diff(d, axis=1) >= 0).all()
    # same neighbour SET wherever the k-th and (k+1)-th distances are separated
    from scipy.spatial import cKDTree

    dd, _ = cKDTree(P[:, :3]).query(P[:, :3], k=11)
    sep = (dd[:, 10] - dd[:, 9]) > 1e-9
    same = np.array([set(a) == set(b) for a, b in zip(nb, ref)])
    assert same[sep].all() and sep.mean() > 0.95
    # fewer points than k: remaining slots keep the query index (:196-199)
    nb3, _ = oracle.knn_bruteforce(P[:4], 10)
    assert (nb3[:, 4:] == np.arange(4)[:, None]).all()


def test_voxelgrid_sampling_matches_numpy():
    P = util.scan_pair()["points"][0]
    T = util.scan_pair()["times"][0]
    out, ot, _ = oracle.voxelgrid_sampling(P, 0.25, times=T)
    coords = np.floor(P[:, :3] * (1.0 / 0.25)).astype(np.int64)
    uniq, inverse, counts = np.unique(coords, axis=0, return_inverse=True, return_counts=True)
    inverse = inverse.reshape(-1)
    mean = np.zeros((len(uniq), 4))
    np.add.at(mean, inverse, P)
    mean /= counts[:, None]
    tm = np.zeros(len(uniq))
    np.add.at(tm, inverse, T)
    tm /= counts
    assert out.shape == mean.shape
    assert np.allclose(out, mean, atol=1e-12) and np.allclose(ot, tm, atol=1e-12)  # ascending packed key == lexicographic
    assert np.allclose(out[:, 3], 1.0)
    e, _, _ = oracle.voxelgrid_sampling(np.zeros((0, 4)), 0.25)
    assert e.shape == (0, 4)
