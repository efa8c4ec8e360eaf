"""Pins the oracle functions added in round 2 (oracle/glim_oracle.c: go_randomgrid_sampling, go_merge_frames,
go_vgicp_linearize_gpumap_sv) against independent numpy restatements of the documented rules -- the CUDA
kernels are checked against these oracle functions by the -m gpu tests, so they must not be self-referential."""
import numpy as np
import pytest

from oracle import oracle
from glim_b200 import synth
from tests import util
from tests.test_oracle_vgicp import numpy_linearize

U64 = np.uint64


def splitmix_hash(seed, i):
    """hash(seed, index) of the ledger: splitmix64 finaliser of seed + golden * (i + 1), in wrapping uint64 arithmetic."""
    with np.errstate(over="ignore"):
        z = U64(seed) + U64(0x9E3779B97F4A7C15) * (np.asarray(i, dtype=U64) + U64(1))
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def numpy_randomgrid(pts4, res, rate, seed):
    n = len(pts4)
    if rate >= 0.99:
        return np.ones(n, bool)
    c = np.floor(pts4[:, :3] * (1.0 / res)).astype(np.int64)
    _, inv = np.unique(c, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    V = inv.max() + 1
    ppv = max(1, int(np.ceil(rate * n / V)))
    h = splitmix_hash(seed, np.arange(n))
    keep = np.zeros(n, bool)
    order = np.lexsort((np.arange(n), h, inv))  # by voxel, then hash, then index
    start = np.r_[0, np.flatnonzero(np.diff(inv[order])) + 1]
    rank = np.arange(n) - np.repeat(start, np.diff(np.r_[start, n]))
    keep[order[rank < ppv]] = True
    cap = int(n * rate * 1.2)
    if keep.sum() > cap > 0:
        thr = np.sort(h[keep])[cap - 1]
        keep &= h <= thr
    return keep


@pytest.mark.parametrize("rate,res", [(0.1, 1.0), (0.3, 0.5), (0.7, 2.0), (1.0, 1.0)])
def test_randomgrid_sampling_matches_numpy(rate, res):
    rng = np.random.default_rng(11)
    # clustered cloud: dense blobs (quota binds) + sparse background (voxels below their quota)
    pts = np.concatenate([rng.normal(size=(4000, 3)) * 0.6 + [5, 0, 0], rng.normal(size=(3000, 3)) * 0.4 - [3, 2, 0], rng.uniform(-20, 20, size=(3000, 3))])
    pts4 = np.concatenate([pts, np.ones((len(pts), 1))], axis=1)
    for seed in (0, 12345):
        got = oracle.randomgrid_sampling(pts4, res, rate, seed)
        assert np.array_equal(got, numpy_randomgrid(pts4, res, rate, seed))
    if rate < 0.99:
        assert 0 < got.sum() <= int(len(pts4) * rate * 1.2)
        c = np.floor(pts * (1.0 / res)).astype(np.int64)
        _, inv, cnt = np.unique(c, axis=0, return_inverse=True, return_counts=True)
        ppv = max(1, int(np.ceil(rate * len(pts) / len(cnt))))
        assert np.bincount(inv.reshape(-1)[got], minlength=len(cnt)).max() <= ppv
        # a different seed draws different points of the crowded voxels
        assert not np.array_equal(got, oracle.randomgrid_sampling(pts4, res, rate, 999))


def _full(c6):
    C = np.zeros((len(c6), 3, 3))
    C[:, 0, 0], C[:, 0, 1], C[:, 0, 2], C[:, 1, 1], C[:, 1, 2], C[:, 2, 2] = c6.T
    C[:, 1, 0], C[:, 2, 0], C[:, 2, 1] = c6[:, 1], c6[:, 2], c6[:, 4]
    return C


def test_merge_frames_matches_numpy():
    pair = util.scan_pair()
    clouds = [oracle.pack_cloud(pair["points"][k], util.cov_colmajor16(pair["covs"][k])) for k in (0, 1)]
    rng = np.random.default_rng(2)
    poses = [synth.pose(*(rng.normal(size=3) * 0.5), *(rng.normal(size=3) * 0.2)), synth.inv_pose(pair["poses"][0]) @ pair["poses"][1]]
    res = 0.5
    pts, covs = oracle.merge_frames(poses, clouds, res)
    # numpy: transform, R C R^T, average per voxel
    P, Cs = [], []
    for (xyz, cov6), T in zip(clouds, poses):
        R, t = T[:3, :3], T[:3, 3]
        P.append(xyz.astype(np.float64) @ R.T + t)
        Cs.append(R @ _full(cov6.astype(np.float64)) @ R.T)
    P, Cs = np.concatenate(P), np.concatenate(Cs)
    key = np.floor(P * (1.0 / res)).astype(np.int64)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    V = len(uniq)
    cnt = np.bincount(inv, minlength=V).astype(np.float64)
    mean_p = np.stack([np.bincount(inv, weights=P[:, k], minlength=V) for k in range(3)], axis=1) / cnt[:, None]
    mean_c = np.stack([np.bincount(inv, weights=Cs[:, r, c], minlength=V) for r in range(3) for c in range(3)], axis=1).reshape(V, 3, 3) / cnt[:, None, None]
    assert len(pts) == V
    # same set of voxels, whatever the order: sort both by the voxel of the averaged point (unique per voxel for a convex cell)
    ko = np.lexsort(np.floor(pts[:, :3] * (1.0 / res)).astype(np.int64).T[::-1])
    kn = np.lexsort(uniq.T[::-1])
    assert np.array_equal(np.floor(pts[ko, :3] * (1.0 / res)).astype(np.int64), uniq[kn])
    assert np.allclose(pts[ko, :3], mean_p[kn], rtol=0, atol=1e-10) and np.all(pts[:, 3] == 1.0)
    assert np.allclose(covs[ko][:, :3, :3], mean_c[kn], rtol=1e-10, atol=1e-14)
    assert np.all(covs[:, 3, :] == 0) and np.all(covs[:, :, 3] == 0)
    # thinning keeps exactly `target` of those voxels, untouched, and is reproducible per seed
    target = V // 3
    tp, tc = oracle.merge_frames(poses, clouds, res, target=target, seed=7)
    assert len(tp) == target
    full = {tuple(p) for p in pts[:, :3]}
    assert all(tuple(p) in full for p in tp[:, :3])
    tp2, _ = oracle.merge_frames(poses, clouds, res, target=target, seed=7)
    assert np.array_equal(tp, tp2)
    tp3, _ = oracle.merge_frames(poses, clouds, res, target=target, seed=8)
    assert not np.array_equal(tp, tp3)


def test_surface_validation_rule_matches_numpy():
    """keep a correspondence iff 3 n^T C_B n <= tr(C_B), n = R n_A (fp32 in oracle and kernel; fp64 here: a handful of
    correspondences within rounding of the threshold may flip, everything else must agree)."""
    pair = util.scan_pair()
    xyz0, cov0 = oracle.pack_cloud(pair["points"][0], util.cov_colmajor16(pair["covs"][0]))
    xyz1, cov1 = oracle.pack_cloud(pair["points"][1], util.cov_colmajor16(pair["covs"][1]))
    normals = np.asarray(pair["normals"][1], np.float64)[:, :3]
    m = oracle.GpuMap(xyz0, cov0, 0.5)
    T_gt = synth.inv_pose(pair["poses"][0]) @ pair["poses"][1]
    roll = np.eye(4)
    a = np.deg2rad(60.0)
    roll[:3, :3] = [[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]]
    for T in (T_gt, T_gt @ roll):
        out_sv, corr_sv = oracle.linearize_gpumap(m, xyz1, cov1, T, normals=normals)
        _, corr = oracle.linearize_gpumap(m, xyz1, cov1, T)
        hit = corr >= 0
        Tf = T.astype(np.float32).astype(np.float64)
        n = normals.astype(np.float32).astype(np.float64) @ Tf[:3, :3].T
        CB = np.zeros((len(xyz1), 3, 3))
        CB[hit] = _full(m.vcov[corr[hit]].astype(np.float64))
        s = np.einsum("ni,nij,nj->n", n, CB, n)
        tr = np.trace(CB, axis1=1, axis2=2)
        margin = np.abs(3.0 * s - tr) / np.maximum(tr, 1e-30)
        keep = hit & (3.0 * s <= tr)
        got_keep = corr_sv >= 0
        assert np.array_equal(corr_sv[~hit], corr[~hit])          # misses stay misses
        assert np.all(corr_sv[hit & ~got_keep] == -2)               # rejected correspondences are marked
        differ = keep != got_keep
        assert np.all(margin[differ] < 1e-5) and differ.sum() <= 3  # only threshold ties may flip (fp32 vs fp64)
        assert 0 < (hit & ~got_keep).sum() < hit.sum()
        # the Hessian is the plain factor restricted to the surviving correspondences
        mu = np.zeros((len(xyz1), 3))
        mu[hit] = m.vmean[corr[hit]]
        P4 = np.concatenate([xyz1.astype(np.float64), np.ones((len(xyz1), 1))], axis=1)
        C4 = np.zeros((len(xyz1), 4, 4))
        C4[:, :3, :3] = _full(cov1.astype(np.float64))
        ref = numpy_linearize(P4, C4, mu, CB, got_keep, Tf)
        o = oracle.split122(out_sv)
        assert o["num_inliers"] == ref["num_inliers"]
        for k in ("H_tt", "H_ss", "H_ts", "b_t", "b_s"):
            assert util.rel_err(o[k], ref[k]) < 1e-10, k
