"""Golden vectors (tests/golden/vgicp_golden.npz, produced by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them exactly.  GPU: the CUDA path reproduces them (indices exact, floats to 1e-4)."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests import util

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vgicp_golden.npz"))


def c16(c):
    return np.ascontiguousarray(np.swapaxes(c, 1, 2)).reshape(len(c), 16)


def test_oracle_reproduces_golden():
    xyz0, cov0 = oracle.pack_cloud(G["points0"], c16(G["covs0"]))
    xyz1, cov1 = oracle.pack_cloud(G["points1"], c16(G["covs1"]))
    for res in (0.25, 0.5):
        tag = f"r{int(res * 100):03d}"
        m = oracle.GpuMap(xyz0, cov0, res)
        assert np.array_equal(m.buckets, G[f"{tag}_buckets"]) and np.array_equal(m.vnum, G[f"{tag}_vnum"])
        assert np.array_equal(m.vmean, G[f"{tag}_vmean"]) and np.array_equal(m.vcov, G[f"{tag}_vcov"])
        for k, T in enumerate(G["poses"]):
            o, c = oracle.linearize_gpumap(m, xyz1, cov1, T)
            assert np.array_equal(c, G[f"{tag}_corr"][k])
            assert np.allclose(o, G[f"{tag}_linearized"][k], rtol=1e-12, atol=1e-9)
            assert oracle.error_gpumap(m, xyz1, cov1, G["poses"][0], T) == pytest.approx(G[f"{tag}_error"][k], rel=1e-12)
            assert oracle.overlap_gpumap([m], xyz1, [T]) == G[f"{tag}_overlap"][k]
    cm = oracle.CpuMap(0.5)
    cm.insert(G["points0"], c16(G["covs0"]))
    fac = oracle.CpuFactor(cm, G["points1"], c16(G["covs1"]))
    for k, T in enumerate(G["poses"]):
        assert np.allclose(fac.linearize(T), G["cpu_linearized"][k], rtol=1e-10, atol=1e-8)
    nb, _ = oracle.knn_bruteforce(G["points0"], 10)
    assert np.array_equal(nb, G["knn0"])
    nrm, cv = oracle.covariance_estimate(G["points0"], nb)
    assert np.allclose(nrm, G["normals0"], atol=1e-12) and np.allclose(cv, G["covs0"], atol=1e-12)
    vg, vt, _ = oracle.voxelgrid_sampling(G["points0"], 0.3, times=G["times0"])
    assert np.array_equal(vg, G["voxelgrid_points"]) and np.array_equal(vt, G["voxelgrid_times"])


@pytest.mark.gpu
def test_gpu_reproduces_golden(ctx):
    from glim_b200 import gpu, preprocess

    tgt = gpu.PointCloudGPU.clone(G["points0"], G["covs0"], ctx=ctx)
    src = gpu.PointCloudGPU.clone(G["points1"], G["covs1"], ctx=ctx)
    for res in (0.25, 0.5):
        tag = f"r{int(res * 100):03d}"
        m = gpu.GaussianVoxelMapGPU(res, ctx=ctx).insert(tgt)
        buckets, vnum, vmean, vcov = m.download()
        assert np.array_equal(buckets, G[f"{tag}_buckets"]) and np.array_equal(vnum, G[f"{tag}_vnum"])
        assert np.array_equal(vmean, G[f"{tag}_vmean"]) and np.array_equal(vcov, G[f"{tag}_vcov"])
        fac = gpu.IntegratedVGICPFactorGPU(np.eye(4), 1, m, src, ctx=ctx)
        for k, T in enumerate(G["poses"]):
            got = fac.linearize({1: T})
            ref = oracle.split122(G[f"{tag}_linearized"][k])
            assert got["num_inliers"] == ref["num_inliers"] == (G[f"{tag}_corr"][k] >= 0).sum()
            for key in ("H_tt", "H_ss", "H_ts"):
                assert util.rel_err(got[key], ref[key]) < util.REL_TOL
            assert abs(got["error"] - ref["error"]) <= util.REL_TOL * ref["error"] + 1e-12
            assert gpu.overlap_gpu(m, src, T) == G[f"{tag}_overlap"][k]
        fac.linearize({1: G["poses"][0]})
        for k, T in enumerate(G["poses"]):
            assert fac.error({1: T}) == pytest.approx(G[f"{tag}_error"][k], rel=util.REL_TOL)
    nb = preprocess.find_neighbors(G["points0"], 10, ctx=ctx).reshape(-1, 10)
    assert np.array_equal(nb, G["knn0"])
    nrm, cv = preprocess.CloudCovarianceEstimation(ctx=ctx).estimate(G["points0"], nb)
    assert np.allclose(cv, G["covs0"], atol=1e-9)
    vg, vt, _ = preprocess.voxelgrid_sampling(G["points0"], 0.3, times=G["times0"], ctx=ctx)
    assert np.array_equal(vg, G["voxelgrid_points"]) and np.array_equal(vt, G["voxelgrid_times"])


def test_oracle_reproduces_round2_golden():
    """random grid, merge_frames and the surface-validation gate (tests/golden/make_golden_round2.py): exact."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden_round2", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_round2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "round2_golden.npz"))
    now = mod.compute(G)
    assert set(now) == set(G2.files)
    for k in G2.files:
        if k in ("sv_linearized", "normals1"):
            assert np.allclose(now[k], G2[k], rtol=1e-12, atol=1e-9), k
        else:
            assert np.array_equal(now[k], G2[k]), k
    # sanity of what was frozen: the gate rejects some but not all correspondences; thinning halves the merged cloud
    assert 0 < (G2["sv_corr"] == -2).sum() < (G2["sv_corr"] != -1).sum()
    assert len(G2["merge_thinned_points"]) == len(G2["merge_points"]) // 2
    assert 0 < len(G2["randomgrid_keep_r100_rate030_seed5"]) <= int(len(G["points0"]) * 0.3 * 1.2)
