"""Generates tests/golden/round2_golden.npz: frozen outputs of the oracle functions added in round 2 (random-grid
downsampling, merge_frames, the surface-validation gate) on the inputs of vgicp_golden.npz.  As for make_golden.py: the
reference has no vectors for these (their rules live in the un-vendored gtsam_points), so the file pins
OUR oracle against drift; the CUDA path is compared with the same oracle functions live in tests/test_gpu_parity.py.

    python tests/golden/make_golden_round2.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def c16(c):
    return np.ascontiguousarray(np.swapaxes(c, 1, 2)).reshape(len(c), 16)


def compute(G):
    out = {}
    out["randomgrid_keep_r100_rate030_seed5"] = np.flatnonzero(oracle.randomgrid_sampling(G["points0"], 1.0, 0.3, seed=5)).astype(np.int32)
    out["randomgrid_keep_r050_rate010_seed0"] = np.flatnonzero(oracle.randomgrid_sampling(G["points0"], 0.5, 0.1, seed=0)).astype(np.int32)
    clouds = [oracle.pack_cloud(G["points0"], c16(G["covs0"])), oracle.pack_cloud(G["points1"], c16(G["covs1"]))]
    poses = [np.eye(4), G["poses"][0]]
    p, c = oracle.merge_frames(poses, clouds, 0.5)
    out["merge_points"], out["merge_covs"] = p, c
    pt, ct = oracle.merge_frames(poses, clouds, 0.5, target=len(p) // 2, seed=7)
    out["merge_thinned_points"], out["merge_thinned_covs"] = pt, ct
    nb, _ = oracle.knn_bruteforce(G["points1"], 10)
    nrm1, _ = oracle.covariance_estimate(G["points1"], nb)
    out["normals1"] = nrm1
    m = oracle.GpuMap(clouds[0][0], clouds[0][1], 0.5)
    lin, corr = [], []
    for k in (0, 4):
        o, cc = oracle.linearize_gpumap(m, clouds[1][0], clouds[1][1], G["poses"][k], normals=nrm1)
        lin.append(o)
        corr.append(cc)
    out["sv_linearized"], out["sv_corr"] = np.stack(lin), np.stack(corr)
    return out


def main():
    G = np.load(os.path.join(HERE, "vgicp_golden.npz"))
    out = compute(G)
    path = os.path.join(HERE, "round2_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
