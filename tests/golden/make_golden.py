"""Generates tests/golden/vgicp_golden.npz: frozen inputs + outputs of the CPU oracle for the hot path.

The reference (GLIM v1.2.2) has no golden vectors and its arithmetic (gtsam_points) cannot be built or imported here, so
these vectors pin OUR oracle (oracle/glim_oracle.c) against regressions and travel to the GPU box, where the CUDA path is
checked against them (tests/test_golden.py).  Re-run only when the oracle's definition changes deliberately:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from glim_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    sc = synth.make_hall_scene()
    traj = synth.arc_trajectory(8)
    pts, covs, times = [], [], []
    for i in (3, 4):
        p, t = synth.scan(sc, "hdl32", traj[i], synth.rng_for(900, i), n_rays=32 * 60)
        nb, _ = oracle.knn_bruteforce(p, 10)
        _, c = oracle.covariance_estimate(p, nb)
        pts.append(p)
        covs.append(c)
        times.append(t)
    T_gt = synth.inv_pose(traj[3]) @ traj[4]
    rng = synth.rng_for(901)
    poses = np.stack([T_gt] + [synth.perturb(T_gt, rng, 0.01, 0.05) for _ in range(3)] + [synth.pose(0.3, -0.2, 0.05, np.pi, 0.0, 0.0)])
    out = {"points0": pts[0], "points1": pts[1], "covs0": covs[0], "covs1": covs[1], "times0": times[0], "poses": poses}

    def c16(c):
        return np.ascontiguousarray(np.swapaxes(c, 1, 2)).reshape(len(c), 16)

    xyz0, cov0 = oracle.pack_cloud(pts[0], c16(covs[0]))
    xyz1, cov1 = oracle.pack_cloud(pts[1], c16(covs[1]))
    for res in (0.25, 0.5):
        m = oracle.GpuMap(xyz0, cov0, res)
        tag = f"r{int(res * 100):03d}"
        out[f"{tag}_buckets"] = m.buckets
        out[f"{tag}_vnum"] = m.vnum
        out[f"{tag}_vmean"] = m.vmean
        out[f"{tag}_vcov"] = m.vcov
        lin, corr = [], []
        for T in poses:
            o, c = oracle.linearize_gpumap(m, xyz1, cov1, T)
            lin.append(o)
            corr.append(c)
        out[f"{tag}_linearized"] = np.stack(lin)
        out[f"{tag}_corr"] = np.stack(corr)
        out[f"{tag}_error"] = np.array([oracle.error_gpumap(m, xyz1, cov1, poses[0], T) for T in poses])
        out[f"{tag}_overlap"] = np.array([oracle.overlap_gpumap([m], xyz1, [T]) for T in poses])
    # CPU (fp64, unordered_map) factor
    cm = oracle.CpuMap(0.5)
    cm.insert(pts[0], c16(covs[0]))
    fac = oracle.CpuFactor(cm, pts[1], c16(covs[1]))
    out["cpu_linearized"] = np.stack([fac.linearize(T) for T in poses])
    # preprocess
    nb, _ = oracle.knn_bruteforce(pts[0], 10)
    nrm, cv = oracle.covariance_estimate(pts[0], nb)
    out["knn0"] = nb
    out["normals0"] = nrm
    vg, vt, _ = oracle.voxelgrid_sampling(pts[0], 0.3, times=times[0])
    out["voxelgrid_points"] = vg
    out["voxelgrid_times"] = vt
    path = os.path.join(ROOT, "tests", "golden", "vgicp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(pts[0]), len(pts[1]), "points")


if __name__ == "__main__":
    main()
