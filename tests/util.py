"""Shared helpers for the test-suite: small seeded scenarios and error metrics."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from glim_b200 import synth  # noqa: E402

# tolerance of north_star: Hessians / gradients within 1e-4 relative (Frobenius) of the fp64 oracle
REL_TOL = 1e-4


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return d / n if n > 0 else d


@functools.lru_cache(maxsize=None)
def scan_pair(sensor="hdl32", n_rays=32 * 300, step=1.0, key=0, scene="hall"):
    """Two scans `step` metres apart along the M2 arc, with PLANE covariances (k = 10).
    -> dict(points=[P0,P1] (N,4), covs=[C0,C1] (N,4,4), normals=[..], poses=[T0,T1])"""
    sc = synth.make_hall_scene() if scene == "hall" else synth.make_blocks_scene()
    traj = synth.arc_trajectory(8, step=step)
    out = {"points": [], "covs": [], "normals": [], "poses": [], "times": []}
    for i in (3, 4):
        pts, tms = synth.scan(sc, sensor, traj[i], synth.rng_for(77, key, i), n_rays=n_rays)
        nrm, cov = synth.with_covariances(pts, 10)
        out["points"].append(pts)
        out["covs"].append(cov)
        out["normals"].append(nrm)
        out["poses"].append(traj[i])
        out["times"].append(tms)
    return out


def test_poses(T_gt, n=4, key=0):
    """Ground-truth delta and a handful of perturbed / adversarial deltas."""
    rng = synth.rng_for(99, key)
    poses = [T_gt]
    for _ in range(n):
        poses.append(synth.perturb(T_gt, rng, 0.01, 0.05))
    return poses


test_poses.__test__ = False


def cov_colmajor16(covs):
    n = covs.shape[0]
    return np.ascontiguousarray(np.swapaxes(covs.reshape(n, 4, 4), 1, 2)).reshape(n, 16)
