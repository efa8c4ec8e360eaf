import sys, numpy as np
sys.path.insert(0,'.')
from glim_b200 import gpu, synth
from oracle import oracle
from tests import util
ctx=gpu.Context(0)
sc = synth.make_hall_scene()
traj = synth.arc_trajectory(8, step=1.5)
clouds, packed, poses = [], [], []
origin = traj[2]
for i in range(5):
    pts, _ = synth.scan(sc, "hdl32", traj[i], synth.rng_for(61, i), n_rays=32 * 250)
    _, cov = synth.with_covariances(pts, 10)
    clouds.append(gpu.PointCloudGPU.clone(pts, cov, ctx=ctx))
    packed.append(oracle.pack_cloud(pts, util.cov_colmajor16(cov)))
    poses.append(synth.inv_pose(origin) @ traj[i])
for K in (1,2,5):
    pts, covs, merged = gpu.merge_frames_gpu(poses[:K], clouds[:K], 0.25, 0, seed=9, ctx=ctx)
    rp, rc = oracle.merge_frames(poses[:K], packed[:K], 0.25, 0, seed=9)
    d=np.abs(pts-rp); 
    print(K, len(pts), len(rp), 'max diff', d.max(), 'n bad rows', (d.max(1)>0).sum(), 'cov max', np.abs(covs-rc).max())
    bad=np.nonzero(d.max(1)>0)[0][:3]
    for b in bad: print(b, pts[b], rp[b])
for target in (0, 3000):
    pts, covs, merged = gpu.merge_frames_gpu(poses, clouds, 0.25, target, seed=9, ctx=ctx)
    rp, rc = oracle.merge_frames(poses, packed, 0.25, target, seed=9)
    print('target', target, len(pts), len(rp), np.array_equal(pts, rp) if len(pts)==len(rp) else None, np.array_equal(covs, rc) if len(pts)==len(rp) else None)
    gx, gc = merged.download()
    xyz, cov6 = oracle.pack_cloud(pts, util.cov_colmajor16(covs))
    print('  cloud', np.array_equal(gx, xyz), np.array_equal(gc, cov6), np.abs(gc-cov6).max())
    print('  sym', np.allclose(covs, covs.transpose(0,2,1)), covs[:,3,:].any(), np.linalg.eigvalsh(covs[:, :3, :3]).min())
