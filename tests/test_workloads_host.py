"""Host logic of the BASELINE workloads (no GPU: the overlap gate falls back to its numpy twin): the factor lists follow the
reference's factor-creation rules."""
import numpy as np

from glim_b200 import workloads


def test_odometry_factor_rule():
    """OdometryEstimationGPU::create_factors (odometry_estimation_gpu.cpp:128-206): one factor per voxel level to each of the
    last `full_connection_window_size` frames and to every keyframe outside that window; levels of a target share a pair id."""
    p = workloads.OdometryParams(max_num_keyframes=4)
    w = workloads.odometry_stream(None, n_frames=10, first_bench_frame=1, n_rays=32 * 40, params=p, step=4.0)
    assert len(w.sets) == 9 and w.resolutions == [0.25, 0.5]
    kf = w.notes["keyframes_final"]
    assert kf[0] >= 0 and len(kf) <= p.max_num_keyframes and kf == sorted(kf)
    for k, s in enumerate(w.sets):
        cur = k + 1
        assert all(f.source == cur for f in s.factors)
        targets = [f.target for f in s.factors]
        assert len(s.factors) % 2 == 0 and len(s.deltas) == len(s.factors)
        for t in range(max(0, cur - p.full_connection_window_size), cur):
            assert targets.count(t) == 2  # both levels of every window frame
        for a, b in zip(s.factors[0::2], s.factors[1::2]):
            assert (a.target, a.pair, a.level, b.level) == (b.target, b.pair, 0, 1)
            assert np.array_equal(s.deltas[s.factors.index(a)], s.deltas[s.factors.index(b)])  # levels share the pose
        assert max(targets) < cur


def test_sub_mapping_bundle_is_fully_connected():
    """SubMapping::insert_frame (sub_mapping.cpp:276-310): every earlier keyframe x every level."""
    w = workloads.sub_mapping_bundle(None, n_keyframes=5, n_rays=64 * 32)
    f = w.sets[0].factors
    assert len(f) == 5 * 4 // 2 * 2 and len({x.pair for x in f}) == 10
    assert all(x.target < x.source for x in f)


def test_global_mapping_gate_and_order():
    """GlobalMapping::create_matching_cost_factors (global_mapping.cpp:441-470): source-major order, targets ascending within a
    source, only pairs whose overlap passes the gate, one factor per level."""
    w = workloads.global_mapping(None, n_submaps=16, laps=4, n_rays=64 * 64, params=workloads.GlobalMappingParams(submap_target_num_points=1500), side=60.0)
    f = w.sets[0].factors
    assert len(f) > 0 and len(f) % 2 == 0
    src = [x.source for x in f]
    assert src == sorted(src)
    ov = w.notes["_pair_overlap"]
    assert len(ov) == w.notes["num_pairs"] and min(ov.values()) >= 0.2
    for a, b in zip(f[0::2], f[1::2]):
        assert (a.target, a.source, a.pair) == (b.target, b.source, b.pair) and a.target < a.source


def test_host_overlap_twin_equals_the_oracle():
    """Workload.overlap without a context (the numpy twin that lets the reference arm and these tests run without the product
    library) must be the oracle's overlap_gpu: fraction of transformed source points whose fp32 voxel is occupied in any target."""
    from glim_b200 import synth
    from oracle import oracle
    from tests import util

    w = workloads.sub_mapping_bundle(None, n_keyframes=4, n_rays=64 * 48)
    packed = [oracle.pack_cloud(c[0], util.cov_colmajor16(c[1])) for c in w.host_clouds]
    for level, res in enumerate(w.resolutions):
        maps = [oracle.GpuMap(*packed[t], res) for t in range(3)]
        for targets in ([0], [1, 2], [0, 1, 2]):
            deltas = [synth.perturb(w.gt_delta(t, 3), synth.rng_for(7, t), 0.01, 0.05) for t in targets]
            got = w.overlap(targets, level, 3, deltas)
            # the twin transforms in fp64 and casts to fp32; the oracle (like the kernel) transforms in fp32 with explicit FMAs: a point
            # on a voxel face may fall on the other side -- a handful of points out of thousands
            want = oracle.overlap_gpumap([maps[t] for t in targets], packed[3][0], deltas)
            assert abs(got - want) <= 5.0 / len(packed[3][0]), (targets, level, got, want)
            assert 0.0 < want <= 1.0
