#!/usr/bin/env python
"""bench.py -- VGICP linearize throughput (M points.factors / s) on B200, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one relinearization pass of the hot path over the workload's whole factor set (every
NonlinearFactorSetGPU::linearize call the workload contains, back to back).  Default workload:
global_mapping_gpu (BASELINE.json configs[3]: 256 submaps x 50 k points, 0.5 / 1.0 m voxels), the one
configuration BASELINE.json defines at 1 / 2 / 4 / 8 GPUs -- total work is fixed, the pair list is sharded
over the ranks and the per-pair Hessian slab is all-reduced over NCCL ("scaling": "strong").  The other
BASELINE configurations (single_pair, odometry_gpu, sub_mapping_gpu, livox_stress) run with --workload.

value      whole-job points.factors / s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
e2e        the same through the C-ABI with HOST buffers: poses H2D from pinned memory + launch (+ all-reduce) +
           gb_linearized6 records (and the slab) D2H, every step
roofline   algorithmic bytes per launch (SURVEY 8(d) B_sweep, from gb_sweep_stats) / measured launch duration vs
           MEASURED_PEAKS.json hbm_gbs
--impl reference   the CPU path (oracle port of gtsam_points::IntegratedVGICPFactor, fp64, all host threads) on a
           bounded sample of the same workload.  The only place besides cpu_baseline where bench.py executes oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "VGICP linearize throughput"
UNIT = "M points*factors/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="global_mapping_gpu", choices=["single_pair", "odometry_gpu", "sub_mapping_gpu", "global_mapping_gpu", "livox_stress"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; a scaled run is not a bench value)")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"], help="multi-GPU result exchange: rows pushed into peer memory by the sweep kernel (fused) or fp32 atomics + NCCL all-reduce (nccl)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget for the cpu_baseline sample")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
# workload construction (identical on every rank: seeded)
# --------------------------------------------------------------------------------------------------------------
def build_workload(name, ctx, scale, use_gpu):
    from glim_b200 import workloads

    s = scale
    if name == "global_mapping_gpu":
        n = max(8, int(round(256 * s)) // 4 * 4)
        return workloads.global_mapping(ctx, n_submaps=n, laps=4, side=300.0, use_gpu=use_gpu, n_rays=None if s >= 1 else 64 * max(64, int(2048 * s)),
                                        params=workloads.GlobalMappingParams(submap_target_num_points=max(2000, int(50000 * s))))
    if name == "odometry_gpu":
        n = max(6, int(round(64 * s)))
        return workloads.odometry_stream(ctx, n_frames=n, first_bench_frame=max(1, n // 4), use_gpu=use_gpu, n_rays=None if s >= 1 else 32 * max(64, int(1875 * s)))
    if name == "sub_mapping_gpu":
        return workloads.sub_mapping_bundle(ctx, n_keyframes=max(3, int(round(15 * s))), use_gpu=use_gpu, n_rays=None if s >= 1 else 64 * max(64, int(2048 * s)))
    if name == "single_pair":
        return workloads.single_pair(ctx, use_gpu=use_gpu, n_rays=None if s >= 1 else 64 * max(64, int(1563 * s)))
    if name == "livox_stress":
        return workloads.livox_stress(ctx, n_rays=max(5000, int(500_000 * s)), use_gpu=use_gpu)
    raise ValueError(name)


def workload_config(name, w, extra):
    cfg = {
        "workload": name,
        "clouds": len(w.host_clouds),
        "points_per_cloud": int(np.median([len(c[0]) for c in w.host_clouds])),
        "voxel_resolutions_m": w.resolutions,
        "factor_sets_per_step": len(w.sets),
        "factors_per_step": int(sum(len(s.factors) for s in w.sets)),
        "point_factors_per_step": int(w.point_factors),
        "l2": "inputs larger than L2 (no flush)" if sum(len(c[0]) for c in w.host_clouds) * 36 > 126e6 else "L2 flushed between steps (256 MB write)",
    }
    cfg.update({k: v for k, v in w.notes.items() if not k.startswith("_")})
    cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.monotonic(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi not available"}
        sel = [l for (t, l) in self.samples if t0 <= t <= t1] or [l for (_, l) in self.samples[-3:]]
        mhz, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in sel:
            f = [x.strip() for x in l.split(",")]
            try:
                mhz.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(mhz)}


# --------------------------------------------------------------------------------------------------------------
# CPU path (oracle port of gtsam_points::IntegratedVGICPFactor), used by cpu_baseline and --impl reference
# --------------------------------------------------------------------------------------------------------------
class CpuSample:
    """A bounded sample of the workload for the CPU path: the factors of the factor set's most connected source
    cloud (one 'insert_submap' worth of factors in global mapping, one frame's factors in odometry)."""

    def __init__(self, w, max_factors=16):
        from oracle import oracle

        self.oracle = oracle
        fset = w.sets[len(w.sets) // 2]
        by_src = {}
        for k, f in enumerate(fset.factors):
            by_src.setdefault(f.source, []).append(k)
        src = max(by_src, key=lambda s: (len(by_src[s]), -s))
        idx = by_src[src][:max_factors]
        self.threads = oracle.num_threads()
        self.items = []
        maps = {}
        pts, cov = w.host_clouds[src]
        c16 = np.ascontiguousarray(np.swapaxes(cov, 1, 2)).reshape(len(cov), 16)
        for k in idx:
            f = fset.factors[k]
            key = (f.target, f.level)
            if key not in maps:
                m = oracle.CpuMap(w.resolutions[f.level])
                tp, tc = w.host_clouds[f.target]
                m.insert(tp, np.ascontiguousarray(np.swapaxes(tc, 1, 2)).reshape(len(tc), 16))
                maps[key] = m
            fac = oracle.CpuFactor(maps[key], pts, c16, num_threads=self.threads)
            self.items.append((fac, oracle.pose_colmajor(fset.deltas[k])))
        self.point_factors = len(pts) * len(self.items)
        self.out = np.zeros(122)
        self.desc = f"{len(self.items)} factors x {len(pts)} source points of one {w.name} source cloud (all of its factors, capped at {max_factors})"

    def set_threads(self, t):
        for fac, _ in self.items:
            fac.num_threads = t

    def run_once(self):
        for fac, Tc in self.items:
            fac.linearize_raw(Tc, self.out)


def cpu_thread_sweep(s, seconds):
    """Time the sample at the shipped thread count (2, config_odometry_cpu.json:36) and at 8 / 16 / 32 / all host
    threads; the baseline reported is the FASTEST of them (OpenMP fork/join can make 'all threads' slower on small
    clouds -- the strongest CPU number is the honest one to compare against)."""
    allt = s.threads
    cands = sorted({t for t in (2, 8, 16, 32, allt) if t <= allt})
    per = {}
    budget = seconds / len(cands)
    for t in cands:
        s.set_threads(t)
        s.run_once()
        t0 = time.perf_counter()
        reps = 0
        while True:
            s.run_once()
            reps += 1
            el = time.perf_counter() - t0
            if el >= budget or reps >= 100000:
                break
        per[t] = s.point_factors * reps / el / 1e6
    best = max(per, key=lambda t: per[t])
    s.set_threads(best)
    return best, per


def cpu_baseline(w, seconds):
    s = CpuSample(w)
    best, per = cpu_thread_sweep(s, seconds)
    return {"value": per[best], "unit": UNIT, "cores": best, "kind": "port", "host_threads_available": s.threads,
            "by_threads": {str(t): round(v, 2) for t, v in per.items()},
            "sample": f"{s.desc}; ~{seconds:.0f} s of CPU work split over thread counts {sorted(per)}; fp64, update_correspondences + evaluate (OpenMP); value = fastest thread count ({best})"}


def run_reference(args, rank):
    """--impl reference: the CPU VGICP path on the box's host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    # The sample needs the workload's clouds and factor lists.  With a GPU present the synthetic inputs are generated
    # there (torch ray casting, overlap gate) exactly as for our arm; without one the same construction runs on the host
    # in numpy (slow at full scale).  Nothing of libglim_b200 is inside the timed region either way.
    import torch

    ctx = None
    if torch.cuda.is_available():
        from glim_b200 import gpu

        torch.cuda.set_device(0)
        ctx = gpu.Context(0)
    w = build_workload(args.workload, ctx, args.scale, use_gpu=ctx is not None)
    s = CpuSample(w)
    best, per = cpu_thread_sweep(s, 6.0)  # pick the fastest thread count for this host, then time K steps with it
    for _ in range(max(1, min(args.warmup, 3))):
        s.run_once()
    # bound the run: at most ~60 s of CPU work
    t_probe = time.perf_counter()
    s.run_once()
    per_step = time.perf_counter() - t_probe
    steps = max(1, min(args.steps, int(60.0 / max(per_step, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        s.run_once()
    el = time.perf_counter() - t0
    val = s.point_factors * steps / el / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, w, {"reference_step": s.desc}),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": best, "kind": "port", "host_threads_available": s.threads, "by_threads": {str(t): round(v, 2) for t, v in per.items()},
                         "sample": s.desc + f"; one step = one pass over the sample with the fastest thread count ({best})"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "GLIM's own CPU path (gtsam_points::IntegratedVGICPFactor) cannot be built here (GTSAM / gtsam_points / Eigen absent); this is the oracle port, fp64, all host threads",
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# main (our arm)
# --------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from glim_b200 import gpu, multi_gpu
    from glim_b200.capi import GB_SLAB_STRIDE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = gpu.Context(local_rank, cuda_stream=stream.cuda_stream)

    sampler = ClockSampler(local_rank) if rank == 0 else None  # started early: nvidia-smi takes ~1 s to produce its first line
    t_build = time.perf_counter()
    w = build_workload(args.workload, ctx, args.scale, use_gpu=True)
    build_s = time.perf_counter() - t_build

    # ---- shard: pairs over ranks for global mapping; every other workload is a single online stream (replicas only) ----
    sharded = args.workload == "global_mapping_gpu"
    sizes = [len(c[0]) for c in w.host_clouds]
    fused = args.collective == "fused"

    def exchange(handle: bytes):
        t = torch.tensor(list(handle), dtype=torch.uint8, device=f"cuda:{local_rank}")
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [bytes(o.cpu().numpy().tobytes()) for o in outs]

    def make_sweeps(inliers_by_set=None, reuse_peers=None):
        """Partition every factor set over the ranks and prepare this rank's sweeps.  inliers_by_set: measured inlier counts per
        factor (from a calibration sweep) for the cost-balanced partition; None -> the overlap estimate of the gate."""
        sweeps, slabs, peers, my_pf, my_bytes, all_pf = [], [], [], 0, 0, 0
        for si, fset in enumerate(w.sets):
            if sharded:
                f_rank, _ = multi_gpu.shard_factors(fset.factors, sizes, world, pair_cost=w.notes.get("_pair_overlap"),
                                                    factor_inliers=None if inliers_by_set is None else inliers_by_set[si])
                mine = [k for k in range(len(fset.factors)) if f_rank[k] == rank]
            else:
                mine = list(range(len(fset.factors)))
            sub = type(fset)([fset.factors[k] for k in mine], fset.deltas[mine] if len(mine) else np.zeros((0, 4, 4)))
            gf = w.gpu_factors(sub)
            num_pairs = max((f.pair for f in fset.factors), default=-1) + 1
            sw = gpu.Sweep(ctx, gf, pair_index=[f.pair for f in sub.factors])
            slab = torch.zeros((max(1, num_pairs), GB_SLAB_STRIDE), dtype=torch.float32, device=f"cuda:{local_rank}")
            if fused:
                ps = reuse_peers[si] if reuse_peers is not None else gpu.PeerSlab(ctx, max(1, num_pairs), world if sharded else 1, rank if sharded else 0, exchange)
                sw.attach_peer_slab(ps)
                peers.append(ps)
            else:
                sw.attach_slab(slab.data_ptr(), max(1, num_pairs))
                peers.append(None)
            sw.set_poses(sub.deltas)
            sw._sub = sub
            sw._mine = mine
            sweeps.append(sw)
            slabs.append(slab)
            my_pf += sw.point_factors
            my_bytes += sw.algorithmic_bytes
            all_pf += sum(sizes[f.source] for f in fset.factors)
        return sweeps, slabs, peers, my_pf, my_bytes, all_pf

    sweeps, slabs, peers, my_pf, my_bytes, all_pf = make_sweeps()
    calibrated = False
    if sharded and world > 1:
        # Calibration sweep: a relinearizing back-end knows every factor's inlier count from its previous sweep; use it to
        # balance the partition by measured cost instead of the gate's overlap estimate, then rebuild this rank's sweeps.
        inl = []
        for si, (sw, ps) in enumerate(zip(sweeps, peers)):
            sw.launch()
            if fused:
                ps.signal_wait()
            rec = sw.fetch()
            full = torch.zeros(len(w.sets[si].factors), dtype=torch.float64, device=f"cuda:{local_rank}")
            if len(sw._mine):
                full[torch.tensor(sw._mine, device=full.device)] = torch.from_numpy(np.ascontiguousarray(rec["num_inliers"])).to(full.device)
            dist.all_reduce(full, op=dist.ReduceOp.SUM)
            inl.append(full.cpu().numpy())
        old = sweeps
        sweeps, slabs, peers, my_pf, my_bytes, all_pf = make_sweeps(inl, reuse_peers=peers if fused else None)
        del old
        calibrated = True
    total_pf = all_pf if sharded else all_pf * world  # replicas: every rank processes the whole stream
    ctx.synchronize()

    small_inputs = sum(sizes) * 36 <= 126e6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}") if small_inputs else None

    def step_device():
        for sw, slab, ps in zip(sweeps, slabs, peers):
            if fused:
                sw.launch()  # epilogue stores finished pair rows into every rank's slab
                ps.signal_wait()  # completion flags: publish ours, wait for the peers'
            else:
                slab.zero_()
                sw.launch()
                if world > 1 and sharded:
                    dist.all_reduce(slab, op=dist.ReduceOp.SUM)

    def timed(fn, steps, flush_l2):
        """K steps bracketed by barrier + synchronize; device time from CUDA events on the launching stream; max over ranks."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] if flush_l2 else None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.monotonic()
        if flush_l2:
            for a, b in ev:
                flush.fill_(1)
                a.record(stream)
                fn()
                b.record(stream)
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in ev)
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(steps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t1 = time.monotonic()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1

    for _ in range(max(3, args.warmup)):
        step_device()
    launches0 = ctx.kernel_launches
    profiling = bool(os.environ.get("GB_PROFILE"))  # ncu --profile-from-start off: capture only the timed region
    if profiling:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ms, t0, t1 = timed(step_device, args.steps, small_inputs)
    if profiling:
        torch.cuda.profiler.stop()
    gpu_launches = (ctx.kernel_launches - launches0) * world  # every rank launches the same sequence
    ms_per_step = ms / args.steps
    value = total_pf / (ms_per_step * 1e-3) / 1e6

    # ---- kernel-only duration for the roofline (same launches, no memset / collective in the bracket) ----
    def step_kernels():
        for sw in sweeps:
            sw.launch()

    for s in slabs:
        s.zero_()
    kms, _, _ = timed(step_kernels, args.steps, small_inputs)
    k_launches = len(sweeps)
    kernel_ms_per_launch = kms / args.steps / max(1, k_launches)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = my_bytes / max(1, k_launches) / (kernel_ms_per_launch * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this workload (1 GPU, full scale only)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if world == 1 and args.scale == 1.0 and args.workload in tj:
            traffic, traffic_src = tj[args.workload]["dram_bytes_per_launch"], tj[args.workload]["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_vgicp_sweep<LINEARIZE>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": my_bytes / max(1, k_launches), "launch_ms": kernel_ms_per_launch,
                "traffic": traffic, "traffic_source": traffic_src, "note": "achieved = SURVEY 8(d) B_sweep (48 B per point.factor + 48 B per target voxel + 16 B per bucket + 552 B per factor) / CUDA-event launch time; source clouds shared by consecutive factors are re-read from L2, so DRAM traffic is below B_sweep (see profiles/)"}

    # ---- end to end through the C-ABI with host buffers ----
    host_deltas = [sw._sub.deltas.copy() for sw in sweeps]
    h2d = sum(d.shape[0] * 128 for d in host_deltas)
    d2h = sum(d.shape[0] * 976 for d in host_deltas) + (sum(s.numel() * 4 for s in slabs) if sharded else 0)
    host_slabs = [torch.empty(s.shape, dtype=torch.float32, pin_memory=True) for s in slabs] if sharded else None

    def step_e2e():
        for k, (sw, slab, ps) in enumerate(zip(sweeps, slabs, peers)):
            sw.set_poses(host_deltas[k])  # gb_sweep_set_poses: pinned staging + H2D
            if fused:
                sw.launch()
                ps.signal_wait()
                if sharded:
                    ps.fetch_async()  # D2H of the complete slab into pinned memory; the sync below covers it
            else:
                slab.zero_()
                sw.launch()
                if world > 1 and sharded:
                    dist.all_reduce(slab, op=dist.ReduceOp.SUM)
                if sharded:
                    host_slabs[k].copy_(slab, non_blocking=True)
            sw.fetch()  # gb_sweep_fetch: D2H of the gb_linearized6 records + stream sync

    for _ in range(3):
        step_e2e()
    e_steps = max(3, min(args.steps, 50))
    ems, _, _ = timed(step_e2e, e_steps, False)
    e2e_val = total_pf / (ems / e_steps * 1e-3) / 1e6

    # checksum of the (all-reduced) Hessian slab of one step: identical at every N up to fp32 reduction order
    step_device()
    torch.cuda.synchronize()
    slab_checksum = float(sum(np.abs(ps.fetch().astype(np.float64)).sum() for ps in peers)) if fused else float(sum(s.double().abs().sum().item() for s in slabs))
    if sampler:
        time.sleep(0.15)
        sampler.stop()
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, w, {"parallelism": ((f"pairs sharded over {world} rank(s); finished pair rows of the [{slabs[0].shape[0]} x {GB_SLAB_STRIDE}] fp32 Hessian slab are stored into every rank's buffer by the sweep kernel's epilogue over NVLink (CUDA IPC peer memory) + completion flags; no NCCL in the step" if fused else f"pairs sharded over {world} rank(s), NCCL all-reduce of the [{slabs[0].shape[0]} x {GB_SLAB_STRIDE}] fp32 Hessian slab") if sharded else f"replicas x{world} (single online stream does not shard)"),
                                                       "tiles_grid_first_sweep": [int(sweeps[0].num_tiles), int(sweeps[0].grid)], "partition": ("contiguous, cost = n_source + 1.25 * measured inliers (calibration sweep)" if calibrated else "contiguous, cost = n_source * (1 + 1.25 * gate overlap)") if sharded else None, "build_seconds": round(build_s, 1), "scale": args.scale}),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ems / e_steps},
            "gpu_launches": int(gpu_launches),
            "clocks": sampler.summary(t0, t1) if sampler else None,
            "roofline": roofline,
            "slab_checksum": slab_checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(w, args.cpu_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
