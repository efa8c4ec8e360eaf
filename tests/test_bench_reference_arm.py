"""The driver's reference arm (`bench.py --impl reference`) runs on the host cores only: it is exercised here, on the CPU-only
box, at a debug scale.  Contract (task statement, measurement section): same metric / unit / config keys as our arm, `impl`,
`cpu_baseline` {kind, cores, sample}, `e2e` with zero copy bytes; under torchrun only rank 0 works and prints; the thread count
comes from the affinity mask, not from OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 -- round 1's arm ran on one thread)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra, *flags):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.1", "--steps", "2", "--warmup", "1", "--cpu-seconds", "1", *flags],
                          capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_line_and_thread_count_under_torchrun_environment(tmp_path):
    # LD_DEBUG=files: the dynamic loader logs every object it maps, dlopen()ed ones included
    r = run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "LD_DEBUG": "files", "LD_DEBUG_OUTPUT": str(tmp_path / "ld")}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "VGICP linearize throughput" and d["unit"] == "M points*factors/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 2 and d["value"] > 0
    assert d["config"]["workload"] == "global_mapping_gpu"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["sample"]
    avail = len(os.sched_getaffinity(0))
    assert cb["host_threads_available"] == avail and cb["omp_num_threads_env"] == "1"
    if avail >= 2:
        assert cb["cores"] >= 2, "the arm must not inherit OMP_NUM_THREADS=1"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # nothing of the product library may be loaded by this arm; the oracle (the CPU path being timed) must be
    loaded = "".join(open(tmp_path / f, errors="replace").read() for f in os.listdir(tmp_path))
    assert "libglim_oracle.so" in loaded, "loader log is empty: the check below would be vacuous"
    assert "libglim_b200.so" not in loaded


def test_reference_arm_other_ranks_exit_without_work():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
