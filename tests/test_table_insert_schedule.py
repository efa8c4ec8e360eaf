"""k_table_insert (glim_b200/csrc/gb_kernels_voxelmap.cu:95-116) inserts all voxels CONCURRENTLY with atomicMin-priority linear
probing and claims that its fixed point is exactly the table a sequential first-free-slot insertion in ascending voxel id builds
(what the oracle's go_gpumap_build does) -- including which voxels run out of probes -- whatever the thread schedule.  The GPU
parity tests sample one schedule (the hardware's); here the kernel's per-thread loop is restated as a step function and driven by
ADVERSARIAL RANDOM SCHEDULES (one atomic step of one randomly chosen thread at a time) on small, heavily loaded tables, and the
final table and the dropped set are compared with the sequential build.  Pure host logic, runs on the CPU-only box."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

EMPTY = 0x7FFFFFFF


def sequential(homes, mask, max_scan):
    """ascending voxel id, first free slot within max_scan probes (oracle/glim_oracle.c go_gpumap_build)"""
    table = np.full(mask + 1, EMPTY, np.int64)
    dropped = []
    for v, h in enumerate(homes):
        for i in range(max_scan):
            s = (h + i) & mask
            if table[s] == EMPTY:
                table[s] = v
                break
        else:
            dropped.append(v)
    return table, sorted(dropped)


class Thread:
    """one thread of k_table_insert: state (cur, s, dist); step() = one iteration of its for (;;) loop (one atomicMin)"""

    def __init__(self, v, homes, mask):
        self.cur, self.s, self.dist, self.done = v, homes[v] & mask, 0, False

    def step(self, table, homes, mask, max_scan, dropped):
        if self.dist >= max_scan:
            dropped.append(self.cur)
            self.done = True
            return
        old = table[self.s]
        table[self.s] = min(old, self.cur)  # atomicMin
        if old == EMPTY:
            self.done = True
            return
        if old > self.cur:  # took the slot from a lower-priority voxel: carry it onward
            self.cur = old
            self.dist = ((self.s - (homes[self.cur] & mask)) & mask) + 1
        else:
            self.dist += 1
        self.s = (self.s + 1) & mask


def concurrent(homes, mask, max_scan, rng, burst):
    table = np.full(mask + 1, EMPTY, np.int64)
    threads = [Thread(v, homes, mask) for v in range(len(homes))]
    dropped = []
    active = list(range(len(threads)))
    while active:
        k = int(rng.integers(len(active)))
        t = threads[active[k]]
        for _ in range(int(rng.integers(1, burst + 1))):  # a thread may run several iterations before another is scheduled
            t.step(table, homes, mask, max_scan, dropped)
            if t.done:
                break
        if t.done:
            active[k] = active[-1]
            active.pop()
    return table, sorted(dropped)


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(st.integers(0, 2**32 - 1), st.sampled_from([8, 16, 64]), st.floats(0.2, 1.3), st.sampled_from([1, 2, 3, 10]), st.sampled_from([1, 4, 64]))
def test_any_schedule_reaches_the_sequential_table(seed, nb, load, max_scan, burst):
    rng = np.random.default_rng(seed)
    V = max(1, int(nb * load))
    # clustered homes (few distinct hash values) make long collision chains and force drops
    homes = rng.integers(0, max(1, nb // int(rng.integers(1, 5))), V).astype(np.int64)
    ref_table, ref_dropped = sequential(homes, nb - 1, max_scan)
    for _ in range(3):
        table, dropped = concurrent(homes, nb - 1, max_scan, rng, burst)
        assert np.array_equal(table, ref_table), (homes.tolist(), max_scan)
        assert dropped == ref_dropped
