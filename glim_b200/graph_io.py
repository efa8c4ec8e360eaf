"""On-disk round trip of the matching-cost part of a global map (second half).

GLIM does NOT serialise matching-cost factors: GlobalMapping::save writes one text line per (target, source) pair into
`graph.txt` (src/glim/mapping/global_mapping.cpp:576-598) and GlobalMapping::load re-creates one factor per voxel-map level of
the target for every line (:846-873).  After a load the whole graph is relinearized at once -- the many-factor sweep this
repository accelerates -- so the pair list -> factor list step is the host-side entry to that sweep.  Host logic only.

    num_submaps: 256
    num_all_frames: 3840
    num_matching_cost_factors: 3415
    matching_cost vgicp_gpu 0 1
    ...
"""
from __future__ import annotations

from .workloads import Factor

TYPES = ("gicp", "vgicp", "vgicp_gpu")


def write_graph_txt(path: str, num_submaps: int, num_all_frames: int, pairs, factor_type: str = "vgicp_gpu") -> None:
    """pairs: iterable of (target_index, source_index).  One line per pair whatever the number of voxel levels: the reference keys
    its table by "i_j" (global_mapping.cpp:563-567), so the levels of a pair collapse into one entry."""
    if factor_type not in TYPES:
        raise ValueError(f"unknown matching cost factor type {factor_type!r}")
    seen, lines = set(), []
    for i, j in pairs:
        if (int(i), int(j)) in seen:
            continue
        seen.add((int(i), int(j)))
        lines.append(f"matching_cost {factor_type} {int(i)} {int(j)}")
    with open(path, "w") as f:
        f.write(f"num_submaps: {int(num_submaps)}\n")
        f.write(f"num_all_frames: {int(num_all_frames)}\n")
        f.write(f"num_matching_cost_factors: {len(lines)}\n")
        f.write("\n".join(lines) + ("\n" if lines else ""))


def read_graph_txt(path: str):
    """-> (num_submaps, num_all_frames, [(type, first, second), ...]); raises ValueError on a malformed file (the reference
    logs an error and returns false, global_mapping.cpp:700-720)."""
    with open(path) as f:
        tokens = f.read().split()
    pos = 0

    def expect(tag):
        nonlocal pos
        if pos + 1 >= len(tokens) or tokens[pos] != tag:
            raise ValueError(f"graph.txt: expected {tag!r}")
        value = int(tokens[pos + 1])
        pos += 2
        return value

    num_submaps = expect("num_submaps:")
    num_all_frames = expect("num_all_frames:")
    count = expect("num_matching_cost_factors:")
    out = []
    for _ in range(count):
        if pos + 3 >= len(tokens) or tokens[pos] != "matching_cost":
            raise ValueError("graph.txt: truncated matching_cost list")
        out.append((tokens[pos + 1], int(tokens[pos + 2]), int(tokens[pos + 3])))
        pos += 4
    return num_submaps, num_all_frames, out


def recreate_matching_cost_factors(entries, num_levels: int, start_from_frame_id: int = 0):
    """GlobalMapping::load (:846-873): for every ("vgicp" | "vgicp_gpu", first, second) one factor per voxel-map level of
    submaps[first], source = subsampled_submaps[second]; other types are skipped with a warning in the reference."""
    factors, skipped = [], []
    for pair, (ftype, first, second) in enumerate(entries):
        if ftype not in ("vgicp", "vgicp_gpu"):
            skipped.append((ftype, first, second))
            continue
        for level in range(num_levels):
            factors.append(Factor(first + start_from_frame_id, level, second + start_from_frame_id, pair))
    return factors, skipped
