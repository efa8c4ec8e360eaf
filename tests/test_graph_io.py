"""graph.txt round trip (global_mapping.cpp:576-598 save, :846-873 load): the pair list written by save re-creates exactly
the factor list of the sweep."""
import numpy as np
import pytest

from glim_b200 import graph_io
from glim_b200.workloads import Factor


def test_round_trip_recreates_the_sweep(tmp_path):
    rng = np.random.default_rng(5)
    factors, pair = [], 0
    for cur in range(1, 40):
        for i in sorted(rng.choice(cur, size=min(cur, 4), replace=False)):
            for level in (0, 1):
                factors.append(Factor(int(i), level, cur, pair))
            pair += 1
    path = tmp_path / "graph.txt"
    graph_io.write_graph_txt(str(path), 40, 600, [(f.target, f.source) for f in factors])  # levels collapse per pair
    text = path.read_text().splitlines()
    assert text[0] == "num_submaps: 40" and text[1] == "num_all_frames: 600" and text[2] == f"num_matching_cost_factors: {pair}"
    assert text[3].startswith("matching_cost vgicp_gpu ")
    n, m, entries = graph_io.read_graph_txt(str(path))
    assert (n, m, len(entries)) == (40, 600, pair)
    again, skipped = graph_io.recreate_matching_cost_factors(entries, num_levels=2)
    assert not skipped and again == factors
    shifted, _ = graph_io.recreate_matching_cost_factors(entries, 2, start_from_frame_id=100)
    assert shifted[0].target == factors[0].target + 100 and shifted[0].source == factors[0].source + 100


def test_unsupported_types_are_skipped_and_malformed_files_rejected(tmp_path):
    p = tmp_path / "g.txt"
    p.write_text("num_submaps: 3\nnum_all_frames: 9\nnum_matching_cost_factors: 2\nmatching_cost gicp 0 1\nmatching_cost vgicp 0 2\n")
    _, _, entries = graph_io.read_graph_txt(str(p))
    f, skipped = graph_io.recreate_matching_cost_factors(entries, 1)
    assert skipped == [("gicp", 0, 1)] and f == [Factor(0, 0, 2, 1)]
    p.write_text("num_submaps: 3\nnum_all_frames: 9\nnum_matching_cost_factors: 2\nmatching_cost vgicp 0 1\n")
    with pytest.raises(ValueError):
        graph_io.read_graph_txt(str(p))
    with pytest.raises(ValueError):
        graph_io.write_graph_txt(str(p), 1, 1, [], factor_type="ndt")
    graph_io.write_graph_txt(str(p), 1, 1, [])
    assert graph_io.read_graph_txt(str(p)) == (1, 1, [])
