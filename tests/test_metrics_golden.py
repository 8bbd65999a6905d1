"""Host trajectory metrics (navigation_model_b200/analysis/metrics.py) against what the
unmodified reference returned for the same trajectories (tests/golden/metrics_golden.npz, written by
tests/golden/make_golden_metrics.py)."""
import os

import numpy as np
import pytest

import helpers as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gm():
    return dict(np.load(os.path.join(GOLDEN, "metrics_golden.npz"), allow_pickle=False))


def test_orientation_and_bout_metrics_match_reference(golden_sim, gm):
    from navigation_model_b200.analysis import metrics as M
    maze = H.product_m9()
    T = maze.MazeTile
    units = [tuple(int(v) for v in a) for a in golden_sim["actions_units"]]
    for m in range(len(golden_sim["seeds"])):
        pos, ori = golden_sim["hist_pos"][m], list(golden_sim["hist_ori"][m])
        disc = M.discretize_positions(pos, maze)
        h8, rel, e8 = M.hist_orientations(ori, disc)
        assert np.array_equal(h8, gm["ori_hist8"][m]) and np.array_equal(e8, gm["edges8"])
        assert len(rel) == gm["n_relative"][m]
        ha, _, ea = M.hist_orientations(ori, disc, possible_actions=units)
        assert np.array_equal(ha, gm["ori_hist_actions"][m]) and np.array_equal(ea, gm["edges_actions"])
        h6, _, e6 = M.hist_orientations(ori, disc, n_bins=6, distance=True)
        assert np.array_equal(h6, gm["ori_hist6_distance"][m]) and np.array_equal(e6, gm["edges6"])
        hf, _, _ = M.hist_orientations_filtered(ori, disc, maze, [T.CORNER, T.WALL])
        assert np.array_equal(hf, gm["ori_hist_filtered"][m])
        assert np.array_equal(M.relative_orientation_signatures(h8)[1], gm["ori_signature"][m])
        runs = M.hist_time_moving(M.motion_analysis(disc)[1])
        n = int(gm["n_runs"][m])
        assert [r[0] for r in runs] == gm["run_lengths"][m][:n].tolist()
        assert [r[1] for r in runs] == gm["run_kinds"][m][:n].tolist()
        assert np.array_equal(M.mov_bouts(runs)[2], gm["bouts"][m])
        assert M.compute_static_intervals(runs) == gm["static_intervals"][m]
        assert np.array_equal(M.subareas(disc, maze), gm["subareas_hist"][m])
        assert M.latency_enter_shock_zone(disc, maze, subareas_shock_zone=(1, 2)) == gm["latency_sub12"][m]
        assert M.shock_zone_entries(disc, maze, subareas_shock_zone=(1, 2)) == gm["entries_sub12"][m]
        assert M.latency_enter_shock_zone(disc, maze, subareas_shock_zone=3) == gm["latency_sub3"][m]
        assert M.shock_zone_entries(disc, maze, subareas_shock_zone=3) == gm["entries_sub3"][m]
        assert M.latency_enter_shock_zone(disc, maze, tile_lines=5) == gm["latency_lines5"][m]
        assert M.shock_zone_entries(disc, maze, tile_lines=5) == gm["entries_lines5"][m]
        assert M.occupancy_shock_zone(disc, tile_lines=5) == gm["occupancy_lines5"][m]
        assert M.occupancy_shock_zone(disc, tile_lines=2) == gm["occupancy_lines2"][m]
        assert M.arms(disc, maze) == gm["corridor_switches"][m]
        assert np.array_equal(M.histogram_tiles(disc, maze), gm["tiles_histogram"][m])


def test_walk_metrics_match_reference(gm):
    """Long synthetic tile walks that do commute between the corridors and the shock zones."""
    from navigation_model_b200.analysis import metrics as M
    maze = H.product_m9()
    for w, walk in enumerate(gm["walks"]):
        walk = [tuple(int(v) for v in c) for c in walk]
        runs = M.hist_time_moving(M.motion_analysis(walk)[1])
        assert M.arms(walk, maze) == gm["walk_switches"][w] > 0
        assert M.shock_zone_entries(walk, maze, subareas_shock_zone=(1, 2)) == gm["walk_entries_sub12"][w]
        assert M.shock_zone_entries(walk, maze, tile_lines=5) == gm["walk_entries_lines5"][w]
        assert M.latency_enter_shock_zone(walk, maze, subareas_shock_zone=3) == gm["walk_latency_sub3"][w]
        assert M.latency_enter_shock_zone(walk, maze, tile_lines=5) == gm["walk_latency_lines5"][w]
        assert M.occupancy_shock_zone(walk, tile_lines=5) == gm["walk_occupancy_lines5"][w]
        assert np.array_equal(M.subareas(walk, maze), gm["walk_subareas_hist"][w])
        assert M.compute_static_intervals(runs) == gm["walk_static_intervals"][w]
        assert np.array_equal(M.mov_bouts(runs)[2], gm["walk_bouts"][w])


def test_metric_error_behaviour():
    from navigation_model_b200.analysis import metrics as M
    maze = H.product_m9()
    with pytest.raises(RuntimeError, match="Wrong input sizes for hist_orientations"):
        M.hist_orientations([], [1])
    with pytest.raises(RuntimeError, match="Cannot use both n_bins and possible_actions"):
        M.hist_orientations([0.0, 1.0], [0, 1], n_bins=4, possible_actions=[(1, 0), (0, 1)])
    with pytest.raises(RuntimeError, match="Wrong input sizes for hist_orientations_filtered"):
        M.hist_orientations_filtered([0.0], [], maze, [])
    assert M.latency_enter_shock_zone([(7, 0), (7, 1)], maze) == 1   # no zone described: len - 1, like the reference
    with pytest.raises(UnboundLocalError):
        M.shock_zone_entries([(7, 0)], maze)


def test_oracle_distance_formulas_match_reference():
    """The oracle's restated distances (np_oracle.normalized_distance / manhattan_distance / js_distance) and the product's
    host mirror (analysis/statistics.py) against the reference's own outputs (tests/golden/distances_golden.npz)."""
    import os
    from conftest import GOLDEN
    from navigation_model_b200.analysis import statistics as S
    from oracle import np_oracle as O
    g = dict(np.load(os.path.join(GOLDEN, "distances_golden.npz"), allow_pickle=False))
    zero = np.zeros(len(g["data"]))
    for n in range(len(g["x"])):
        x = g["x"][n]
        assert O.normalized_distance(x, g["data"]) == g["nd"][n] and S.normalized_distance(x, g["data"]) == g["nd"][n]
        assert O.manhattan_distance(x, g["data"]) == g["l1"][n] and S.manhattan_distance(x, g["data"]) == g["l1"][n]
        assert abs(O.js_distance(x, g["data"]) - g["js"][n]) <= 1e-15
        assert abs(S.js_distance((x,), (g["data"],)) - g["js"][n]) <= 1e-15
        assert O.normalized_distance(x, zero) == g["nd_zero_data"][n]
