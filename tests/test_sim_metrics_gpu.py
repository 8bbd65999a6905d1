"""Fused trajectory metrics of the simulation kernel (orientation histograms, not-moving
runs, subareas, shock zone, corridor switches) against (1) what the unmodified reference returned for its own
trajectories (tests/golden/metrics_golden.npz; the kernel replays the recorded draws) and (2) the host metric
functions applied to the histories the kernel itself returns, for a few hundred random mice.  All integer: exact."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H
from test_sim_gpu import _classes, _random_params

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gm():
    return dict(np.load(os.path.join(GOLDEN, "metrics_golden.npz"), allow_pickle=False))


def test_fused_metrics_match_reference_golden(cuda, golden_sim):
    from navigation_model_b200.batched_sim import BatchedExperiment
    g, gm = golden_sim, _gm()
    maze = H.product_m9()
    T_ = maze.MazeTile
    names = [str(s) for s in g["param_names"]]
    P = {n: g["params"][:, i] for i, n in enumerate(names)}
    T = int(g["T"])
    units = [tuple(int(v) for v in a) for a in g["actions_units"]]
    exp = BatchedExperiment(maze, g["actions"], g["actions_units"], _classes(), v_table=g["v_table"])
    kw = dict(init_pos=g["init_pose"][:, :2], init_ori=g["init_pose"][:, 2], init_disc=g["init_disc"], draws=g["draws"])
    r = exp.run(P, T, shock_zone={"subareas": (1, 2)}, corridors=(1, 7), static_intervals=True, subareas=True,
                orientation_histogram={}, **kw)
    assert not r.status.any()
    assert np.array_equal(r.orientations_histogram, gm["ori_hist8"])
    assert np.array_equal(r.orientations_histogram_bins, gm["edges8"])
    assert np.array_equal(r.rel_ori_signature(), gm["ori_signature"])
    assert np.array_equal(r.static_intervals, gm["static_intervals"])
    assert np.array_equal(r.mov_bouts(), gm["bouts"])
    assert np.array_equal(r.subareas_histogram(maze), gm["subareas_hist"])
    assert np.array_equal(r.zone_latency, gm["latency_sub12"])
    assert np.array_equal(r.zone_entries, gm["entries_sub12"])
    assert np.array_equal(r.corridor_switches, gm["corridor_switches"])
    assert np.array_equal(r.tiles_histogram(maze), gm["tiles_histogram"])
    r = exp.run(P, T, shock_zone={"subareas": 3}, orientation_histogram={"possible_actions": units}, **kw)
    assert np.array_equal(r.orientations_histogram, gm["ori_hist_actions"])
    assert np.array_equal(r.zone_latency, gm["latency_sub3"]) and np.array_equal(r.zone_entries, gm["entries_sub3"])
    r = exp.run(P, T, shock_zone={"tile_lines": 5}, orientation_histogram={"n_bins": 6, "distance": True}, **kw)
    assert np.array_equal(r.orientations_histogram, gm["ori_hist6_distance"])
    assert np.array_equal(r.zone_latency, gm["latency_lines5"])
    assert np.array_equal(r.zone_entries, gm["entries_lines5"])
    assert np.array_equal(r.zone_occupancy, gm["occupancy_lines5"])
    r = exp.run(P, T, shock_zone={"tile_lines": 2}, orientation_histogram={"tiles_filter": [T_.CORNER, T_.WALL]}, **kw)
    assert np.array_equal(r.orientations_histogram, gm["ori_hist_filtered"])
    assert np.array_equal(r.zone_occupancy, gm["occupancy_lines2"])


def test_fused_metrics_match_host_functions_on_device_histories(cuda):
    """300 mice x 500 steps with their own PCG64 streams: every fused metric equals the host metric function applied
    to the history the same launch returned."""
    from navigation_model_b200.analysis import metrics as M
    from navigation_model_b200.batched_sim import BatchedExperiment
    maze = H.product_m9()
    T_ = maze.MazeTile
    rng = np.random.default_rng(5)
    N, T = 300, 500
    P = _random_params(rng, N)
    P["beta"] = rng.uniform(0.05, 1.0, N)  # exploratory mice: they do reach the corridors and the zones
    P["W_exp"] = rng.uniform(5, 10, N)
    V = rng.random((9, 9))
    cells = np.array(H.oracle_m9().vis_cells)
    start = cells[rng.integers(len(cells), size=N)]
    pos0 = np.array([maze.get_tile_centre(int(c[0]), int(c[1])) for c in start]) + rng.uniform(-0.03, 0.03, (N, 2))
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, _classes(), v_table=V)
    kw = dict(init_pos=pos0, init_ori=rng.uniform(-3, 3, N), seeds=list(range(100, 100 + N)), history=True)
    runs = [
        (dict(shock_zone={"tile_lines": 5}, corridors=(1, 7), static_intervals=True, subareas=True,
              orientation_histogram={}), {}),
        (dict(shock_zone={"subareas": (3, 4)}, corridors=(3, 5), orientation_histogram={"n_bins": 6, "distance": True}),
         dict(n_bins=6, distance=True)),
        (dict(orientation_histogram={"tiles_filter": [T_.CENTRE]}), None),
        (dict(orientation_histogram={"possible_actions": [tuple(a) for a in O.A8_UNITS.tolist()]}),
         dict(possible_actions=[tuple(a) for a in O.A8_UNITS.tolist()])),
    ]
    seen_switch = seen_entries = 0
    for opts, ori_kw in runs:
        r = exp.run(P, T, **opts, **kw)
        assert not r.status.any()
        for m in range(N):
            disc = maze.cont2disc_list(r.history_position[m])
            ori = list(r.history_orientation[m])
            if ori_kw is None:
                want = M.hist_orientations_filtered(ori, disc, maze, [T_.CENTRE])[0]
            else:
                want = M.hist_orientations(ori, disc, **ori_kw)[0]
            assert np.array_equal(r.orientations_histogram[m], want), (opts, m)
            if "shock_zone" in opts:
                z = opts["shock_zone"]
                zk = dict(subareas_shock_zone=z["subareas"]) if "subareas" in z else dict(tile_lines=z["tile_lines"])
                assert r.zone_latency[m] == M.latency_enter_shock_zone(disc, maze, **zk), (opts, m)
                assert r.zone_entries[m] == M.shock_zone_entries(disc, maze, **zk), (opts, m)
                if "tile_lines" in z:
                    assert r.zone_occupancy[m] == M.occupancy_shock_zone(disc, tile_lines=z["tile_lines"])
                seen_entries += int(r.zone_entries[m] > 1)
            if opts.get("corridors") == (1, 7):
                assert r.corridor_switches[m] == M.arms(disc, maze), (opts, m)
                seen_switch += int(r.corridor_switches[m] > 0)
            if opts.get("static_intervals"):
                moved = M.motion_analysis(disc)[1]
                hist = M.hist_time_moving(moved)
                assert r.static_intervals[m] == M.compute_static_intervals(hist)
                assert np.array_equal(r.mov_bouts()[m], M.mov_bouts(hist)[2])
            if opts.get("subareas"):
                assert np.array_equal(r.subareas_histogram(maze)[m], M.subareas(disc, maze))
                assert np.array_equal(r.tiles_histogram(maze)[m], M.histogram_tiles(disc, maze))
    assert seen_entries > 10 and seen_switch > 0   # the state machines were actually exercised


def test_fused_metric_argument_errors(cuda):
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    from navigation_model_b200.simulation.maze import Maze
    maze = H.product_m9()
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, [bc.AnxietyComponent])
    P = {"beta": 1.0, "W_anx": 1.0, "p1": 1.0, "p2": 0.5, "p3": 0.5}
    kw = dict(init_pos=maze.get_tile_centre(1, 1), draws=np.full((1, 5), 0.5))
    with pytest.raises(ValueError, match="six-bin"):
        exp.run(P, 5, orientation_histogram={"distance": True}, **kw)
    with pytest.raises(RuntimeError, match="Cannot use both n_bins and possible_actions"):
        exp.run(P, 5, orientation_histogram={"n_bins": 4, "possible_actions": [(1, 0), (0, 1)]}, **kw)
    with pytest.raises(ValueError, match="shock_zone must be"):
        exp.run(P, 5, shock_zone={"subareas": 1, "tile_lines": 2}, **kw)
    with pytest.raises(TypeError, match="unknown orientation_histogram options"):
        exp.run(P, 5, orientation_histogram={"bins": 3}, **kw)
    bare = Maze("CWC\nWMW\nCWC", "", size_tile=0.1)  # no subareas
    exp2 = BatchedExperiment(bare, H.a8() / 0.111 * 0.1, O.A8_UNITS, [bc.AnxietyComponent])
    with pytest.raises(RuntimeError, match="no subareas"):
        exp2.run(P, 5, init_pos=bare.get_tile_centre(1, 1), draws=np.full((1, 5), 0.5), subareas=True)
