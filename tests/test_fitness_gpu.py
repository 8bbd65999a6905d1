"""Batched distances / multi-objective fitness on the device vs distances the UNMODIFIED reference computed
(tests/golden/distances_golden.npz, analysis/statistics.py:10-54, 67-103) and vs the oracle's restated formulas (pinned to
the same fixture on the CPU, tests/test_metrics_golden.py).  Reduction order differs from numpy's: 1e-12 relative."""
import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu


def test_distance_rows_vs_reference_formulas(cuda):
    from navigation_model_b200.batched_fitness import BatchedMultiObjectiveFitness
    rng = np.random.default_rng(0)
    N, L = 300, 81
    data = rng.integers(0, 50, L).astype(np.float64)
    x = rng.integers(0, 50, (N, L)).astype(np.float64)
    x[0] = 0.0            # empty histogram
    x[1] = data           # identical
    x[2, :] = 0.0
    x[2, 5] = 7.0         # single bin
    fit = BatchedMultiObjectiveFitness({"h": data, "z": np.zeros(L)})
    fit.add_objective("nd", "normalized", "h")
    fit.add_objective("l1", "manhattan", "h")
    fit.add_objective("js", "js", "h")
    fit.add_objective("nd_zero_data", "normalized", "z")
    with pytest.raises(RuntimeError):
        fit.add_objective("nd", "normalized", "h")
    d = fit({"h": x, "z": x})
    for n in range(N):
        np.testing.assert_allclose(d["nd"][n], O.normalized_distance(x[n], data), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(d["l1"][n], O.manhattan_distance(x[n], data), rtol=1e-12)
        np.testing.assert_allclose(d["js"][n], O.js_distance(x[n], data), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(d["nd_zero_data"][n], O.normalized_distance(x[n], np.zeros(L)), rtol=1e-12)
    assert d["nd"][1] == 0.0 and abs(d["js"][0] - np.sqrt(np.log(2.0) / 2.0)) < 1e-12  # identical; empty vs data
    # ... and against the reference's own outputs on the same seeded inputs
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "distances_golden.npz"), allow_pickle=False))
    assert np.array_equal(g["x"], x) and np.array_equal(g["data"], data)
    np.testing.assert_allclose(d["nd"], g["nd"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(d["l1"], g["l1"], rtol=1e-12)
    np.testing.assert_allclose(d["js"], g["js"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(d["nd_zero_data"], g["nd_zero_data"], rtol=1e-12)


def test_simulation_to_fitness_stays_on_device(cuda):
    """simulate -> fused occupancy / tile counts -> distances to data metrics, one value per mouse."""
    from navigation_model_b200.batched_fitness import BatchedMultiObjectiveFitness
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    maze, omz = H.product_m9(), H.oracle_m9()
    rng = np.random.default_rng(3)
    N, T = 48, 250
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, [bc.AnxietyComponent, bc.ExperienceComponent])
    u = rng.uniform
    P = {"beta": u(0.5, 4, N), "W_anx": u(0, 5, N), "p1": u(0, 1, N), "p2": u(0, 1, N), "p3": u(0, 1, N),
         "W_exp": u(0, 5, N), "xi": u(0.006, 0.6, N), "kappa_home": u(0.01, 5, N), "kappa_explo": u(0.01, 5, N),
         "theta_explo": u(80, 120, N), "theta_home": u(30, 70, N)}
    draws = rng.random((N, T))
    res = exp.run(P, T, maze.get_tile_centre(1, 1), 0.0, init_disc=(1, 1), draws=draws, occupancy=True)
    data_occ = rng.integers(0, 20, (9, 9)).astype(np.float64) * (maze.visitable_mask)
    data = {"occupancy": data_occ, "tile_counts": np.array([0, 90, 30, 131, 0, 0, 0, 0.0]), "count_moving": 120.0}
    fit = BatchedMultiObjectiveFitness(data)
    fit.add_objective("occ", "normalized", "occupancy")
    fit.add_objective("tiles", "js", "tile_counts")
    fit.add_objective("moving", "manhattan", "count_moving")
    d = fit({"occupancy": res.occupancy, "tile_counts": res.tile_counts, "count_moving": res.count_moving})
    for n in range(N):
        np.testing.assert_allclose(d["occ"][n], O.normalized_distance(res.occupancy[n].ravel(), data_occ.ravel()),
                                   rtol=1e-12)
        np.testing.assert_allclose(d["tiles"][n], O.js_distance(res.tile_counts[n], data["tile_counts"]), rtol=1e-10)
        assert d["moving"][n] == abs(res.count_moving[n] - 120.0)
    # the best mouse by a weighted sum of the objectives (one rank: no exchange; `offset` = the block's first index)
    w = {"occ": 1.0, "tiles": 2.0, "moving": 0.01}
    total = d["occ"] + 2.0 * d["tiles"] + 0.01 * d["moving"]
    v, i = fit.best({"occupancy": res.occupancy, "tile_counts": res.tile_counts, "count_moving": res.count_moving},
                    weights=w, offset=1000)
    assert i == 1000 + int(np.argmin(total)) and v == total.min()


def test_reference_style_objectives_from_fused_metrics(cuda, golden_sim):
    """The whole reference pipeline (experiment -> Analyzer metrics -> MultiObjectiveFitness) for a
    batch: fused metrics under the reference's product names -> batched distances; every value equals the host
    ``MultiObjectiveFitness`` applied to the host ``Analyzer`` results of the same mouse's history."""
    from navigation_model_b200.analysis import metrics as M
    from navigation_model_b200.analysis.statistics import js_distance, manhattan_distance, normalized_distance
    from navigation_model_b200.batched_fitness import BatchedMultiObjectiveFitness
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.optimization.fitness import MultiObjectiveFitness
    from test_sim_gpu import _classes
    g = golden_sim
    maze = H.product_m9()
    names = [str(s) for s in g["param_names"]]
    P = {n: g["params"][:, i] for i, n in enumerate(names)}
    T = int(g["T"])
    exp = BatchedExperiment(maze, g["actions"], g["actions_units"], _classes(), v_table=g["v_table"])
    r = exp.run(P, T, g["init_pose"][:, :2], g["init_pose"][:, 2], init_disc=g["init_disc"], draws=g["draws"],
                history=True, occupancy=True, shock_zone={"tile_lines": 5}, corridors=(1, 7), static_intervals=True,
                subareas=True, orientation_histogram={})
    results = r.as_results(maze)
    data = {"orientations_histogram": np.array([1, 10, 5, 9, 3, 8, 0, 120.0]),
            "subareas_histogram": np.array([0, 150, 60, 20, 10, 40, 80, 241.0]),
            "moving_dynamics_histogram": np.array([180.0, 421.0]), "static_intervals": 60.0,
            "shock_zone_entries": 3.0, "tiles_histogram": np.array([9.0, 7.5, 6.0])}
    spec = [("ori", "js", js_distance, "orientations_histogram"), ("sub", "normalized", normalized_distance,
                                                                   "subareas_histogram"),
            ("bouts", "normalized", normalized_distance, "moving_dynamics_histogram"),
            ("static", "manhattan", manhattan_distance, "static_intervals"),
            ("entries", "manhattan", manhattan_distance, "shock_zone_entries"),
            ("tiles", "normalized", normalized_distance, "tiles_histogram")]
    dev_fit = BatchedMultiObjectiveFitness(data)
    host_fit = MultiObjectiveFitness(data)
    for name, kind, fn, key in spec:
        dev_fit.add_objective(name, kind, key)
        host_fit.add_objective(name, fn, key)
    d = dev_fit(results)
    ana = M.Analyzer(M.discretize_positions, M.motion_analysis, M.hist_time_moving, M.mov_bouts,
                     M.compute_static_intervals, M.hist_orientations, M.subareas, M.shock_zone_entries,
                     M.histogram_tiles)
    for m in range(len(g["seeds"])):
        host_res = ana.analize(continuous_positions=r.history_position[m], maze=maze, tile_lines=5,
                               absolute_orientations=list(r.history_orientation[m]))
        want = host_fit(host_res)
        for name, *_ in spec:
            np.testing.assert_allclose(d[name][m], want[name], rtol=1e-10, atol=1e-12, err_msg=f"{name} mouse {m}")
