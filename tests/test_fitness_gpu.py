"""Batched distances / multi-objective fitness on the device vs the restated reference formulas
(analysis/statistics.py).  Reduction order differs from numpy's: 1e-12 relative."""
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
