"""BatchedExperiment.run_pipelined: a population cut into blocks of whole CTAs that travel through two host threads / CUDA
streams must give exactly what one run() gives -- every per-mouse array, the population occupancy, the fused metrics."""
import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu


def _same(a, b):
    for name, va in vars(a).items():
        vb = getattr(b, name)
        if va is None:
            assert vb is None, name
        elif isinstance(va, np.ndarray):
            assert va.shape == vb.shape and np.array_equal(va, vb), name
        else:
            assert va == vb, name


def _experiment(v_table):
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    return BatchedExperiment(H.product_m9(), H.a8(), O.A8_UNITS,
                             [bc.AnxietyComponent, bc.BiomechanicalCostComponent, bc.ExperienceComponent,
                              bc.BiomechPersistenceComponent, bc.ValueTableComponent], v_table=v_table)


def test_pipelined_equals_one_run_device_seeded(cuda):
    import bench
    rng = np.random.default_rng(9)
    N, T = 5000, 120   # 40 CTAs: blocks of 10 CTAs, the last one ragged
    maze = H.product_m9()
    P = bench.sim_population(rng, N)
    P["beta"] = 2.5    # a scalar column passes through
    exp = _experiment(rng.random((9, 9)))
    pos0 = np.array(maze.get_tile_centre(1, 1)) + rng.uniform(-0.03, 0.03, (N, 2))
    ori0 = rng.uniform(-3, 3, N)
    kw = dict(init_disc=(1, 1), device_seed=31, n_mice=N, occupancy=True, occupancy_total=True,
              shock_zone={"tile_lines": 5}, corridors=(1, 7), static_intervals=True, subareas=True,
              orientation_histogram={})
    one = exp.run(P, T, pos0, ori0, **kw)
    for chunks in (4, 3, 64):
        _same(one, exp.run_pipelined(P, T, pos0, ori0, chunks=chunks, **kw))
    _same(one, exp.run(P, T, pos0, ori0, **kw))   # and the object still runs the plain way afterwards
    with pytest.raises(ValueError):
        exp.run_pipelined(P, T, pos0, ori0, history=True, **kw)


def test_pipelined_equals_one_run_numpy_seeds_and_per_mouse_tables(cuda):
    import bench
    rng = np.random.default_rng(10)
    N, T = 700, 60
    maze = H.product_m9()
    P = bench.sim_population(rng, N)
    exp = _experiment(rng.random((N, 9, 9)))   # a value table per mouse: sliced with the blocks
    pos0 = np.array(maze.get_tile_centre(1, 1)) + rng.uniform(-0.03, 0.03, (N, 2))
    seeds = list(range(100, 100 + N))
    init_visits = rng.integers(0, 5, (N, 9, 9))
    one = exp.run(P, T, pos0, 0.3, init_disc=(1, 1), seeds=seeds, init_visits=init_visits)
    two = exp.run_pipelined(P, T, pos0, 0.3, init_disc=(1, 1), seeds=seeds, init_visits=init_visits, chunks=3)
    _same(one, two)
    draws = rng.random((N, T))
    _same(exp.run(P, T, pos0, 0.3, init_disc=(1, 1), draws=draws),
          exp.run_pipelined(P, T, pos0, 0.3, init_disc=(1, 1), draws=draws, chunks=5))
