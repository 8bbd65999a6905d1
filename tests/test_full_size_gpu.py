"""BASELINE.json sizes, grids with beta fastest: the run-sharing kernels (one table per warp) against the per-set
kernels on the WHOLE grid -- bit-identical log-likelihoods -- plus oracle spot checks.  The per-set kernels are the
ones the smaller parity tests pin entry by entry."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu


def _grid(n_gamma, n_alpha, n_beta):
    from navigation_model_b200.sweep import parameter_grid
    g, a, b = parameter_grid(np.linspace(0.5, 0.99, n_gamma), np.linspace(0.01, 1, n_alpha), np.logspace(-1, 1.5, n_beta))
    return a, b, g


def test_fit_sweep_configs1_grid(cuda, monkeypatch):
    """configs[1]: 64 x 64 x 32 parameter sets x 5 000 steps, PositiveConditioning + log-likelihood."""
    from navigation_model_b200.sweep import ConditioningSweep
    omz, maze = H.oracle_m9(), H.product_m9()
    traj, rew = O.synthetic_session(omz, 0, 5000)
    a, b, g = _grid(32, 64, 64)
    assert a.size == 131072
    sw = ConditioningSweep(maze, "positive", possible_actions=H.a8())
    sw.add_session(traj, rew)
    ll = sw.log_likelihood(a, b, g)
    assert sw.last_sweep_mode() == 1
    monkeypatch.setenv("NM_FIT_GROUP", "0")
    ll_sets = sw.log_likelihood(a, b, g)
    assert sw.last_sweep_mode() == 2  # a table per parameter set in device memory
    monkeypatch.setenv("NM_FIT_TABLES", "shared")
    ll_shared = sw.log_likelihood(a, b, g)
    assert sw.last_sweep_mode() == 0  # ... in shared memory
    monkeypatch.delenv("NM_FIT_TABLES")
    assert np.array_equal(ll, ll_sets) and np.array_equal(ll, ll_shared)
    idx = np.random.default_rng(3).choice(a.size, 64, replace=False)
    tr = H.oracle_transitions(omz, traj, rew, H.a8())
    _, want = CO.fit(O.ALG_POSITIVE, tr, a[idx], g[idx], 60, beta=b[idx], want_v=False)
    np.testing.assert_allclose(ll[0, idx], want, rtol=1e-9, atol=0)
    # the table does not depend on beta: beta = 0 gives the closed-form likelihood for every (alpha, gamma)
    obs = tr["obs"] != O.OBS_NONE
    closed = 0.0
    for k in tr["ncand"][obs]:
        closed = closed + (0.0 - np.log(float(k)))
    monkeypatch.delenv("NM_FIT_GROUP")
    ll0 = sw.log_likelihood(a[:64], np.zeros(64), g[:64])[0]
    np.testing.assert_allclose(ll0, closed, rtol=1e-12)


def test_qtable_sweep_configs1_grid(cuda, monkeypatch):
    """The same grid and session through the explicit Q[s, a] learner."""
    from navigation_model_b200.sweep import QLearningSweep
    omz, maze = H.oracle_m9(), H.product_m9()
    moves = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    traj, rew = O.synthetic_session(omz, 0, 5000)
    a, b, g = _grid(32, 64, 64)
    ql = QLearningSweep(maze, moves)
    ql.add_session(traj, rew)
    ll = ql.log_likelihood(a, b, g)
    assert ql.last_sweep_mode() == 1
    monkeypatch.setenv("NM_Q_GROUP", "0")
    ll_sets = ql.log_likelihood(a, b, g)
    assert ql.last_sweep_mode() == 2  # a table per parameter set in device memory
    monkeypatch.setenv("NM_Q_TABLES", "shared")
    ll_octet = ql.log_likelihood(a, b, g)
    assert ql.last_sweep_mode() == 0  # ... in shared memory (eight lanes per set)
    monkeypatch.delenv("NM_Q_TABLES")
    assert np.array_equal(ll, ll_sets) and np.array_equal(ll, ll_octet)
    tr = O.extract_transitions(omz, traj, rew, None)
    nxt = O.transition_table(omz, moves)
    for p in np.random.default_rng(4).choice(a.size, 3, replace=False):
        _, want = O.run_q_learning(tr, nxt, a[p], b[p], g[p])
        assert abs(ll[0, p] - want) <= 1e-9 * abs(want)


def test_model_based_sweep_configs3_grid(cuda, monkeypatch):
    """configs[3] shape: 400-state maze, 8 actions, 56 x 56 x 32 = 100 352 parameter sets (1 000 steps here)."""
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.sweep import ModelBasedSweep
    layout = "\n".join(["M" * 20] * 20)
    maze, omz = Maze(layout, "", size_tile=0.05), O.OracleMaze(layout, 0.05)
    moves = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    traj, rew = O.synthetic_session(omz, 0, 1000, start=(3, 3), goal=(15, 12), jitter=0.02)
    a, b, g = _grid(56, 56, 32)
    assert a.size == 100352
    mb = ModelBasedSweep(maze, moves, n_sweeps=1)
    mb.add_session(traj, rew)
    ll = mb.log_likelihood(a, b, g)
    assert mb.last_sweep_mode() == 1
    monkeypatch.setenv("NM_MB_GROUP", "0")
    ll_sets = mb.log_likelihood(a, b, g)
    assert mb.last_sweep_mode() == 0
    assert np.array_equal(ll, ll_sets)
    tr = O.extract_transitions(omz, traj, rew, None)
    nxt = O.transition_table(omz, moves)
    for p in np.random.default_rng(5).choice(a.size, 2, replace=False):
        _, _, want = O.run_model_based(tr, nxt, a[p], b[p], g[p], n_sweeps=1)
        assert abs(ll[0, p] - want) <= 1e-9 * abs(want)


@pytest.mark.parametrize("agent", ["qtable", "modelbased"])
def test_regrouping_of_grids_given_in_another_order(cuda, agent):
    """A grid with gamma fastest (``parameter_grid(alpha, beta, gamma)``): the numpy API brings sets with equal
    (alpha, gamma) together, runs the run-sharing kernel and hands the results back in the caller's order --
    bit-identical to the sweep over the arrays as given."""
    from navigation_model_b200.sweep import ModelBasedSweep, QLearningSweep, parameter_grid
    omz, maze = H.oracle_m9(), H.product_m9()
    moves = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    a, b, g = parameter_grid(np.linspace(0.05, 1, 6), np.logspace(-1, 1.5, 64), np.linspace(0.5, 0.99, 5))
    sw = QLearningSweep(maze, moves) if agent == "qtable" else ModelBasedSweep(maze, moves, n_sweeps=2)
    for sid, T in ((0, 400), (2, 150)):
        sw.add_session(*O.synthetic_session(omz, sid, T))
    assert sw.regroup_order(a, g) is not None
    out = sw.log_likelihood(a, b, g, return_tables=True)
    assert sw.last_sweep_mode() == 1
    ref = sw.log_likelihood(a, b, g, return_tables=True, regroup=False)
    assert sw.last_sweep_mode() == (2 if agent == "qtable" else 0)  # tables per parameter set
    for x, y in zip(out, ref):
        assert np.array_equal(x, y)
    # nothing to do for a grid that is already laid out in runs, nothing possible for 20 beta points
    g2, a2, b2 = parameter_grid(np.linspace(0.5, 0.99, 5), np.linspace(0.05, 1, 6), np.logspace(-1, 1.5, 64))
    assert sw.regroup_order(a2, g2) is None
    a3, b3, g3 = parameter_grid(np.linspace(0.05, 1, 6), np.logspace(-1, 1.5, 20), np.linspace(0.5, 0.99, 5))
    assert sw.regroup_order(a3, g3) is None

