"""Randomised layouts: mazes with scattered void tiles (states without some neighbours, dead ends), random action
subsets, ragged sessions.  Every likelihood sweep is run twice -- tables per run of 32 parameter sets and tables per
parameter set -- and both against the oracle: tables bit-exact, log-likelihoods 1e-9, the two mappings bit-identical."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu

NEIGH = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1)]


def _random_maze(seed, X=7, Y=11, size_tile=0.1):
    from navigation_model_b200.simulation.maze import Maze
    rng = np.random.default_rng(seed)
    rows = [[("V" if rng.random() < 0.22 else "CWM"[rng.integers(3)]) for _ in range(Y)] for _ in range(X)]
    rows[1][1] = "M"  # start tile
    rows[1][2] = "W"  # ... with at least one visitable neighbour
    layout = "\n".join("".join(r) for r in rows)
    return Maze(layout, "", size_tile=size_tile), O.OracleMaze(layout, size_tile), rng


def _grid(rng, pairs=3, n_beta=32, drop=6):
    a = np.repeat(rng.uniform(0.01, 1, pairs), n_beta)
    g = np.repeat(rng.uniform(0.5, 0.99, pairs), n_beta)
    b = np.tile(np.logspace(-1, 1.5, n_beta), pairs)
    P = a.size - drop
    return a[:P].copy(), b[:P].copy(), g[:P].copy()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_value_table_sweeps_on_random_mazes(cuda, seed, monkeypatch):
    from navigation_model_b200.sweep import ConditioningSweep
    maze, omz, rng = _random_maze(seed)
    sessions = [O.synthetic_session(omz, 100 * seed + j, T, goal=(1, 2), jitter=0.03) for j, T in enumerate((260, 75))]
    # the null action stays in (every step then has a reachable endpoint: np.max([]) raises in the reference otherwise)
    acts = H.a8(0.1)[np.concatenate([[0], 1 + rng.permutation(7)[:int(rng.integers(2, 8))]])]
    a, b, g = _grid(rng)
    for alg, code in (("td0", O.ALG_TD0), ("positive", O.ALG_POSITIVE), ("negative", O.ALG_NEGATIVE)):
        trs = [H.oracle_transitions(omz, t, r, acts) for t, r in sessions]
        assert all((tr["ncand"] > 0).all() for tr in trs)
        sw = ConditioningSweep(maze, alg, possible_actions=acts, replay_n_times=1, replay_seed=seed)
        for t, r in sessions:
            sw.add_session(t, r)
        ll, V = sw.log_likelihood(a, b, g, return_v=True)
        assert sw.last_sweep_mode() == 1
        monkeypatch.setenv("NM_FIT_GROUP", "0")
        ll_s, V_s = sw.log_likelihood(a, b, g, return_v=True)
        monkeypatch.delenv("NM_FIT_GROUP")
        assert np.array_equal(ll, ll_s) and np.array_equal(V, V_s)
        for j, tr in enumerate(trs):
            Vw, llw = CO.fit(code, tr, a, g, len(omz.vis_cells), beta=b, orders=O.replay_orders(len(tr["s"]), 1, seed))
            np.testing.assert_allclose(ll[j], llw, rtol=1e-9, atol=1e-300)
            for p in (0, 40, a.size - 1):
                assert np.array_equal(V[j, p], O.dense_table(omz, Vw[p]))


@pytest.mark.parametrize("seed", [5, 6, 7, 8])
def test_tile_action_agents_on_random_mazes(cuda, seed, monkeypatch):
    from navigation_model_b200.sweep import ModelBasedSweep, QLearningSweep
    maze, omz, rng = _random_maze(seed)
    sessions = [O.synthetic_session(omz, 100 * seed + j, T, goal=(1, 2), jitter=0.03) for j, T in enumerate((220, 60))]
    moves = [NEIGH[i] for i in sorted(rng.permutation(9)[:int(rng.integers(3, 9))])]
    nxt = O.transition_table(omz, moves)
    a, b, g = _grid(rng)
    picks = (0, 33, a.size - 1)
    ql = QLearningSweep(maze, moves, replay_n_times=1, replay_seed=seed)
    mb = ModelBasedSweep(maze, moves, n_sweeps=2)
    assert np.array_equal(ql.next_table, nxt)
    for t, r in sessions:
        ql.add_session(t, r)
        mb.add_session(t, r)
    for sw, env in ((ql, "NM_Q_GROUP"), (mb, "NM_MB_GROUP")):
        out = sw.log_likelihood(a, b, g, return_tables=True, regroup=False)
        assert sw.last_sweep_mode() == 1
        monkeypatch.setenv(env, "0")
        ref = sw.log_likelihood(a, b, g, return_tables=True, regroup=False)
        assert sw.last_sweep_mode() == (2 if sw is ql else 0)  # tables per parameter set (Q: in device memory)
        for x, y in zip(out, ref):
            assert np.array_equal(x, y)
        if sw is ql:  # ... and the shared-memory octet kernel
            monkeypatch.setenv("NM_Q_TABLES", "shared")
            ref = sw.log_likelihood(a, b, g, return_tables=True, regroup=False)
            assert sw.last_sweep_mode() == 0
            monkeypatch.delenv("NM_Q_TABLES")
            for x, y in zip(out, ref):
                assert np.array_equal(x, y)
        monkeypatch.delenv(env)
    ll_q, Q = ql.log_likelihood(a, b, g, return_tables=True)
    ll_m, V, R = mb.log_likelihood(a, b, g, return_tables=True)
    for j, (t, r) in enumerate(sessions):
        tr = O.extract_transitions(omz, t, r, None)
        orders = O.replay_orders(len(tr["s"]), 1, seed)
        for p in picks:
            Qw, llw = O.run_q_learning(tr, nxt, a[p], b[p], g[p], orders=orders)
            dense = np.zeros((omz.X, omz.Y, len(moves)))
            for i, (x, y) in enumerate(omz.vis_cells):
                dense[x, y] = Qw[i]
            assert np.array_equal(Q[j, p], dense)
            assert abs(ll_q[j, p] - llw) <= 1e-9 * abs(llw) + 1e-300
            Vw, Rw, llw = O.run_model_based(tr, nxt, a[p], b[p], g[p], n_sweeps=2)
            assert np.array_equal(V[j, p], O.dense_table(omz, Vw)) and np.array_equal(R[j, p], O.dense_table(omz, Rw))
            assert abs(ll_m[j, p] - llw) <= 1e-9 * abs(llw) + 1e-300
