"""Tabular Q-learning agent with an explicit Q[s, a] table (EXTENSION, parity unpinned by the reference): CUDA
octet-per-parameter-set kernel vs this repo's numpy oracle.  Q tables bit-exact, log-likelihoods to 1e-9 relative."""
import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu

ACTS8 = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
ACTS5 = [(1, 0), (-1, 0), (0, 1), (0, -1), (0, 0)]


def _dense_q(omz, Q, fill=None):
    out = np.zeros((omz.X, omz.Y, Q.shape[1])) if fill is None else np.array(fill, dtype=np.float64)
    for i, (x, y) in enumerate(omz.vis_cells):
        out[x, y] = Q[i]
    return out


@pytest.mark.parametrize("tables", ["global", "shared"])
@pytest.mark.parametrize("acts,n_times", [(ACTS8, 0), (ACTS8, 2), (ACTS5, 0), (ACTS5, 1)])
def test_qtable_vs_oracle_m9(cuda, acts, n_times, tables, monkeypatch):
    from navigation_model_b200.sweep import QLearningSweep
    omz, maze = H.oracle_m9(), H.product_m9()
    nxt = O.transition_table(omz, acts)
    sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((0, 500), (3, 150), (5, 33))]
    rng = np.random.default_rng(11)
    P = 37  # not a multiple of the four parameter sets a warp carries
    a, b, g = H.grid_params(rng, P)
    ql = QLearningSweep(maze, acts, replay_n_times=n_times, replay_seed=5)
    assert np.array_equal(ql.next_table, nxt)
    for traj, rew in sessions:
        ql.add_session(traj, rew)
    monkeypatch.setenv("NM_Q_TABLES", tables)  # per-set tables in device memory (default) / in shared memory
    ll, Q = ql.log_likelihood(a, b, g, return_tables=True)
    assert ql.last_sweep_mode() == (2 if tables == "global" else 0)
    assert ll.shape == (3, P) and Q.shape == (3, P, omz.X, omz.Y, len(acts))
    for j, (traj, rew) in enumerate(sessions):
        tr = O.extract_transitions(omz, traj, rew, None)
        orders = O.replay_orders(len(tr["s"]), n_times, 5)
        for p in range(P):
            Qw, llw = O.run_q_learning(tr, nxt, a[p], b[p], g[p], orders=orders)
            assert np.array_equal(Q[j, p], _dense_q(omz, Qw)), (j, p)
            assert abs(ll[j, p] - llw) <= 1e-9 * abs(llw) + 1e-300, (j, p, ll[j, p], llw)


def test_qtable_m20_with_initial_table_and_negative_rewards(cuda):
    """20 x 20 all-visitable maze (400 states, 25.6 KB of Q per parameter set), 8 actions, q_init, rewards of both
    signs (the max-backup must not assume non-negative values)."""
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.sweep import QLearningSweep
    layout = "\n".join(["M" * 20] * 20)
    maze = Maze(layout, "", size_tile=0.05)
    omz = O.OracleMaze(layout, 0.05)
    nxt = O.transition_table(omz, ACTS8)
    traj, rew = O.synthetic_session(omz, 1, 400, start=(3, 3), goal=(15, 12), jitter=0.02)
    rng = np.random.default_rng(8)
    rew = np.where(rng.random(len(rew)) < 0.05, -0.7, rew)
    tr = O.extract_transitions(omz, traj, rew, None)
    P = 21
    a, b, g = H.grid_params(rng, P)
    q0 = rng.normal(size=(20, 20, 8))
    ql = QLearningSweep(maze, ACTS8)
    ql.add_session(traj, rew)
    ll, Q = ql.log_likelihood(a, b, g, q_init=q0, return_tables=True)
    q0c = np.array([q0[x, y] for x, y in omz.vis_cells])
    for p in range(0, P, 4):
        Qw, llw = O.run_q_learning(tr, nxt, a[p], b[p], g[p], q0=q0c)
        assert np.array_equal(Q[0, p], _dense_q(omz, Qw, fill=q0))
        assert abs(ll[0, p] - llw) <= 1e-9 * abs(llw)


def test_qtable_extreme_beta_empty_session_and_best_fit(cuda):
    from navigation_model_b200.sweep import QLearningSweep
    omz, maze = H.oracle_m9(), H.product_m9()
    nxt = O.transition_table(omz, ACTS8)
    traj, rew = O.synthetic_session(omz, 2, 300)
    ql = QLearningSweep(maze, ACTS8)
    ql.add_session(traj, rew)
    ql.add_session(traj[:1], rew[:1])  # a session without a single step
    a = np.array([0.5, 0.9, 0.2, 0.3])
    b = np.array([0.0, 5e3, -40.0, 2.0])  # flat policy, near-deterministic policy, inverted preferences
    g = np.array([0.9, 0.99, 0.5, 0.8])
    ll = ql.log_likelihood(a, b, g)
    assert np.all(ll[1] == 0.0) and np.all(np.isfinite(ll[0]))
    tr = O.extract_transitions(omz, traj, rew, None)
    want = np.array([O.run_q_learning(tr, nxt, a[p], b[p], g[p])[1] for p in range(4)])
    assert np.allclose(ll[0], want, rtol=1e-9, atol=0)
    nll, idx = ql.best_fit(a, b, g)
    assert idx == int(np.argmin(-want)) and abs(nll + want[idx]) <= 1e-9 * abs(want[idx])
    with pytest.raises(ValueError):
        ql.log_likelihood(a, b, g, q_init=np.zeros((3, 3, 8)))


@pytest.mark.parametrize("maze_name,acts", [("m9", ACTS8), ("m9", ACTS5), ("m20", ACTS8)])
def test_qtable_group_kernel_matches_octet_kernel_and_oracle(cuda, maze_name, acts, monkeypatch):
    """A grid whose beta axis varies fastest: aligned runs of 32 parameter sets share (alpha, gamma), the table lives
    once per warp (nm_qgroup_kernel).  Bit-identical to the octet kernel (same operations per parameter set), tables
    bit-exact and likelihoods to 1e-9 against the oracle; incomplete last run, replays, initial table, beta < 0."""
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.sweep import QLearningSweep
    if maze_name == "m9":
        omz, maze = H.oracle_m9(), H.product_m9()
        sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((0, 600), (3, 97), (5, 1))]
    else:
        layout = "\n".join(["M" * 20] * 20)
        maze, omz = Maze(layout, "", size_tile=0.05), O.OracleMaze(layout, 0.05)
        sessions = [O.synthetic_session(omz, 1, 300, start=(3, 3), goal=(15, 12), jitter=0.02)]
    nxt = O.transition_table(omz, acts)
    rng = np.random.default_rng(21)
    pairs, n_beta = 4, 32
    a = np.repeat(rng.uniform(0.01, 1, pairs), n_beta)
    g = np.repeat(rng.uniform(0.5, 0.99, pairs), n_beta)
    b = np.tile(np.logspace(-1, 1.5, n_beta), pairs)
    P = a.size - 5
    a, b, g = a[:P].copy(), b[:P].copy(), g[:P].copy()
    b[7] = -b[7]
    q0 = rng.normal(size=(omz.X, omz.Y, len(acts)))
    ql = QLearningSweep(maze, acts, replay_n_times=1, replay_seed=9)
    for traj, rew in sessions:
        ql.add_session(traj, rew)
    ll, Q = ql.log_likelihood(a, b, g, q_init=q0, return_tables=True)
    assert ql.last_sweep_mode() == 1  # the warp-shared-table kernel really ran
    monkeypatch.setenv("NM_Q_GROUP", "0")
    ll_x, Q_x = ql.log_likelihood(a, b, g, q_init=q0, return_tables=True)
    assert ql.last_sweep_mode() == 2  # a table per parameter set, in device memory
    assert np.array_equal(ll, ll_x) and np.array_equal(Q, Q_x)
    monkeypatch.setenv("NM_Q_TABLES", "shared")
    ll_o, Q_o = ql.log_likelihood(a, b, g, q_init=q0, return_tables=True)
    assert ql.last_sweep_mode() == 0  # ... and in shared memory (the octet kernel)
    assert np.array_equal(ll, ll_o) and np.array_equal(Q, Q_o)
    monkeypatch.delenv("NM_Q_GROUP")
    monkeypatch.delenv("NM_Q_TABLES")
    q0c = np.array([q0[x, y] for x, y in omz.vis_cells])
    for j, (traj, rew) in enumerate(sessions):
        tr = O.extract_transitions(omz, traj, rew, None)
        orders = O.replay_orders(len(tr["s"]), 1, 9)
        for p in (0, 31, 32, P - 1):
            Qw, llw = O.run_q_learning(tr, nxt, a[p], b[p], g[p], q0=q0c, orders=orders)
            assert np.array_equal(Q[j, p], _dense_q(omz, Qw, fill=q0)), (j, p)
            assert abs(ll[j, p] - llw) <= 1e-9 * abs(llw) + 1e-300, (j, p)
    # a run that is not uniform: a table per parameter set
    g2 = g.copy()
    g2[40] = 0.777
    ql.log_likelihood(a, b, g2, regroup=False)
    assert ql.last_sweep_mode() == 2


def test_loglikelihood_fitness_drives_the_tile_action_agents(cuda):
    """LogLikelihoodFitness (the objective the sweeps plug into, optimization/fitness.py) with the two extensions."""
    from navigation_model_b200.optimization.fitness import LogLikelihoodFitness
    omz, maze = H.oracle_m9(), H.product_m9()
    sessions = [O.synthetic_session(omz, sid, 200) for sid in (0, 1)]
    rng = np.random.default_rng(2)
    a, b, g = H.grid_params(rng, 40)
    nxt = O.transition_table(omz, ACTS8)
    for name in ("qtable", "modelbased"):
        fit = LogLikelihoodFitness(maze, sessions, algorithm=name, discrete_actions=ACTS8, n_sweeps=2)
        nll = fit(a, b, g)
        assert nll.shape == (40,)
        want = np.zeros(40)
        for traj, rew in sessions:
            tr = O.extract_transitions(omz, traj, rew, None)
            for p in range(40):
                want[p] -= (O.run_q_learning(tr, nxt, a[p], b[p], g[p])[1] if name == "qtable"
                            else O.run_model_based(tr, nxt, a[p], b[p], g[p], n_sweeps=2)[2])
        np.testing.assert_allclose(nll, want, rtol=1e-9)
        v, i = fit.best(a, b, g)
        assert i == int(np.argmin(nll)) and abs(v - nll[i]) <= 1e-9 * abs(nll[i])
    with pytest.raises(ValueError):
        LogLikelihoodFitness(maze, sessions, algorithm="qtable")

