"""Model-based agent (EXTENSION, parity unpinned by the reference): CUDA warp-per-parameter-set kernel vs this
repo's numpy oracle.  V / Rhat tables bit-exact, log-likelihoods to 1e-9 relative."""
import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu

ACTS8 = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
ACTS4 = [(1, 0), (-1, 0), (0, 1), (0, -1), (0, 0)]


def _dense(omz, table, fill=0.0):
    out = np.full((omz.X, omz.Y), fill)
    for i, (x, y) in enumerate(omz.vis_cells):
        out[x, y] = table[i]
    return out


@pytest.mark.parametrize("acts,n_sweeps", [(ACTS8, 1), (ACTS8, 3), (ACTS4, 2), (ACTS8, 0)])
def test_model_based_vs_oracle_m9(cuda, acts, n_sweeps):
    from navigation_model_b200.sweep import ModelBasedSweep
    omz, maze = H.oracle_m9(), H.product_m9()
    nxt = O.transition_table(omz, acts)
    sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((0, 400), (3, 150))]
    rng = np.random.default_rng(4)
    P = 37
    a, b, g = H.grid_params(rng, P)
    mb = ModelBasedSweep(maze, acts, n_sweeps=n_sweeps)
    assert np.array_equal(mb.next_table, nxt)  # product transition table == oracle's
    for traj, rew in sessions:
        mb.add_session(traj, rew)
    ll, V, R = mb.log_likelihood(a, b, g, return_tables=True)
    for j, (traj, rew) in enumerate(sessions):
        tr = O.extract_transitions(omz, traj, rew, None)
        for p in range(P):
            Vw, Rw, llw = O.run_model_based(tr, nxt, a[p], b[p], g[p], n_sweeps=n_sweeps)
            assert np.array_equal(V[j, p], _dense(omz, Vw)), (j, p)
            assert np.array_equal(R[j, p], _dense(omz, Rw)), (j, p)
            assert abs(ll[j, p] - llw) <= 1e-9 * abs(llw) + 1e-300, (j, p, ll[j, p], llw)


def test_model_based_m20_config4_shape(cuda):
    """BASELINE configs[3] shape: 20 x 20 all-visitable maze (400 states), 8 actions."""
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.sweep import ModelBasedSweep
    layout = "\n".join(["M" * 20] * 20)
    maze = Maze(layout, "", size_tile=0.05)
    omz = O.OracleMaze(layout, 0.05)
    nxt = O.transition_table(omz, ACTS8)
    traj, rew = O.synthetic_session(omz, 1, 300, start=(3, 3), goal=(15, 12), jitter=0.02)
    tr = O.extract_transitions(omz, traj, rew, None)
    rng = np.random.default_rng(8)
    P = 40
    a, b, g = H.grid_params(rng, P)
    mb = ModelBasedSweep(maze, ACTS8, n_sweeps=2)
    mb.add_session(traj, rew)
    v0 = rng.random((20, 20))
    ll, V, R = mb.log_likelihood(a, b, g, v_init=v0, return_tables=True)
    for p in range(0, P, 5):
        Vw, Rw, llw = O.run_model_based(tr, nxt, a[p], b[p], g[p], n_sweeps=2, v0=H.compact(omz, v0))
        assert np.array_equal(V[0, p], _dense(omz, Vw)) and np.array_equal(R[0, p], _dense(omz, Rw))
        assert abs(ll[0, p] - llw) <= 1e-9 * abs(llw)


@pytest.mark.parametrize("maze_name,acts,n_sweeps", [("m9", ACTS8, 1), ("m9", ACTS4, 3), ("m20", ACTS8, 2)])
def test_model_based_run_kernel_matches_set_kernel_and_oracle(cuda, maze_name, acts, n_sweeps, monkeypatch):
    """A grid whose beta axis varies fastest: aligned runs of 32 parameter sets share (alpha, gamma), and V, Rhat and
    the planning sweeps are computed once per run (nm_mb_kernel<true>).  Bit-identical to the warp-per-set mapping,
    tables bit-exact and likelihoods to 1e-9 against the oracle; incomplete last run, initial table, beta < 0."""
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.sweep import ModelBasedSweep
    if maze_name == "m9":
        omz, maze = H.oracle_m9(), H.product_m9()
        sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((0, 300), (3, 97))]
    else:
        layout = "\n".join(["M" * 20] * 20)
        maze, omz = Maze(layout, "", size_tile=0.05), O.OracleMaze(layout, 0.05)
        sessions = [O.synthetic_session(omz, 1, 200, start=(3, 3), goal=(15, 12), jitter=0.02)]
    nxt = O.transition_table(omz, acts)
    rng = np.random.default_rng(17)
    pairs, n_beta = 3, 32
    a = np.repeat(rng.uniform(0.01, 1, pairs), n_beta)
    g = np.repeat(rng.uniform(0.5, 0.99, pairs), n_beta)
    b = np.tile(np.logspace(-1, 1.5, n_beta), pairs)
    P = a.size - 7
    a, b, g = a[:P].copy(), b[:P].copy(), g[:P].copy()
    b[5] = -b[5]
    v0 = rng.random((omz.X, omz.Y))
    mb = ModelBasedSweep(maze, acts, n_sweeps=n_sweeps)
    for traj, rew in sessions:
        mb.add_session(traj, rew)
    ll, V, R = mb.log_likelihood(a, b, g, v_init=v0, return_tables=True)
    assert mb.last_sweep_mode() == 1  # the warp-per-run mapping really ran
    monkeypatch.setenv("NM_MB_GROUP", "0")
    ll_x, V_x, R_x = mb.log_likelihood(a, b, g, v_init=v0, return_tables=True)
    assert mb.last_sweep_mode() == 0
    monkeypatch.delenv("NM_MB_GROUP")
    assert np.array_equal(ll, ll_x) and np.array_equal(V, V_x) and np.array_equal(R, R_x)
    for j, (traj, rew) in enumerate(sessions):
        tr = O.extract_transitions(omz, traj, rew, None)
        for p in (0, 31, 32, P - 1):
            Vw, Rw, llw = O.run_model_based(tr, nxt, a[p], b[p], g[p], n_sweeps=n_sweeps, v0=H.compact(omz, v0))
            want_v = v0.copy()  # tiles that are not states keep the initial value
            for i, (x, y) in enumerate(omz.vis_cells):
                want_v[x, y] = Vw[i]
            assert np.array_equal(V[j, p], want_v)
            assert np.array_equal(R[j, p], _dense(omz, Rw))
            assert abs(ll[j, p] - llw) <= 1e-9 * abs(llw) + 1e-300, (j, p)
    g2 = g.copy()
    g2[40] = 0.777  # a run that is not uniform keeps the warp-per-set mapping
    mb.log_likelihood(a, b, g2)
    assert mb.last_sweep_mode() == 0

