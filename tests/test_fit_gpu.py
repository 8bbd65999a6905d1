"""Parity of the CUDA fitness sweep (through the C-ABI) with the oracle and the reference's golden
vectors.  V-tables: bit-exact.  Log-likelihoods: 1e-9 relative (BASELINE.json north_star tolerance; the
device `exp`/`log` differ from glibc's by <= 1 ulp)."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9
ALGS = {"td0": O.ALG_TD0, "positive": O.ALG_POSITIVE, "negative": O.ALG_NEGATIVE}


def _sweep(alg, sessions, actions=True, replay_n_times=0, replay_seed=None, maze=None):
    from navigation_model_b200.sweep import ConditioningSweep
    maze = maze or H.product_m9()
    sw = ConditioningSweep(maze, alg, possible_actions=H.a8(maze.size_tile) if actions else None,
                           replay_n_times=replay_n_times, replay_seed=replay_seed)
    for traj, rew in sessions:
        sw.add_session(traj, rew)
    return sw


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
@pytest.mark.parametrize("n_times", [0, 2])
def test_vtables_bit_exact_vs_oracle(cuda, alg, n_times):
    omz = H.oracle_m9()
    traj, rew = O.synthetic_session(omz, 1, 900)
    if alg == "negative":
        rew = -rew
    tr = H.oracle_transitions(omz, traj, rew, H.a8())
    rng = np.random.default_rng(42)
    P = 257  # not a multiple of the block size
    a, _, gm = H.grid_params(rng, P)
    want, _ = CO.fit(ALGS[alg], tr, a, gm, 60, orders=O.replay_orders(len(tr["s"]), n_times, 5))
    sw = _sweep(alg, [(traj, rew)], replay_n_times=n_times, replay_seed=5)
    got = sw.v_tables(a, gm)  # [1, P, 9, 9]
    assert got.shape == (1, P, 9, 9)
    for p in range(P):
        assert np.array_equal(got[0, p], O.dense_table(omz, want[p])), (alg, n_times, p)


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
def test_loglik_vs_oracle(cuda, alg):
    omz = H.oracle_m9()
    sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((0, 1500), (1, 333), (2, 64), (3, 65))]  # ragged
    if alg == "negative":
        sessions = [(t, -r) for t, r in sessions]
    rng = np.random.default_rng(7)
    P = 300
    a, b, gm = H.grid_params(rng, P)
    sw = _sweep(alg, sessions)
    ll, V = sw.log_likelihood(a, b, gm, return_v=True)
    assert ll.shape == (4, P)
    for j, (traj, rew) in enumerate(sessions):
        tr = H.oracle_transitions(omz, traj, rew, H.a8())
        Vw, llw = CO.fit(ALGS[alg], tr, a, gm, 60, beta=b)
        np.testing.assert_allclose(ll[j], llw, rtol=LL_RTOL, atol=0)
        for p in range(0, P, 17):
            assert np.array_equal(V[j, p], O.dense_table(omz, Vw[p]))


def test_golden_vectors_through_reference_api(cuda, golden_fit):
    """ConditioningExperiment / RLAlgorithm / ShuffledReplays classes of this repo (device path) against
    the V-tables the reference produced for the same calls."""
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.simulation.mouse import Mouse
    from navigation_model_b200.simulation.rl import (ConditioningExperiment, NegativeConditioning,
                                                     PositiveConditioning, ShuffledReplays, TD0)
    g = golden_fit
    maze = H.product_m9()
    traj, rew = g["m9_traj"], g["m9_rew"]

    def run(alg_name, a, gm, r, n_times, seed, v_init=None, extra=False, mz=maze, tj=traj, acts=g["A8"]):
        mouse = Mouse(tj[0], 0.0, acts)
        alg = {"td0": lambda: TD0(a, gm), "positive": lambda: PositiveConditioning(a, gm, mouse, mz),
               "negative": lambda: NegativeConditioning(a, gm, mouse, mz)}[alg_name]()
        rep = ShuffledReplays(n_times=n_times, seed=seed)
        exp = ConditioningExperiment(mz, alg, rep, v_table=None if v_init is None else v_init.copy())
        assert exp._device_eligible()
        exp.run(list(tj), [0.0] * len(tj), list(r), do_replays=n_times > 0)
        if extra:
            exp.do_replays()
        return exp.v_table, rep, mouse

    for alg in ("td0", "positive", "negative"):
        r = -rew if alg == "negative" else rew
        for n_times in (0, 3):
            for p, (a, gm) in enumerate(g["params"]):
                V, rep, mouse = run(alg, a, gm, r, n_times, 11)
                assert np.array_equal(V, g[f"m9_{alg}_rep{n_times}_V"][p]), (alg, n_times, p)
                assert len(rep._buffer) == len(traj) - 1
    # the algorithm's mouse ends where the reference's does
    _, _, mouse = run("positive", 0.1, 0.9, rew, 0, None)
    assert np.array_equal([*mouse.position, mouse.orientation], g["m9_mouse_final"])
    # continued learning + a second do_replays()
    assert np.array_equal(run("td0", 0.2, 0.8, rew, 2, 4, v_init=g["m9_v0"])[0], g["m9_td0_v0_V"])
    assert np.array_equal(run("positive", 0.2, 0.8, rew, 2, 4, v_init=g["m9_v0"], extra=True)[0],
                          g["m9_positive_v0_V"])
    # 7x7 maze with an interior void; points outside the maze
    m7 = Maze(str(g["m7_layout"]), "", size_tile=0.2)
    a7 = O.A8_UNITS * 0.2
    for p, (a, gm) in enumerate(g["params"]):
        assert np.array_equal(run("td0", a, gm, g["m7_rew"], 2, 21, mz=m7, tj=g["m7_traj"], acts=a7)[0],
                              g["m7_td0_rep2_V"][p])
        for alg in ("positive", "negative"):
            r = -g["m7b_rew"] if alg == "negative" else g["m7b_rew"]
            assert np.array_equal(run(alg, a, gm, r, 2, 21, mz=m7, tj=g["m7b_traj"], acts=a7)[0],
                                  g[f"m7b_{alg}_rep2_V"][p])


def test_benchmark_session_kat(cuda, golden_fit):
    omz = H.oracle_m9()
    traj, rew = O.synthetic_session(omz, 0, 5000)
    sw = _sweep("td0", [(traj, rew)])
    V = sw.v_tables([0.1], [0.9])[0, 0]
    assert np.array_equal(V, golden_fit["kat_td0_V"])
    assert V.max() == 0.5949675648481678 and V[7, 4] == 0.48326347528794783


def test_v_init_shared_and_per_unit(cuda):
    omz = H.oracle_m9()
    traj, rew = O.synthetic_session(omz, 4, 400)
    tr = H.oracle_transitions(omz, traj, rew, H.a8())
    rng = np.random.default_rng(3)
    P = 70
    a, b, gm = H.grid_params(rng, P)
    sw = _sweep("positive", [(traj, rew)])
    v0 = rng.random((9, 9))
    got = sw.v_tables(a, gm, v_init=v0)[0]
    want, _ = CO.fit(O.ALG_POSITIVE, tr, a, gm, 60, v0=H.compact(omz, v0))
    for p in range(P):
        assert np.array_equal(got[p], O.dense_table(omz, want[p], v0))
    v0p = rng.random((P, 9, 9))
    got = sw.v_tables(a, gm, v_init=v0p)[0]
    for p in range(0, P, 7):
        w, _ = CO.fit(O.ALG_POSITIVE, tr, a[p:p + 1], gm[p:p + 1], 60, v0=H.compact(omz, v0p[p]))
        assert np.array_equal(got[p], O.dense_table(omz, w[0], v0p[p]))


def test_empty_and_single_step_sessions(cuda):
    omz = H.oracle_m9()
    t1, r1 = O.synthetic_session(omz, 5, 1)   # no transition at all
    t2, r2 = O.synthetic_session(omz, 5, 2)   # exactly one transition
    sw = _sweep("td0", [(t1, r1), (t2, r2)])
    V = sw.v_tables([0.5, 0.25], [0.9, 0.8])
    assert not V[0].any()
    ll = sw.log_likelihood([0.5, 0.25], [1.0, 2.0], [0.9, 0.8])
    assert np.array_equal(ll[0], [0.0, 0.0])
    tr = H.oracle_transitions(omz, t2, r2, H.a8())
    _, llw = CO.fit(O.ALG_TD0, tr, [0.5, 0.25], [0.9, 0.8], 60, beta=[1.0, 2.0])
    np.testing.assert_allclose(ll[1], llw, rtol=LL_RTOL)


def test_full_grid_properties(cuda):
    """BASELINE.json configs[1] at full size (64 x 64 x 32 = 131 072 parameter sets, T = 5 000) through
    size-independent properties + spot checks against the oracle."""
    from navigation_model_b200.sweep import parameter_grid
    omz = H.oracle_m9()
    traj, rew = O.synthetic_session(omz, 0, 5000)
    tr = H.oracle_transitions(omz, traj, rew, H.a8())
    alpha, beta, gamma = parameter_grid(np.linspace(0.01, 1, 64), np.logspace(-1, 1.5, 64), np.linspace(0.5, 0.99, 32))
    P = alpha.size
    assert P == 131072
    sw = _sweep("positive", [(traj, rew)])
    ll = sw.log_likelihood(alpha, beta, gamma)[0]
    assert np.isfinite(ll).all() and (ll < 0).all()
    # (1) spot check against the oracle
    idx = np.random.default_rng(1).choice(P, 96, replace=False)
    _, llw = CO.fit(O.ALG_POSITIVE, tr, alpha[idx], gamma[idx], 60, beta=beta[idx], want_v=False)
    np.testing.assert_allclose(ll[idx], llw, rtol=LL_RTOL)
    # (2) determinism / independence of the block decomposition: a permuted grid gives permuted results
    perm = np.random.default_rng(2).permutation(P)
    ll_perm = sw.log_likelihood(alpha[perm], beta[perm], gamma[perm])[0]
    assert np.array_equal(ll_perm, ll[perm])
    # (3) closed form: alpha = 0 keeps V = 0, so every observed choice has probability 1/K
    obs = tr["obs"] != O.OBS_NONE
    closed = 0.0
    for k in tr["ncand"][obs]:
        closed = closed + (0.0 - np.log(float(k)))
    ll0 = sw.log_likelihood(np.zeros(4), beta[:4], gamma[:4])[0]
    np.testing.assert_allclose(ll0, closed, rtol=1e-12)
    # (4) the arg-min epilogue agrees with numpy
    from navigation_model_b200.sweep import reduce_best
    import torch
    ll_t = torch.from_numpy(ll.reshape(1, P)).to(sw.torch_device)
    best_v, best_i = reduce_best(sw.argmin_device(ll_t), 0)
    assert best_i == int(np.argmin(-ll)) and best_v == float((-ll).min())


@pytest.mark.parametrize("alg,sign", [("positive", 1.0), ("positive", -1.0), ("negative", 1.0), ("negative", -1.0)])
def test_sign_of_beta_taken_once_per_warp(cuda, alg, sign, monkeypatch):
    """The likelihood kernels test the sign of beta once per warp and run a step body compiled for the answer (for the
    max-backup learner and beta >= 0 -- the min-backup learner and beta <= 0 -- the softmax maximum is fl(beta * extremum
    V)).  Warps whose 32 betas all have the favourable sign, warps whose betas all have the other sign and a warp with
    both: the three kernels stay bit-identical to one another and agree with the oracle."""
    omz = H.oracle_m9()
    sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((0, 500), (4, 97))]
    if alg == "negative":
        sessions = [(t, -r) for t, r in sessions]
    rng = np.random.default_rng(17)
    pairs, n_beta = 4, 32
    a = np.repeat(rng.uniform(0.01, 1, pairs), n_beta)
    g = np.repeat(rng.uniform(0.5, 0.99, pairs), n_beta)
    b = sign * np.tile(np.logspace(-1, 1.3, n_beta), pairs)
    b[32:64] = -b[32:64]      # one warp with the other sign throughout
    b[70] = -b[70]            # one warp with both signs
    b[100] = 0.0              # beta = 0 has both signs
    sw = _sweep(alg, sessions)
    ll, V = sw.log_likelihood(a, b, g, return_v=True)
    assert sw.last_sweep_mode() == 1
    monkeypatch.setenv("NM_FIT_GROUP", "0")
    ll_x, V_x = sw.log_likelihood(a, b, g, return_v=True)
    assert sw.last_sweep_mode() == 2
    monkeypatch.setenv("NM_FIT_TABLES", "shared")
    ll_s, V_s = sw.log_likelihood(a, b, g, return_v=True)
    assert sw.last_sweep_mode() == 0
    monkeypatch.delenv("NM_FIT_GROUP")
    monkeypatch.delenv("NM_FIT_TABLES")
    assert np.array_equal(ll, ll_x) and np.array_equal(V, V_x) and np.array_equal(ll, ll_s) and np.array_equal(V, V_s)
    for j, (traj, rew) in enumerate(sessions):
        tr = H.oracle_transitions(omz, traj, rew, H.a8())
        Vw, llw = CO.fit(ALGS[alg], tr, a, g, 60, beta=b)
        np.testing.assert_allclose(ll[j], llw, rtol=LL_RTOL, atol=1e-300)
        for p in (0, 40, 70, 100, 127):
            assert np.array_equal(V[j, p], O.dense_table(omz, Vw[p]))


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
@pytest.mark.parametrize("n_beta", [64, 128, 32, 160])
def test_group_kernels_match_exact_kernel_and_oracle(cuda, alg, n_beta, monkeypatch):
    """Grids whose beta axis varies fastest: aligned runs of 32 consecutive parameter sets share (alpha, gamma) and
    the warp-shared-table kernel takes over.  Its results must be bit-identical to the exact kernel's (same operation
    sequence per parameter set) and agree with the oracle."""
    omz = H.oracle_m9()
    sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((0, 700), (4, 97), (5, 1))]
    if alg == "negative":
        sessions = [(t, -r) for t, r in sessions]
    rng = np.random.default_rng(n_beta)
    pairs = 5
    a = np.repeat(rng.uniform(0.01, 1, pairs), n_beta)
    g = np.repeat(rng.uniform(0.5, 0.99, pairs), n_beta)
    b = np.tile(np.logspace(-1, 1.5, n_beta), pairs)
    P = a.size - 11  # the last run is incomplete
    a, b, g = a[:P].copy(), b[:P].copy(), g[:P].copy()
    b[3] = -b[3]  # mixed signs of beta inside a warp
    v0 = rng.random((9, 9))
    sw = _sweep(alg, sessions, replay_n_times=1, replay_seed=3)
    ll, V = sw.log_likelihood(a, b, g, v_init=v0, return_v=True)
    assert sw.last_sweep_mode() == 1  # the shared-table kernel really ran
    monkeypatch.setenv("NM_FIT_GROUP", "0")
    ll_x, V_x = sw.log_likelihood(a, b, g, v_init=v0, return_v=True)
    assert sw.last_sweep_mode() == 2  # a table per parameter set, in device memory
    monkeypatch.setenv("NM_FIT_TABLES", "shared")
    ll_s, V_s = sw.log_likelihood(a, b, g, v_init=v0, return_v=True)
    assert sw.last_sweep_mode() == 0  # ... and in shared memory
    monkeypatch.delenv("NM_FIT_GROUP")
    monkeypatch.delenv("NM_FIT_TABLES")
    assert np.array_equal(ll, ll_x) and np.array_equal(V, V_x) and np.array_equal(ll, ll_s) and np.array_equal(V, V_s)
    for j, (traj, rew) in enumerate(sessions):
        tr = H.oracle_transitions(omz, traj, rew, H.a8())
        orders = O.replay_orders(len(tr["s"]), 1, 3)
        Vw, llw = CO.fit(ALGS[alg], tr, a, g, 60, beta=b, v0=H.compact(omz, v0), orders=orders)
        np.testing.assert_allclose(ll[j], llw, rtol=LL_RTOL, atol=1e-300)
        for p in (0, 63, 64, P - 1):
            assert np.array_equal(V[j, p], O.dense_table(omz, Vw[p], v0))
    # runs that are uniform in alpha but not in gamma must NOT be grouped: perturb one gamma inside a run
    g2 = g.copy()
    g2[70] = 0.6123
    ll2 = sw.log_likelihood(a, b, g2)
    assert sw.last_sweep_mode() == 2
    monkeypatch.setenv("NM_FIT_TABLES", "shared")
    assert np.array_equal(ll2, sw.log_likelihood(a, b, g2)) and sw.last_sweep_mode() == 0


def test_td0_linearity_in_reward(cuda):
    """TD0 is linear in the reward: scaling rewards by a power of two scales V exactly."""
    omz = H.oracle_m9()
    traj, rew = O.synthetic_session(omz, 6, 2000)
    rng = np.random.default_rng(9)
    a, _, gm = H.grid_params(rng, 1024)
    V1 = _sweep("td0", [(traj, rew)], actions=False).v_tables(a, gm)
    V4 = _sweep("td0", [(traj, 4.0 * rew)], actions=False).v_tables(a, gm)
    assert np.array_equal(4.0 * V1, V4)


def test_best_fit_single_rank(cuda):
    omz = H.oracle_m9()
    sessions = [O.synthetic_session(omz, sid, 500) for sid in (0, 1, 2)]
    rng = np.random.default_rng(11)
    a, b, gm = H.grid_params(rng, 777)
    sw = _sweep("td0", sessions)
    ll = sw.log_likelihood(a, b, gm)
    nll = -(ll[0] + ll[1] + ll[2])
    v, i = sw.best_fit(a, b, gm)
    assert i == int(np.argmin(nll)) and v == float(nll[i])


def test_errors(cuda):
    from navigation_model_b200.sweep import ConditioningSweep
    maze = H.product_m9()
    omz = H.oracle_m9()
    traj, rew = O.synthetic_session(omz, 0, 50)
    with pytest.raises(RuntimeError):  # rl.py:82-83
        ConditioningSweep(maze, "td0").add_session(traj, rew[:-1])
    with pytest.raises(ValueError):
        ConditioningSweep(maze, "positive").add_session(traj, rew)  # needs possible_actions
    sw = ConditioningSweep(maze, "td0")
    sw.add_session(traj, rew)
    with pytest.raises(ValueError):  # likelihood without candidates
        sw.log_likelihood([0.1], [1.0], [0.9])
    # a mouse stranded outside the maze has no candidate: np.max([]) raises ValueError in the reference
    far = traj.copy()
    far[10] = (5.0, 5.0)
    with pytest.raises(ValueError):
        ConditioningSweep(maze, "positive", possible_actions=H.a8()).add_session(far, rew)


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
def test_memo_and_exact_bodies_agree_and_guard_falls_back(cuda, alg, monkeypatch):
    """The memoised-exponential kernel (default) and the exact-formula kernel (NM_FIT_MEMO=0) both match
    the oracle; parameter sets whose |beta * V| leaves the guard (huge beta, v_init far from 0) are
    recomputed by the exact kernel and still match."""
    omz = H.oracle_m9()
    traj, rew = O.synthetic_session(omz, 8, 1200)
    if alg == "negative":
        rew = -rew
    tr = H.oracle_transitions(omz, traj, rew, H.a8())
    rng = np.random.default_rng(21)
    P = 480
    a, b, gm = H.grid_params(rng, P)
    b[::7] = rng.uniform(2e3, 5e4, len(b[::7]))      # |beta * V| >> 700 -> guard -> exact redo
    b[3::11] = rng.uniform(1e-6, 1e-3, len(b[3::11]))  # nearly uniform policy
    a[5::13] = 1.0
    gm[2::17] = 0.999
    _, want = CO.fit(ALGS[alg], tr, a, gm, 60, beta=b, want_v=False)
    assert np.isfinite(want).all()
    sw = _sweep(alg, [(traj, rew)])
    monkeypatch.setenv("NM_FIT_MEMO", "1")
    ll_memo, V_memo = sw.log_likelihood(a, b, gm, return_v=True)
    monkeypatch.setenv("NM_FIT_MEMO", "0")
    ll_exact, V_exact = sw.log_likelihood(a, b, gm, return_v=True)
    np.testing.assert_allclose(ll_memo[0], want, rtol=LL_RTOL, atol=0)
    np.testing.assert_allclose(ll_exact[0], want, rtol=LL_RTOL, atol=0)
    assert np.array_equal(V_memo, V_exact)  # the V recurrence is the same operations in both
    # a starting table far from zero: every parameter set with beta * V0 > 700 is flagged at t = 0
    v0 = np.full((9, 9), 40.0)
    monkeypatch.setenv("NM_FIT_MEMO", "1")
    ll0 = sw.log_likelihood(a, b, gm, v_init=v0)[0]
    _, want0 = CO.fit(ALGS[alg], tr, a, gm, 60, beta=b, v0=H.compact(omz, v0), want_v=False)
    np.testing.assert_allclose(ll0, want0, rtol=LL_RTOL, atol=0)


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
def test_tables_in_device_memory_match_shared_memory_kernels(cuda, alg, monkeypatch):
    """NM_FIT_TABLES=global: the same sweeps with T[state][parameter set] in device memory (the path mazes beyond 863
    states take) -- bit-identical V-tables and log-likelihoods; per-set initial tables, replays, ragged sessions, a
    parameter count that is not a multiple of the CTA size."""
    omz = H.oracle_m9()
    sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((0, 600), (4, 97), (5, 1))]
    if alg == "negative":
        sessions = [(t, -r) for t, r in sessions]
    rng = np.random.default_rng(31)
    P = 301
    a, b, g = H.grid_params(rng, P)
    v0 = rng.random((P, 9, 9))
    sw = _sweep(alg, sessions, replay_n_times=1, replay_seed=2)
    monkeypatch.setenv("NM_FIT_TABLES", "shared")
    want_ll, want_v = sw.log_likelihood(a, b, g, v_init=v0, return_v=True)
    assert sw.last_sweep_mode() == 0
    want_vt = sw.v_tables(a, g, v_init=v0[0])
    monkeypatch.setenv("NM_FIT_TABLES", "global")
    got_ll, got_v = sw.log_likelihood(a, b, g, v_init=v0, return_v=True)
    assert sw.last_sweep_mode() == 2
    got_vt = sw.v_tables(a, g, v_init=v0[0])
    assert np.array_equal(got_ll, want_ll) and np.array_equal(got_v, want_v) and np.array_equal(got_vt, want_vt)


def test_maze_beyond_the_shared_memory_layout(cuda):
    """A 60 x 60 maze with a few void tiles (3 590 states > 863): sweeps and ConditioningExperiment.run keep their
    tables in device memory; V-tables bit-exact and log-likelihoods 1e-9 against the oracle."""
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.simulation.mouse import Mouse
    from navigation_model_b200.simulation.rl import ConditioningExperiment, PositiveConditioning, ShuffledReplays
    from navigation_model_b200.sweep import ConditioningSweep
    rows = [["M"] * 60 for _ in range(60)]
    for x, y in ((7, 7), (7, 8), (20, 33), (41, 5), (58, 58), (8, 6), (30, 30), (31, 30), (12, 50), (50, 12)):
        rows[x][y] = "V"
    layout = "\n".join("".join(r) for r in rows)
    st = 0.02
    maze, omz = Maze(layout, "", size_tile=st), O.OracleMaze(layout, st)
    assert len(omz.vis_cells) == 3590
    acts = H.a8(st)
    traj, rew = O.synthetic_session(omz, 6, 500, start=(6, 6), goal=(7, 6), jitter=0.006)
    tr = H.oracle_transitions(omz, traj, rew, acts)
    rng = np.random.default_rng(6)
    P = 70
    a, b, g = H.grid_params(rng, P)
    sw = ConditioningSweep(maze, "positive", possible_actions=acts, replay_n_times=1, replay_seed=4)
    sw.add_session(traj, rew)
    ll, V = sw.log_likelihood(a, b, g, return_v=True)
    orders = O.replay_orders(len(tr["s"]), 1, 4)
    Vw, llw = CO.fit(O.ALG_POSITIVE, tr, a, g, len(omz.vis_cells), beta=b, orders=orders)
    np.testing.assert_allclose(ll[0], llw, rtol=LL_RTOL, atol=0)
    for p in (0, 33, P - 1):
        assert np.array_equal(V[0, p], O.dense_table(omz, Vw[p]))
    # a grid with beta fastest takes the same device-memory path on this maze
    a2, g2 = np.repeat(a[:2], 32), np.repeat(g[:2], 32)
    b2 = np.tile(np.logspace(-1, 1.5, 32), 2)
    ll2 = sw.log_likelihood(a2, b2, g2)
    _, want2 = CO.fit(O.ALG_POSITIVE, tr, a2, g2, len(omz.vis_cells), beta=b2, orders=orders, want_v=False)
    np.testing.assert_allclose(ll2[0], want2, rtol=LL_RTOL, atol=0)
    # the reference-style single run
    mouse = Mouse(traj[0], 0.0, acts)
    exp = ConditioningExperiment(maze, PositiveConditioning(0.4, 0.8, mouse, maze), ShuffledReplays(0))
    assert exp._device_eligible()
    exp.run(list(traj), [0.0] * len(traj), list(rew))
    want, _ = O.run_session(O.ALG_POSITIVE, tr, 0.4, 0.8, len(omz.vis_cells))
    assert np.array_equal(exp.v_table, O.dense_table(omz, want))


def test_scratch_release_between_sweeps(cuda):
    """nm_scratch_release gives the device-memory tables back; the next sweep allocates again and agrees."""
    from navigation_model_b200 import _native
    omz = H.oracle_m9()
    traj, rew = O.synthetic_session(omz, 2, 300)
    rng = np.random.default_rng(12)
    a, b, g = H.grid_params(rng, 100)
    sw = _sweep("positive", [(traj, rew)])
    first = sw.log_likelihood(a, b, g)
    assert sw.last_sweep_mode() == 2
    assert _native.lib().nm_scratch_release() == 0
    assert np.array_equal(first, sw.log_likelihood(a, b, g))



@pytest.mark.parametrize("alg", ["td0", "positive"])
def test_value_table_sweep_regroups_grids_given_in_any_axis_order(cuda, alg):
    """A grid with gamma fastest (``parameter_grid(alpha, beta, gamma)``): ConditioningSweep brings sets sharing
    (alpha, gamma) side by side, the table-per-warp kernel runs, results come back in the caller's order -- bit-identical
    to the sweep over the arrays as given; best_fit answers in the caller's order and broadcasts scalars."""
    from navigation_model_b200.sweep import parameter_grid
    omz = H.oracle_m9()
    sessions = [O.synthetic_session(omz, sid, T) for sid, T in ((1, 900), (2, 400))]  # both reach the goal
    a, b, g = parameter_grid(np.linspace(0.05, 1, 6), np.logspace(-1, 1.5, 64), np.linspace(0.5, 0.99, 5))
    sw = _sweep(alg, sessions)
    ll, V = sw.log_likelihood(a, b, g, return_v=True)
    assert sw.last_sweep_mode() == 1
    ll_x, V_x = sw.log_likelihood(a, b, g, return_v=True, regroup=False)
    assert sw.last_sweep_mode() == 2
    assert np.array_equal(ll, ll_x) and np.array_equal(V, V_x)
    nll = -(ll[0] + ll[1])
    v, i = sw.best_fit(a, b, g)
    assert i == int(np.argmin(nll)) and v == nll[i]
    assert sw.best_fit(a, b, g, regroup=False) == (v, i)
    # scalar alpha and gamma against a beta axis (only one set was evaluated)
    betas = np.logspace(1.5, -1, 40)  # random-walk sessions: the flattest policy fits best -> the LAST set
    ll1 = sw.log_likelihood(0.3, betas, 0.9)
    v1, i1 = sw.best_fit(0.3, betas, 0.9)
    assert i1 == int(np.argmin(-(ll1[0] + ll1[1]))) and i1 != 0
    with pytest.raises(ValueError):
        sw.best_fit([0.1, 0.2], betas, 0.9)
