"""The oracle (numpy and C restatements) against the outputs of the unmodified reference.

Fixtures: tests/golden/fit_golden.npz, written by tests/golden/make_golden.py from the reference's own
ConditioningExperiment / RLAlgorithm / ShuffledReplays (simulation/rl.py).  Bit-exact comparisons.
"""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O

import helpers as H

ALGS = {"td0": O.ALG_TD0, "positive": O.ALG_POSITIVE, "negative": O.ALG_NEGATIVE}


def _m9_transitions(g):
    mz = H.oracle_m9()
    return mz, H.oracle_transitions(mz, g["m9_traj"], g["m9_rew"], g["A8"])


def test_transitions_match_reference(golden_fit):
    g = golden_fit
    mz, tr = _m9_transitions(g)
    cells = np.array(mz.vis_cells)
    assert np.array_equal(cells[tr["s"]], g["m9_tr_s"])
    assert np.array_equal(cells[tr["s1"]], g["m9_tr_s1"])
    assert np.array_equal(tr["r"], g["m9_tr_r"])
    assert np.array_equal(tr["ncand"], g["m9_tr_ncand"])
    for i in range(len(tr["s"])):
        k = tr["ncand"][i]
        assert np.array_equal(cells[tr["cand"][i, :k]], g["m9_tr_cand"][i, :k])


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
@pytest.mark.parametrize("n_times", [0, 3])
def test_vtables_match_reference(golden_fit, alg, n_times):
    g = golden_fit
    mz, tr = _m9_transitions(g)
    if alg == "negative":
        tr = dict(tr, r=-tr["r"])
    n = len(tr["s"])
    want = g[f"m9_{alg}_rep{n_times}_V"]
    alphas, gammas = g["params"][:, 0], g["params"][:, 1]
    Vc, _ = CO.fit(ALGS[alg], tr, alphas, gammas, len(mz.vis_cells), orders=O.replay_orders(n, n_times, 11))
    for p, (a, gm) in enumerate(g["params"]):
        Vp, _ = O.run_session(ALGS[alg], tr, a, gm, len(mz.vis_cells), orders=O.replay_orders(n, n_times, 11))
        assert np.array_equal(O.dense_table(mz, Vp), want[p]), (alg, n_times, p)
        assert np.array_equal(Vc[p], Vp)


def test_continued_learning_and_second_replay(golden_fit):
    g = golden_fit
    mz, tr = _m9_transitions(g)
    n = len(tr["s"])
    v0 = H.compact(mz, g["m9_v0"])
    V, _ = O.run_session(O.ALG_TD0, tr, 0.2, 0.8, len(v0), v0=v0, orders=O.replay_orders(n, 2, 4))
    assert np.array_equal(O.dense_table(mz, V, g["m9_v0"]), g["m9_td0_v0_V"])
    # run(do_replays=True) then do_replays() again: 4 cumulative shuffles of the same generator
    V, _ = O.run_session(O.ALG_POSITIVE, tr, 0.2, 0.8, len(v0), v0=v0, orders=O.replay_orders(n, 4, 4))
    assert np.array_equal(O.dense_table(mz, V, g["m9_v0"]), g["m9_positive_v0_V"])
    Vc, _ = CO.fit(O.ALG_POSITIVE, tr, [0.2], [0.8], len(v0), v0=v0, orders=O.replay_orders(n, 4, 4))
    assert np.array_equal(Vc[0], V)


def test_maze_with_void_and_outside_points(golden_fit):
    g = golden_fit
    mz = O.OracleMaze(str(g["m7_layout"]), 0.2)
    tr = O.extract_transitions(mz, g["m7_traj"], g["m7_rew"], None)
    cells = np.array(mz.vis_cells)
    assert np.array_equal(cells[tr["s"]], g["m7_tr_s"])
    assert np.array_equal(cells[tr["s1"]], g["m7_tr_s1"])
    n = len(tr["s"])
    for p, (a, gm) in enumerate(g["params"]):
        V, _ = O.run_session(O.ALG_TD0, tr, a, gm, len(cells), orders=O.replay_orders(n, 2, 21))
        assert np.array_equal(O.dense_table(mz, V), g["m7_td0_rep2_V"][p])
    a7 = O.A8_UNITS * 0.2
    for alg in ("positive", "negative"):
        rew = -g["m7b_rew"] if alg == "negative" else g["m7b_rew"]
        trb = O.extract_transitions(mz, g["m7b_traj"], rew, O.OracleMouse(g["m7b_traj"][0], 0.0, a7))
        for p, (a, gm) in enumerate(g["params"]):
            V, _ = O.run_session(ALGS[alg], trb, a, gm, len(cells), orders=O.replay_orders(n, 2, 21))
            assert np.array_equal(O.dense_table(mz, V), g[f"m7b_{alg}_rep2_V"][p])


def test_benchmark_session_known_answers(golden_fit):
    """SURVEY 8c KAT: session 0, T=5000, TD0(0.1, 0.9), no replays."""
    g = golden_fit
    mz = H.oracle_m9()
    traj, rew = O.synthetic_session(mz, 0, 5000)
    assert np.array_equal(traj[:16], g["kat_traj_head"])
    assert int(rew.sum()) == int(g["kat_rewarded_steps"]) == 108
    tr = H.oracle_transitions(mz, traj, rew, H.a8())
    V, _ = O.run_session(O.ALG_TD0, tr, 0.1, 0.9, 60)
    D = O.dense_table(mz, V)
    assert np.array_equal(D, g["kat_td0_V"])
    assert D.max() == 0.5949675648481678 and D[7, 4] == 0.48326347528794783
    Vp, _ = CO.fit(O.ALG_POSITIVE, tr, [0.1], [0.9], 60, orders=O.replay_orders(4999, 1, 0))
    assert np.array_equal(O.dense_table(mz, Vp[0]), g["kat_positive_rep1_V"])


def test_loglik_c_vs_numpy_oracle():
    """The two restatements of the (unpinned) log-likelihood extension agree."""
    mz = H.oracle_m9()
    traj, rew = O.synthetic_session(mz, 2, 700)
    tr = H.oracle_transitions(mz, traj, rew, H.a8())
    rng = np.random.default_rng(0)
    a, b, gm = H.grid_params(rng, 6)
    for alg in ALGS.values():
        Vc, llc = CO.fit(alg, tr, a, gm, 60, beta=b)
        for p in range(len(a)):
            V, ll = O.run_session(alg, tr, a[p], gm[p], 60, beta=b[p])
            assert np.array_equal(V, Vc[p])
            assert abs(ll - llc[p]) <= 1e-12 * abs(ll)


def test_tile_action_agent_oracles_closed_forms():
    """The Q-table and model-based agents are extensions without a reference implementation: their numpy oracles are
    at least held to closed forms.  alpha = 0 never moves the tables, so every observed step has probability 1 / K;
    beta = 0 gives the same likelihood whatever was learned; and on the deterministic maze the Q-table learner's
    max-backup equals PositiveConditioning's on after-states: Q(s, a) -> r + gamma * max Q(next(s, a), .)."""
    mz = H.oracle_m9()
    acts = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    nxt = O.transition_table(mz, acts)
    traj, rew = O.synthetic_session(mz, 3, 600)
    tr = O.extract_transitions(mz, traj, rew, None)
    closed = 0.0
    for s, s1 in zip(tr["s"].tolist(), tr["s1"].tolist()):
        avail = [a for a in range(len(acts)) if nxt[s, a] >= 0]
        if any(nxt[s, a] == s1 for a in avail):
            closed = closed + (0.0 - np.log(float(len(avail))))
    Q, ll = O.run_q_learning(tr, nxt, 0.0, 3.0, 0.9)
    assert not Q.any() and abs(ll - closed) <= 1e-12 * abs(closed)
    Q, ll = O.run_q_learning(tr, nxt, 0.4, 0.0, 0.9)
    assert Q.any() and abs(ll - closed) <= 1e-12 * abs(closed)
    V, R, ll = O.run_model_based(tr, nxt, 0.0, 3.0, 0.9)
    assert not V.any() and not R.any() and abs(ll - closed) <= 1e-12 * abs(closed)
    # replays only ever update: the likelihood is that of the online pass
    orders = O.replay_orders(len(tr["s"]), 2, 1)
    Q2, ll2 = O.run_q_learning(tr, nxt, 0.3, 2.0, 0.8, orders=orders)
    Q1, ll1 = O.run_q_learning(tr, nxt, 0.3, 2.0, 0.8)
    assert ll1 == ll2 and not np.array_equal(Q1, Q2)
    # fixed point of the backup with alpha = 1 on a two-step stream: Q(s, a) = r + gamma * max_a' Q(s1, a')
    two = {k: v[:2] for k, v in tr.items()}
    Q, _ = O.run_q_learning(two, nxt, 1.0, 1.0, 0.5, q0=np.ones_like(nxt, dtype=np.float64))
    s, s1, r = int(two["s"][0]), int(two["s1"][0]), float(two["r"][0])
    a = next(a for a in range(len(acts)) if nxt[s, a] == s1)
    assert Q[s, a] == r + 0.5 * 1.0 or (s1 == s)  # the bootstrap row was still all ones


def test_floor_division_boundaries():
    """Python float floor-division differs from floor(x / w) on tile multiples."""
    mz = H.oracle_m9()
    assert mz.cont2disc(0.999, 0.443) == (8, 3)
    assert int(np.floor(0.999 / 0.111)) == 9
