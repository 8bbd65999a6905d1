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


def test_floor_division_boundaries():
    """Python float floor-division differs from floor(x / w) on tile multiples."""
    mz = H.oracle_m9()
    assert mz.cont2disc(0.999, 0.443) == (8, 3)
    assert int(np.floor(0.999 / 0.111)) == 9
