"""The oracles' log-likelihood against values EXECUTED BY THE REFERENCE (tests/golden/ll_golden.npz).

make_golden_ll.py drives the unmodified reference: tables through ConditioningExperiment.run and the learners'
update (rl.py:88-102, 176-194, 219-237, 256-262), probabilities through BehaviouralModel.decision_making
(exploration_model.py:84-93) with a table-lookup component, log(p[observed]) per step.  This pins the numpy and C
restatements of the likelihood (they were "parity unpinned" in round 1): per-step log-probabilities to 1e-12
(the oracles take the logarithm of the softmax denominator instead of the quotient's, a rounding-level
difference), per-session sums to 1e-12 relative, final tables bit for bit."""
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O

import helpers as H
from conftest import GOLDEN

ALGS = {"td0": O.ALG_TD0, "positive": O.ALG_POSITIVE, "negative": O.ALG_NEGATIVE}
RTOL = 1e-12


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "ll_golden.npz"), allow_pickle=False))


def _transitions(g, j, alg):
    mz = H.oracle_m9()
    rew = -g[f"s{j}_rew"] if alg == "negative" else g[f"s{j}_rew"]
    return mz, H.oracle_transitions(mz, g[f"s{j}_traj"], rew, g["A8"])


@pytest.mark.parametrize("j", [0, 1])
def test_observed_candidate_index_matches_reference(g, j):
    mz, tr = _transitions(g, j, "positive")
    obs = tr["obs"].astype(np.int64)
    obs[obs == O.OBS_NONE] = -1
    assert np.array_equal(obs, g[f"s{j}_obs"])


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
@pytest.mark.parametrize("j", [0, 1])
def test_per_step_logp_and_session_sum(g, alg, j):
    mz, tr = _transitions(g, j, alg)
    n = len(tr["s"])
    for p, (a, b, gm) in enumerate(g["params"]):
        V = [0.0] * len(mz.vis_cells)
        want = g[f"{alg}_s{j}_logp"][p]
        for i in range(n):
            cand = tr["cand"][i, :tr["ncand"][i]].tolist()
            if tr["obs"][i] != O.OBS_NONE:
                lp = O.step_logp(V, cand, int(tr["obs"][i]), b)
                assert abs(lp - want[i]) <= RTOL * max(1.0, abs(want[i])), (alg, j, p, i)
            else:
                assert np.isnan(want[i])
            O.step_update(ALGS[alg], V, int(tr["s"][i]), int(tr["s1"][i]), float(tr["r"][i]), cand, a, gm)
        Vn, ll = O.run_session(ALGS[alg], tr, a, gm, len(mz.vis_cells), beta=b)
        assert abs(ll - g[f"{alg}_s{j}_ll"][p]) <= RTOL * abs(g[f"{alg}_s{j}_ll"][p])
        assert np.array_equal(O.dense_table(mz, Vn), g[f"{alg}_s{j}_V"][p])


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
@pytest.mark.parametrize("j", [0, 1])
def test_c_oracle(g, alg, j):
    mz, tr = _transitions(g, j, alg)
    a, b, gm = g["params"].T
    V, ll = CO.fit(ALGS[alg], tr, a, gm, len(mz.vis_cells), beta=b)
    np.testing.assert_allclose(ll, g[f"{alg}_s{j}_ll"], rtol=RTOL, atol=0)
    for p in range(len(a)):
        assert np.array_equal(O.dense_table(mz, V[p]), g[f"{alg}_s{j}_V"][p])


def test_replays_leave_the_likelihood_alone(g):
    mz, tr = _transitions(g, 0, "positive")
    n = len(tr["s"])
    a, b, gm = g["params"].T
    V, ll = CO.fit(O.ALG_POSITIVE, tr, a, gm, len(mz.vis_cells), beta=b, orders=O.replay_orders(n, 2, 17))
    np.testing.assert_allclose(ll, g["rep_positive_s0_ll"], rtol=RTOL, atol=0)
    for p in range(len(a)):
        assert np.array_equal(O.dense_table(mz, V[p]), g["rep_positive_s0_V"][p])


@pytest.mark.parametrize("alg", ["td0", "positive"])
def test_given_start_table(g, alg):
    mz, tr = _transitions(g, 0, alg)
    v0 = H.compact(mz, g["v0"])
    for p, (a, b, gm) in enumerate(g["params"]):
        V, ll = O.run_session(ALGS[alg], tr, a, gm, len(v0), beta=b, v0=v0)
        assert abs(ll - g[f"v0_{alg}_s0_ll"][p]) <= RTOL * abs(g[f"v0_{alg}_s0_ll"][p])
        assert np.array_equal(O.dense_table(mz, V, g["v0"]), g[f"v0_{alg}_s0_V"][p])
    a, b, gm = g["params"].T
    Vc, llc = CO.fit(ALGS[alg], tr, a, gm, len(v0), beta=b, v0=v0)
    np.testing.assert_allclose(llc, g[f"v0_{alg}_s0_ll"], rtol=RTOL, atol=0)
