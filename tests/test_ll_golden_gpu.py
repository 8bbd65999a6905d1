"""The CUDA likelihood kernels against log-likelihoods EXECUTED BY THE REFERENCE (tests/golden/ll_golden.npz,
written by tests/golden/make_golden_ll.py: rl.py updates + exploration_model.py:84-93 softmax).  Every one of the
three fit kernels -- a table per warp (runs of 32 sets sharing alpha and gamma), a table per set in device memory, a
table per set in shared memory -- must give the reference's per-session sums to 1e-9 relative (north_star) and the
reference's final tables bit for bit."""
import os

import numpy as np
import pytest

import helpers as H
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "ll_golden.npz"), allow_pickle=False))


def _sweep(g, alg, js=(0, 1), **kw):
    from navigation_model_b200.sweep import ConditioningSweep
    sw = ConditioningSweep(H.product_m9(), alg, possible_actions=g["A8"], **kw)
    for j in js:
        rew = -g[f"s{j}_rew"] if alg == "negative" else g[f"s{j}_rew"]
        sw.add_session(g[f"s{j}_traj"], rew)
    return sw


def _grid_with(params, n_beta=32):
    """Every golden (alpha, beta, gamma) embedded in a run of n_beta sets sharing (alpha, gamma): the golden beta sits at
    a different lane in each run, the others are log-spaced fillers."""
    a, b, gm, where = [], [], [], []
    for i, (ai, bi, gi) in enumerate(params):
        betas = np.logspace(-1, 1.4, n_beta)
        lane = (7 * i + 3) % n_beta
        betas[lane] = bi
        a.append(np.full(n_beta, ai)), b.append(betas), gm.append(np.full(n_beta, gi))
        where.append(i * n_beta + lane)
    return np.concatenate(a), np.concatenate(b), np.concatenate(gm), np.array(where)


@pytest.mark.parametrize("alg", ["td0", "positive", "negative"])
def test_three_kernels_match_reference_executed_likelihood(cuda, g, alg, monkeypatch):
    sw = _sweep(g, alg)
    want = np.stack([g[f"{alg}_s{j}_ll"] for j in (0, 1)])
    wantV = np.stack([g[f"{alg}_s{j}_V"] for j in (0, 1)])
    # (1) a table per warp: beta-fastest runs of 32
    a, b, gm, where = _grid_with(g["params"])
    ll, V = sw.log_likelihood(a, b, gm, return_v=True)
    assert sw.last_sweep_mode() == 1
    np.testing.assert_allclose(ll[:, where], want, rtol=LL_RTOL, atol=0)
    assert np.array_equal(V[:, where], wantV)
    # (2) a table per set in device memory: the four golden sets as they are
    a, b, gm = g["params"].T
    ll, V = sw.log_likelihood(a, b, gm, return_v=True)
    assert sw.last_sweep_mode() == 2
    np.testing.assert_allclose(ll, want, rtol=LL_RTOL, atol=0)
    assert np.array_equal(V, wantV)
    # (3) a table per set in shared memory
    monkeypatch.setenv("NM_FIT_TABLES", "shared")
    ll, V = sw.log_likelihood(a, b, gm, return_v=True)
    assert sw.last_sweep_mode() == 0
    np.testing.assert_allclose(ll, want, rtol=LL_RTOL, atol=0)
    assert np.array_equal(V, wantV)


def test_replays_and_given_table(cuda, g):
    a, b, gm = g["params"].T
    sw = _sweep(g, "positive", js=(0,), replay_n_times=2, replay_seed=17)
    ll, V = sw.log_likelihood(a, b, gm, return_v=True)
    np.testing.assert_allclose(ll[0], g["rep_positive_s0_ll"], rtol=LL_RTOL, atol=0)
    assert np.array_equal(V[0], g["rep_positive_s0_V"])
    for alg in ("td0", "positive"):
        sw = _sweep(g, alg, js=(0,))
        ll, V = sw.log_likelihood(a, b, gm, v_init=g["v0"], return_v=True)
        np.testing.assert_allclose(ll[0], g[f"v0_{alg}_s0_ll"], rtol=LL_RTOL, atol=0)
        assert np.array_equal(V[0], g[f"v0_{alg}_s0_V"])
        ag, bg, gg, where = _grid_with(g["params"])
        ll = sw.log_likelihood(ag, bg, gg, v_init=g["v0"])
        np.testing.assert_allclose(ll[0, where], g[f"v0_{alg}_s0_ll"], rtol=LL_RTOL, atol=0)


def test_scalar_experiment_plus_fitness_objective(cuda, g):
    """The drop-in classes: LogLikelihoodFitness over one session gives the reference-executed value."""
    from navigation_model_b200.optimization.fitness import LogLikelihoodFitness
    fit = LogLikelihoodFitness(H.product_m9(), [(g["s0_traj"], g["s0_rew"])], algorithm="positive",
                               possible_actions=g["A8"])
    for p, (a, b, gm) in enumerate(g["params"]):
        nll = fit(a, b, gm)
        assert abs(-np.ravel(nll)[0] - g["positive_s0_ll"][p]) <= LL_RTOL * abs(g["positive_s0_ll"][p])
