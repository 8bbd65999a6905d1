"""Golden LOG-LIKELIHOOD fixtures whose every arithmetic operation is executed by the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_ll.py

The reference has no likelihood function, but both halves of one are reference code:

* the value table evolves through ``ConditioningExperiment.run`` -> ``TD0 / PositiveConditioning /
  NegativeConditioning.update`` (``simulation/rl.py:88-102, 176-194, 219-237, 256-262``), replays through
  ``ShuffledReplays.offline_replays`` (``rl.py:419-424``);
* the choice probabilities are ``BehaviouralModel.decision_making`` (``simulation/exploration_model.py:68-98``): a
  ONE-LINE table-lookup ``Component`` (built like ``test/test_exploration.py:25-36``'s ``FixedComponent``) with
  weight 1 hands ``V[candidates]`` to the model, which forms ``beta * val``, subtracts the maximum, exponentiates
  and normalises with ``np.sum`` (``:84-93``).  The probability vector is read from the model's own ``choice(p=...)``
  call (``:98``) through a recording generator proxy (the trick of make_golden_sim.py).

So, right before every ONLINE update of the learner, this script asks the reference model for the probabilities of
the step's candidates under the table as it is then, and stores ``log(p[observed])`` (``np.log``) per step and their
step-order sum per session.  The candidates of a step are the ``next_states`` the reference's own
``PositiveConditioning`` caches for that transition (``rl.py:177-182``; parameter-independent, taken from a first
pass), the observed one is the first candidate equal to the arrival tile ``t.s1``; steps whose arrival tile is not a
candidate contribute nothing.  Nothing from this repo's product code or oracle is used.

ll_golden.npz (a few hundred KB):
  params[4,3] = (alpha, beta, gamma); for alg in td0 / positive / negative and session j in 0, 1:
  ``{alg}_s{j}_logp[P, T-1]`` (NaN where a step is skipped), ``{alg}_s{j}_ll[P]``, ``{alg}_s{j}_V[P, X, Y]`` final tables;
  ``s{j}_traj / _rew / _obs`` (observed candidate index per step, -1 = skipped); plus
  ``rep_*``  PositiveConditioning with 2 shuffled replay passes (the likelihood is online-only, the table is not) and
  ``v0_*``   TD0 / PositiveConditioning continuing from a given table (rl.py:63-64).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refharness  # noqa: E402

if not refharness.install():
    raise SystemExit("reference not available at /root/reference")

from navigation_model.simulation.behavioural_components import Component, weight  # noqa: E402  (the reference)
from navigation_model.simulation.exploration_model import BehaviouralModel  # noqa: E402
from navigation_model.simulation.maze import Maze  # noqa: E402
from navigation_model.simulation.mouse import Mouse  # noqa: E402
from navigation_model.simulation.rl import (ConditioningExperiment, NegativeConditioning,  # noqa: E402
                                            PositiveConditioning, ShuffledReplays, TD0)

from make_golden import A8, M9_FILE, SIZE_TILE, synthetic_session  # noqa: E402


@weight("W_table")
class TableLookup(Component):
    """val_k = V[candidate k] of the learner's CURRENT table (weight defaults to 1: ``x * 1.0`` is exact)."""
    _name = "table"

    def __init__(self, alg, **kwargs):
        super().__init__(**kwargs)
        self._alg = alg

    def update(self, next_positions, **kwargs):
        self._val = np.array([self._alg.v_table[c[0], c[1]] for c in next_positions])


class RecordingRng:
    """Wraps the model's numpy Generator; keeps the probability vector of the last choice(p=...) call."""

    def __init__(self, rng):
        self._rng = rng
        self.last_p = None

    def choice(self, a, p=None):
        self.last_p = np.array(p)
        return self._rng.choice(a, p=p)


def candidates_of_session(maze, traj, rew):
    """First pass: the next_states lists the reference's PositiveConditioning caches per transition."""
    mouse = Mouse(traj[0], 0.0, A8)
    rep = ShuffledReplays(n_times=0, seed=None)
    exp = ConditioningExperiment(maze, PositiveConditioning(0.1, 0.9, mouse, maze), rep)
    exp.run(list(traj), [0.0] * len(traj), list(rew), do_replays=False)
    return [[(int(c[0]), int(c[1])) for c in u.transition.extra["next_states"]] for u in rep._buffer]


def reference_loglik(maze, alg_name, alpha, beta, gamma, traj, rew, cands, n_times=0, seed=None, v_init=None):
    mouse = Mouse(traj[0], 0.0, A8)
    alg = {"td0": lambda: TD0(alpha, gamma), "positive": lambda: PositiveConditioning(alpha, gamma, mouse, maze),
           "negative": lambda: NegativeConditioning(alpha, gamma, mouse, maze)}[alg_name]()
    rep = ShuffledReplays(n_times=n_times, seed=seed)
    exp = ConditioningExperiment(maze, alg, rep, v_table=None if v_init is None else v_init.copy())
    model = BehaviouralModel([TableLookup(alg, W_table=1.0)], (0, 0), beta=beta, seed=0)
    rec = RecordingRng(model._rng)
    model._rng = rec
    n = len(traj) - 1
    logp = np.full(n, np.nan)
    obs = np.full(n, -1, dtype=np.int64)
    state = {"i": 0, "ll": 0.0, "min_p": 1.0}
    plain_update = alg.update

    def update_with_likelihood(t):
        i = state["i"]
        if i < n:  # online pass only; replays (i >= n) update the table and nothing else
            cand = cands[i]
            s1 = (int(t.s1[0]), int(t.s1[1]))
            if s1 in cand:
                k = cand.index(s1)
                model.decision_making(next_positions_cont=np.array(cand, dtype=np.float64), next_positions=cand)
                p = rec.last_p
                assert len(p) == len(cand)
                logp[i] = np.log(p[k])
                obs[i] = k
                state["ll"] = state["ll"] + logp[i]
                state["min_p"] = min(state["min_p"], p[k])
        state["i"] = i + 1
        return plain_update(t)

    alg.update = update_with_likelihood
    exp.run(list(traj), [0.0] * len(traj), list(rew), do_replays=n_times > 0)
    assert state["i"] == n * (1 + n_times)
    assert state["min_p"] > 1e-250, "a probability underflowed: pick milder parameters"
    return logp, obs, state["ll"], np.array(exp.v_table)


def main():
    maze = Maze.from_file(M9_FILE, size_tile=SIZE_TILE)
    params = np.array([(0.1, 2.0, 0.9), (0.5, 0.7, 0.5), (0.9, 12.0, 0.97), (0.013, 31.0, 0.731)])
    out = {"params": params, "A8": A8}
    sessions = [synthetic_session(maze, 12, 700), synthetic_session(maze, 13, 500, start=(7, 7), goal=(1, 2))]
    for j, (traj, rew) in enumerate(sessions):
        cands = candidates_of_session(maze, traj, rew)
        out[f"s{j}_traj"], out[f"s{j}_rew"] = traj, rew
        for alg in ("td0", "positive", "negative"):
            r = -rew if alg == "negative" else rew
            lps, lls, Vs = [], [], []
            for (a, b, g) in params:
                logp, obs, ll, V = reference_loglik(maze, alg, a, b, g, traj, r, cands)
                lps.append(logp), lls.append(ll), Vs.append(V)
            out[f"{alg}_s{j}_logp"], out[f"{alg}_s{j}_ll"], out[f"{alg}_s{j}_V"] = np.array(lps), np.array(lls), np.array(Vs)
        out[f"s{j}_obs"] = obs
        print(f"session {j}: {int((obs >= 0).sum())} of {len(obs)} steps observed; ll(positive) = {out[f'positive_s{j}_ll']}")
    # replay passes: likelihood online-only, table after 2 shuffled passes
    traj, rew = sessions[0]
    cands = candidates_of_session(maze, traj, rew)
    lls, Vs = [], []
    for (a, b, g) in params:
        _, _, ll, V = reference_loglik(maze, "positive", a, b, g, traj, rew, cands, n_times=2, seed=17)
        lls.append(ll), Vs.append(V)
    out["rep_positive_s0_ll"], out["rep_positive_s0_V"] = np.array(lls), np.array(Vs)
    assert np.array_equal(out["rep_positive_s0_ll"], out["positive_s0_ll"])
    # continued learning from a given table
    v0 = np.random.default_rng(9).random(maze.shape) * 0.5
    out["v0"] = v0
    for alg in ("td0", "positive"):
        lls, Vs, lps = [], [], []
        for (a, b, g) in params:
            logp, _, ll, V = reference_loglik(maze, alg, a, b, g, traj, rew, cands, v_init=v0)
            lls.append(ll), Vs.append(V), lps.append(logp)
        out[f"v0_{alg}_s0_ll"], out[f"v0_{alg}_s0_V"], out[f"v0_{alg}_s0_logp"] = np.array(lls), np.array(Vs), np.array(lps)
    np.savez_compressed(os.path.join(HERE, "ll_golden.npz"), **out)
    print("ll_golden.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
