"""Fitness objectives: the reference's ``MultiObjectiveFitness`` plus the log-likelihood objective of the
parameter-grid sweep.

``MultiObjectiveFitness`` mirrors ``optimization/fitness.py:4-117`` (named objectives, each a
``ResultsDistance`` against fixed data metrics; ``RuntimeError`` on a duplicate name).
``LogLikelihoodFitness`` is the extension BASELINE.json's north_star asks for: the negative
log-likelihood of recorded sessions under a value learner + softmax policy, evaluated for whole
parameter grids on the GPU (:class:`navigation_model_b200.sweep.ConditioningSweep`).  The reference has
no likelihood: its semantics are specified in DESIGN.md and checked against this repo's own CPU
restatement only ("parity unpinned").
"""
import numpy as np

from ..analysis.statistics import ResultsDistance


class MultiObjectiveFitness:
    """A fitness with many named objectives (``fitness.py:4-117``)."""

    def __init__(self, compare_against):
        self._compare_against = compare_against
        self._distances = {}

    def add_objective(self, name, distance_functions, *keys):
        if name in self._distances:
            raise RuntimeError(f"Objective {name} already in fitness")
        self._distances[name] = ResultsDistance(distance_functions, self._compare_against, *keys)

    def __call__(self, results):
        return {name: dist(results) for name, dist in self._distances.items()}

    def reset(self):
        for dist in self._distances.values():
            dist.reset()

    def _collect(self, what):
        return {name: getattr(dist, what)() for name, dist in self._distances.items()}

    def medians(self):
        return self._collect("median")

    def means(self):
        return self._collect("mean")

    def stds(self):
        return self._collect("std")

    def percentiles(self):
        return self._collect("percentile")

    def data(self):
        return self._collect("data")

    def medians_list(self):
        return list(self.medians().values())

    def means_list(self):
        return list(self.means().values())

    def stds_list(self):
        return list(self.stds().values())

    def percentiles_list(self):
        return list(self.percentiles().values())

    def data_list(self):
        return list(self.data().values())


class LogLikelihoodFitness:
    """Negative log-likelihood of recorded sessions for grids of ``(alpha, beta, gamma)``.

    ``algorithm``: ``"td0"`` / ``"positive"`` / ``"negative"`` (the reference's value learners, with
    ``possible_actions`` = the mouse's relative actions), or one of the two tile-action agents (extensions) --
    ``"qtable"`` (explicit ``Q[s, a]`` table) and ``"modelbased"`` (reward model + ``n_sweeps`` value-iteration sweeps
    per step) -- with ``discrete_actions`` = tile displacements.

    >>> fit = LogLikelihoodFitness(maze, sessions, algorithm="positive", possible_actions=A8)
    >>> nll = fit(alpha, beta, gamma)                  # float64[P], summed over sessions
    >>> best_nll, best_index = fit.best(alpha, beta, gamma)   # sharded over the process group, if any
    """

    def __init__(self, maze, sessions, algorithm="positive", possible_actions=None, mouse_init=None,
                 replay_n_times=0, replay_seed=None, device=None, discrete_actions=None, n_sweeps=1):
        from ..sweep import ConditioningSweep, ModelBasedSweep, QLearningSweep
        name = algorithm.lower() if isinstance(algorithm, str) else None
        if name in ("qtable", "modelbased"):
            if discrete_actions is None:
                raise ValueError(f"algorithm {algorithm!r} needs discrete_actions (tile displacements)")
            if name == "qtable":
                self._sweep = QLearningSweep(maze, discrete_actions, device=device, replay_n_times=replay_n_times,
                                             replay_seed=replay_seed)
            else:
                if replay_n_times:
                    raise ValueError("the model-based agent plans instead of replaying: replay_n_times must be 0")
                self._sweep = ModelBasedSweep(maze, discrete_actions, n_sweeps=n_sweeps, device=device)
        else:
            self._sweep = ConditioningSweep(maze, algorithm, possible_actions=possible_actions, mouse_init=mouse_init,
                                            replay_n_times=replay_n_times, replay_seed=replay_seed, device=device)
        self._tile_actions = name in ("qtable", "modelbased")
        for s in sessions:
            if isinstance(s, (tuple, list)) and len(s) == 2:
                self._sweep.add_session(s[0], s[1])
            else:
                if not s.has_reward:
                    raise RuntimeError("LogLikelihoodFitness: the session has no reward")
                self._sweep.add_session(s.trajectory, s.reward)

    @property
    def sweep(self):
        return self._sweep

    def per_session(self, alpha, beta, gamma, v_init=None):
        """``float64[S, P]`` log-likelihoods (``v_init``: the initial value table of the value learners / the
        model-based agent)."""
        if isinstance(self._sweep, _sweep_types()[2]):
            if v_init is not None:
                raise ValueError("the Q-table agent starts from q_init, not v_init: use fit.sweep.log_likelihood")
            return self._sweep.log_likelihood(alpha, beta, gamma)
        return self._sweep.log_likelihood(alpha, beta, gamma, v_init=v_init)

    def __call__(self, alpha, beta, gamma, v_init=None):
        ll = self.per_session(alpha, beta, gamma, v_init=v_init)
        acc = np.zeros(ll.shape[1])
        for row in ll:  # sessions added in index order, like the device epilogue
            acc = acc + row
        return -acc

    def best(self, alpha, beta, gamma, v_init=None, group=None):
        if self._tile_actions:
            if v_init is not None:
                raise ValueError("best() of the tile-action agents starts from zero tables")
            return self._sweep.best_fit(alpha, beta, gamma, group=group)
        return self._sweep.best_fit(alpha, beta, gamma, v_init=v_init, group=group)


def _sweep_types():
    from ..sweep import ConditioningSweep, ModelBasedSweep, QLearningSweep
    return ConditioningSweep, ModelBasedSweep, QLearningSweep
