"""Exploration models: how the next position is chosen among the reachable endpoints.

Host-side mirror of the reference's ``simulation/exploration_model.py`` (``:4-144``): same classes,
constructor arguments and ``decision_making(**kwargs)`` plugin call.  ``decision_making`` is the scalar,
per-step interface (one softmax + one draw) that user code and the reference's tests drive; whole
simulations do not loop over it -- ``StandardExperiment.run`` / :mod:`navigation_model_b200.batched_sim`
evaluate the same arithmetic inside the CUDA kernel for every step and every mouse.
"""
import numpy as np


class ExplorationModel:
    """Interface (``:4-21``)."""

    @property
    def parameters(self):
        raise NotImplementedError()

    def decision_making(self, **kwargs):
        raise NotImplementedError()


class BehaviouralModel(ExplorationModel):
    """Softmax policy over the summed, weighted component values (``:24-109``).

    ``p = softmax(beta * sum_c w_c * val_c)`` over the currently reachable endpoints; one
    ``Generator.choice(next_positions_cont, p=p)`` draw per decision.  ``_moving`` is ``True`` when the
    previous choice equalled the position before it (i.e. the agent *stayed*) -- the flag is
    passed on to the components unchanged.
    """

    def __init__(self, components, init_pos_discrete, beta=None, seed=None, debug=False):
        self._components = components
        self._prev_pos = init_pos_discrete
        self._moving = False
        self._beta = beta
        self._rng = np.random.default_rng(seed)
        self._model_parameters = {"beta": beta}
        for comp in self._components:
            self._model_parameters.update(comp.parameters)
        self._debug = debug

    @property
    def parameters(self):
        return self._model_parameters

    def decision_making(self, next_positions_cont, **kwargs):
        for comp in self._components:
            comp.update(moving=self._moving, **kwargs)
        total = np.zeros(len(next_positions_cont))
        for comp in self._components:
            total = total + comp.weighted_values
        total = self._beta * total
        total = total - np.max(total)
        weights = np.exp(total)
        p = weights / np.sum(np.exp(total))
        new_pos = self._rng.choice(next_positions_cont, p=p)
        self._moving = np.all(self._prev_pos == new_pos)
        self._prev_pos = new_pos
        return new_pos


class RandomModel(ExplorationModel):
    """Uniform choice among the reachable endpoints (``:112-144``)."""

    def __init__(self, seed=None, debug=False):
        self._rng = np.random.default_rng(seed)
        self._debug = debug

    @property
    def parameters(self):
        return {}

    def decision_making(self, next_positions_cont, **kwargs):
        return self._rng.choice(next_positions_cont)
