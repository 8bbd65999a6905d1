"""Forward simulation of one agent: the reference's ``StandardExperiment`` on top of the CUDA kernel.

Mirror of ``simulation/experiment.py`` (``:1-51``; the matplotlib ``plot`` helper ``:53-69`` is out of
scope).  ``run(iterations)`` with a ``BehaviouralModel`` made of the built-in components does **not**
loop in Python: the whole run is one launch of ``nm_simulate`` with a single mouse whose generator state
is taken from -- and written back to -- the model's numpy ``Generator``; afterwards the mouse (pose +
history), the model (``_prev_pos``, ``_moving``) and the ``ExperienceComponent`` (visit map, familiarity,
mental state) hold exactly what the per-step loop would have left -- and the next ``run`` on the same objects
starts its launch from that state (seeded visit map, familiarity, mental state, ``_moving``), as the reference's
loop would.  That path needs a CUDA device and fails loudly without one.  Models with user-defined components / exploration models are host plugins by
nature and run through the generic per-step loop.
"""
import numpy as np

from .. import _native
from . import behavioural_components as bc
from .exploration_model import BehaviouralModel, RandomModel
from .maze import Maze
from .mouse import Mouse

_BUILTIN = (bc.AnxietyComponent, bc.AnxietyComponent_Grid, bc.BiomechanicalCostComponent, bc.ExperienceComponent,
            bc.BiomechPersistenceComponent, bc.ValueTableComponent)


class StandardExperiment:
    """An exploration model, a maze and a mouse (``experiment.py:1-51``)."""

    def __init__(self, exp_model, maze, mouse, debug=False):
        self._exp_model = exp_model
        self._maze = maze
        self._mouse = mouse
        self._debug = debug

    def _device_eligible(self, iterations=0):
        model = self._exp_model
        if type(self._maze) is not Maze or type(self._mouse) is not Mouse:
            return False
        X, Y = self._maze.shape
        seeded = 0  # visits an ExperienceComponent already holds (a second run on the same objects)
        if type(model) is BehaviouralModel:
            for c in model._components:
                if isinstance(c, bc.ExperienceComponent):
                    m = np.asarray(c._maze_map_exp)
                    if (m.shape != (X, Y) or not np.isfinite(m).all() or (m < 0).any() or (m != np.floor(m)).any()
                            or m.max() + iterations >= 2 ** 32 - 1):
                        return False  # not a visit count the kernel's integer map can carry: host loop
                    seeded = int(m.max())
        if type(model) in (BehaviouralModel, RandomModel) and \
                not _native.lib().nm_simulate_fits(int(X), int(Y), int(iterations) + seeded):
            # beyond ~66 000 tiles not even the tile tables fit shared memory; smaller mazes whose VISIT MAP does not fit
            # it (beyond ~50 x 50 tiles) run on the device with the map in device memory
            raise RuntimeError(f"a maze of {X} x {Y} tiles is too large for the simulation kernel; there is no CPU "
                               "fallback for the built-in models (a user-defined exploration model takes the per-step "
                               "plugin loop)")
        if type(model) is RandomModel:
            return (1 <= len(self._mouse.possible_actions) <= _native.NM_MAX_CAND
                    and type(model._rng.bit_generator).__name__ == "PCG64")
        if type(model) is not BehaviouralModel:
            return False
        comps = model._components
        if any(type(c) not in _BUILTIN for c in comps):
            return False
        kinds = [c.device_kind for c in comps]
        if len(set(kinds)) != len(kinds) or len(kinds) > _native.NM_MAX_COMPONENTS:
            return False
        n_act = len(self._mouse.possible_actions)
        if not 1 <= n_act <= _native.NM_MAX_CAND:
            return False
        for c in comps:
            if isinstance(c, (bc.BiomechanicalCostComponent, bc.BiomechPersistenceComponent)):
                if len(c._discrete_actions) != n_act:
                    return False
        return type(model._rng.bit_generator).__name__ == "PCG64"

    def run(self, iterations):
        self._mouse.enable_history(True)  # experiment.py:32
        if self._device_eligible(int(iterations)):
            self._run_device(int(iterations))
        else:
            self._run_host(int(iterations))

    # -- generic per-step loop (host plugins) -----------------------------------------------------------
    def _run_host(self, iterations):
        maze, mouse, model = self._maze, self._mouse, self._exp_model
        for _ in range(iterations):
            ends = mouse.get_endpoints()
            reachable = maze.are_movements_possible_cont(mouse.position, ends)
            ends = ends[reachable]
            cells = maze.cont2disc_list(ends)
            x, y = maze.cont2disc(mouse.position)
            new_position = model.decision_making(next_positions=cells, next_tiles=maze.get_tiles(cells), x=x, y=y,
                                                 is_visitable=reachable, next_positions_cont=ends)
            mouse.move_to(new_position)

    # -- device path ----------------------------------------------------------------------------------------
    def _run_device_random(self, iterations):
        """RandomModel: uniform ``rng.choice`` among the reachable endpoints (``exploration_model.py:131-144``),
        continued on the device from the model's generator (incl. numpy's buffered 32-bit half)."""
        from ..batched_sim import BatchedExperiment, set_pcg64_state
        model, mouse, maze = self._exp_model, self._mouse, self._maze
        actions = mouse.possible_actions
        exp = BatchedExperiment(maze, actions, np.ones((len(actions), 2)), [], model="random")
        res = exp.run({}, iterations, mouse.position, mouse.orientation, generators=[model._rng], history=True)
        if res.status[0] != 0:  # rng.choice of an empty array
            raise ValueError("a must be non-empty")
        if iterations <= 0:
            return
        hp, ho = res.history_position[0], res.history_orientation[0]
        _native.check(_native.lib().nm_mouse_adopt_trajectory(
            mouse._h, _native.np_ptr(np.ascontiguousarray(hp[1:])), _native.np_ptr(np.ascontiguousarray(ho[1:])),
            iterations))
        set_pcg64_state(model._rng, res.pcg_state[0], res.pcg_buffered[0])

    def _run_device(self, iterations):
        from ..batched_sim import BatchedExperiment, set_pcg64_state
        model, mouse, maze = self._exp_model, self._mouse, self._maze
        if type(model) is RandomModel:
            return self._run_device_random(iterations)
        comps = model._components
        actions = mouse.possible_actions
        disc_actions, v_norm = None, None
        for c in comps:
            if isinstance(c, (bc.BiomechanicalCostComponent, bc.BiomechPersistenceComponent)):
                disc_actions = c._discrete_actions
            if isinstance(c, bc.ValueTableComponent):
                v_norm = c._v_table
        if disc_actions is None:
            disc_actions = np.ones((len(actions), 2))  # no component looks at them
        exp = BatchedExperiment(maze, actions, disc_actions, comps, v_table=v_norm, v_table_normalised=True)
        params = {"beta": model._beta}
        for c in comps:
            params.update(c.parameters)
        init_disc = np.asarray(model._prev_pos, dtype=np.float64).reshape(2)
        # the model and the ExperienceComponent carry on where an earlier run() left them (the reference keeps all of
        # that in the Python objects, exploration_model.py:101-107, behavioural_components.py:437-441)
        seed_visits, state = None, [float(bool(model._moving)), 0.0, 0.0]
        for c in comps:
            if isinstance(c, bc.ExperienceComponent):
                seed_visits = np.asarray(c._maze_map_exp, dtype=np.float64)
                state[1:] = [float(c._familiarity), float(int(c._ment_state))]
        if seed_visits is not None and not seed_visits.any():
            seed_visits = None
        res = exp.run(params, iterations, mouse.position, mouse.orientation, init_disc=init_disc,
                      generators=[model._rng], history=True, occupancy=True, init_visits=seed_visits,
                      init_model_state=state)
        if res.status[0] != 0:  # a step without any reachable endpoint: np.max([]) in exploration_model.py:92
            raise ValueError("zero-size array to reduction operation maximum which has no identity")
        if iterations <= 0:
            return
        # leave every object in the state the per-step loop would have left it in
        hp, ho = res.history_position[0], res.history_orientation[0]
        lib = _native.lib()
        _native.check(lib.nm_mouse_adopt_trajectory(mouse._h, _native.np_ptr(np.ascontiguousarray(hp[1:])),
                                                    _native.np_ptr(np.ascontiguousarray(ho[1:])), iterations))
        set_pcg64_state(model._rng, res.pcg_state[0], res.pcg_buffered[0])
        ms = res.model_state[0]
        model._prev_pos = np.array([ms[0], ms[1]])
        model._moving = np.bool_(ms[2] != 0.0)
        for c in comps:
            if isinstance(c, bc.ExperienceComponent):
                visits = res.occupancy[0].astype(np.uint32).astype(np.float64)  # seeded visits + this run's positions
                last = maze.cont2disc(hp[-1])
                visits[last[0], last[1]] -= 1.0  # the final position has not been "experienced" yet
                c._maze_map_exp[...] = visits
                c._familiarity = float(ms[3])
                c._ment_state = c.State.HOME if ms[4] != 0.0 else c.State.EXPLO
