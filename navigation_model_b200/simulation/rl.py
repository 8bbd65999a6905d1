"""Value learning on a conditioning session: the reference's ``simulation/rl.py`` interface on top of
the CUDA sweep kernels.

Same names, signatures and error behaviour as the reference (``rl.py:12-424``):

* :class:`Transition`, :class:`Update` records (``rl.py:12-35``);
* :class:`TD0`, :class:`PositiveConditioning`, :class:`NegativeConditioning` with the scalar per-step
  plugin call ``update(t) -> Update`` (``rl.py:154-262``) -- this is the plugin interface user code and
  the reference's tests drive one transition at a time, so it stays host Python;
* :class:`ShuffledReplays` (``rl.py:291-424``) and :class:`ConditioningExperiment` (``rl.py:38-121``).

``ConditioningExperiment.run`` / ``do_replays`` with the three built-in algorithms and ``ShuffledReplays``
do **not** loop in Python: the trajectory is flattened to a step stream by the native builder
(``nm_stream_build``) and the whole recurrence -- online pass and replay passes -- runs in one launch
(``nm_vtable_trace``: one parameter set, the table plus the ``dv`` / ``|rpe|`` every ``update`` returns, which fill the
replay buffer's ``Update`` entries like the reference's; many parameter sets: :mod:`navigation_model_b200.sweep`).  Replay passes are parameter-independent step records too: shuffled
passes are permutations of the online records; ``imaginary=True`` passes (``rl.py:388-418``) are transitions
sampled from the maze with the stdlib ``random`` module exactly as the reference samples them (seed it with
``random.seed`` for a replayable run), flattened by ``nm_stream_build_pairs``.  That path needs a CUDA device
and fails loudly without one.  User-defined ``RLAlgorithm`` / ``ReplayStrategy`` subclasses are host plugins by
nature and go through the generic per-transition loop.
"""
import random
import sys
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .. import _native


# --------------------------------------------------------------------------------------- records
@dataclass
class Transition:
    """One step of a session (``rl.py:12-24``)."""
    cs: Any = None    # starting continuous position
    ori: Any = None   # starting orientation
    s: Any = None     # starting discrete position
    cs1: Any = None   # arriving continuous position
    ori1: Any = None  # arriving orientation
    s1: Any = None    # arriving discrete state
    r: Any = None     # reward
    extra: dict = field(default_factory=dict)  # per-algorithm scratch ("next_states")


@dataclass
class Update:
    """A value update caused by a transition (``rl.py:28-35``)."""
    transition: Transition = None
    dv: Any = None   # value change applied to V[s]
    rpe: Any = None  # |reward prediction error|


# ------------------------------------------------------------------------------------ algorithms
class RLAlgorithm:
    """V-table holder + per-transition update rule (``rl.py:126-151``)."""

    #: NM_ALG_* code when the sweep kernel implements the rule, else None (host plugin)
    device_kind = None

    def __init__(self):
        self._v_table = None

    @property
    def v_table(self):
        if self._v_table is None:
            raise RuntimeError("V-table not initialized")
        return self._v_table

    @v_table.setter
    def v_table(self, v_table):
        self._v_table = v_table

    def update_v_table(self, s, dv):
        self._v_table[s[0], s[1]] += dv

    def update(self, t):
        raise NotImplementedError()


class TD0(RLAlgorithm):
    """``V(s) += alpha * (r + gamma * V(s') - V(s))`` (``rl.py:240-262``)."""

    device_kind = _native.ALG_TD0

    def __init__(self, alpha, discount_rate):
        super().__init__()
        self._alpha = alpha
        self._discount_rate = discount_rate

    def update(self, t):
        v = self._v_table
        rpe = t.r + self._discount_rate * v[t.s1[0], t.s1[1]] - v[t.s[0], t.s[1]]
        dv = self._alpha * rpe
        self.update_v_table(t.s, dv)
        return Update(transition=t, dv=dv, rpe=abs(rpe))


class _LookaheadConditioning(RLAlgorithm):
    """Shared body of the max / min one-step look-ahead learners (``rl.py:154-237``).

    The candidate next states of a transition are the tile-visitable cells under the endpoints of the
    algorithm's own mouse placed at the start point (``rl.py:177-182``); they are cached in
    ``t.extra["next_states"]`` so that replays reuse them.
    """

    _pick = None  # np.max / np.min

    def __init__(self, alpha, discount_rate, mouse, maze):
        super().__init__()
        self._alpha = alpha
        self._discount_rate = discount_rate
        self._mouse = mouse
        self._maze = maze

    def _next_states(self, t):
        if "next_states" not in t.extra:
            self._mouse.move_to(t.cs)
            cells = self._maze.cont2disc_list(self._mouse.get_endpoints())
            keep = self._maze.are_visitable(cells)
            t.extra["next_states"] = [c for c, ok in zip(cells, keep) if ok]
        return t.extra["next_states"]

    def update(self, t):
        v = self._v_table
        boot = type(self)._pick([v[c[0], c[1]] for c in self._next_states(t)])
        rpe = t.r + self._discount_rate * boot - v[t.s[0], t.s[1]]
        dv = self._alpha * rpe
        self.update_v_table(t.s, dv)
        self._mouse.move_to(t.cs1)
        return Update(transition=t, dv=dv, rpe=abs(rpe))


class PositiveConditioning(_LookaheadConditioning):
    """``V(s) += alpha * (R + gamma * max V(s') - V(s))`` (``rl.py:154-194``)."""
    device_kind = _native.ALG_POSITIVE
    _pick = staticmethod(np.max)


class NegativeConditioning(_LookaheadConditioning):
    """``V(s) += alpha * (P + gamma * min V(s') - V(s))`` (``rl.py:197-237``)."""
    device_kind = _native.ALG_NEGATIVE
    _pick = staticmethod(np.min)


# ------------------------------------------------------------------------------- replay strategies
class ReplayStrategy:
    """Interface of a replay strategy (``rl.py:267-288``)."""

    def append(self, update):
        raise NotImplementedError

    def clear_buffer(self):
        raise NotImplementedError

    def offline_replays(self, alg):
        raise NotImplementedError


class ShuffledReplays(ReplayStrategy):
    """Present the whole buffer ``n_times``, reshuffled (cumulatively, in place) before every pass
    (``rl.py:291-315, 368-424``)."""

    def __init__(self, n_times, seed=None, imaginary=False, maze=None, possible_actions=None):
        super().__init__()
        self._n_times = n_times
        self._buffer = []
        self._rng = np.random.default_rng(seed)
        self._imaginary = imaginary
        self._maze = maze
        self._possible_actions = possible_actions

    def append(self, update):
        self._buffer.append(update)

    def clear_buffer(self):
        self._buffer = []

    # -- device path support ------------------------------------------------------------------
    def next_orders(self):
        """Advance the generator exactly as ``n_times`` in-place shuffles of the buffer would
        (``rl.py:420``; shuffling ``arange(n)`` draws the same permutation as shuffling the list) and
        return the ``[n_times, n]`` index orders, each expressed in the buffer order *before* the call.
        The buffer itself is left in the final permuted order, like the reference's."""
        n = len(self._buffer)
        cur = np.arange(n)
        orders = np.empty((self._n_times, n), dtype=np.int64)
        for k in range(self._n_times):
            self._rng.shuffle(cur)
            orders[k] = cur
        if self._n_times > 0:
            self._buffer = [self._buffer[i] for i in cur]
        return orders

    def _rewarded_arrivals(self, alg):
        wanted = -1.0 if isinstance(alg, NegativeConditioning) else (1.0 if isinstance(alg, PositiveConditioning)
                                                                      else None)
        if wanted is None:
            return set()
        return {u.transition.s1 for u in self._buffer if u.transition.r == wanted}

    def imaginary_transitions(self, alg):
        """The ``n_times * len(buffer)`` transitions of the imaginary replays, sampled exactly as ``rl.py:388-418``
        samples them and in the same order: a start tile with ``random.choice`` over the visitable tiles, a heading
        with ``random.uniform(-pi, pi)``, an arrival tile with ``random.choice`` among the tiles under the endpoints the
        ray check accepts; reward +-1 when that tile was ever reached with a stimulation.  The draws come from the
        stdlib ``random`` module like the reference's (seed it with ``random.seed`` for a replayable sequence); they do
        not depend on the value table, so they can be drawn ahead of the updates."""
        hot = self._rewarded_arrivals(alg)
        sign = -1 if isinstance(alg, NegativeConditioning) else 1
        coords = self._maze.get_visitable_coordinates()
        acts = [np.asarray(a, dtype=np.float64) for a in self._possible_actions]
        out = []
        for _ in range(self._n_times * len(self._buffer)):
            t = Transition()
            t.s = random.choice(coords)
            t.cs = self._maze.disc2cont(t.s)
            t.ori = random.uniform(-np.pi, np.pi)
            c, s = np.cos(t.ori), np.sin(t.ori)
            rot = np.array([[c, -s], [s, c]])
            ends = np.zeros((len(acts), 2))
            for i, act in enumerate(acts):
                ends[i] = t.cs + np.matmul(rot, act)
            ok = self._maze.are_movements_possible_cont(t.cs, ends)
            t.s1 = random.choice(self._maze.cont2disc_list(ends[ok]))
            t.cs1 = self._maze.disc2cont(t.s1)
            t.ori1 = np.arctan2(t.cs1[1] - t.cs[1], t.cs1[0] - t.cs[0])
            t.r = sign if t.s1 in hot else 0
            out.append(t)
        return out

    def offline_replays(self, alg):
        """Host loop (plugin interface).  ``ConditioningExperiment`` bypasses it for the built-in
        algorithms and runs the passes on the device."""
        if self._imaginary:
            for t in self.imaginary_transitions(alg):
                alg.update(t)
            return
        for _ in range(self._n_times):
            self._rng.shuffle(self._buffer)
            for u in self._buffer:
                alg.update(u.transition)


# ------------------------------------------------------------------------------------ experiment
class ConditioningExperiment:
    """Follow a recorded trajectory, learn a value table, optionally replay (``rl.py:38-121``)."""

    def __init__(self, maze, algorithm, replay_strategy=None, v_table=None, debug=False):
        self._maze = maze
        self._alg = algorithm
        self._alg.v_table = v_table if v_table is not None else np.zeros(self._maze.shape)
        self._replay_strategy = replay_strategy
        self._debug = debug

    # -- dispatch ---------------------------------------------------------------------------------
    def _device_eligible(self):
        from .maze import Maze
        from .mouse import Mouse
        alg, rs = self._alg, self._replay_strategy
        if type(alg) not in (TD0, PositiveConditioning, NegativeConditioning):
            return False
        if type(rs) is not ShuffledReplays:
            return False
        if rs._imaginary and (rs._maze is not self._maze or rs._possible_actions is None):
            return False
        if type(self._maze) is not Maze:
            return False
        if alg.device_kind != _native.ALG_TD0:
            if type(alg._mouse) is not Mouse or alg._maze is not self._maze:
                return False
            if not 1 <= len(alg._mouse.get_endpoints()) <= _native.NM_MAX_CAND:
                return False
        v = alg._v_table
        return isinstance(v, np.ndarray) and v.shape == tuple(self._maze.shape) and v.dtype == np.float64

    def run(self, trajectory, orientations, reward, do_replays=False):
        if len(trajectory) != len(reward):
            raise RuntimeError("trajectory and reward must have the same length")
        if self._device_eligible():
            self._run_device(trajectory, orientations, reward, do_replays)
        else:
            self._run_host(trajectory, orientations, reward, do_replays)

    def do_replays(self):
        if self._replay_strategy is None:
            print("Asked to do replay, but no replay strategy given", file=sys.stderr)
        elif self._device_eligible():
            self._replays_device()
        else:
            self._replay_strategy.offline_replays(self._alg)

    @property
    def v_table(self):
        return self._alg.v_table

    # -- generic host loop (user-defined plugins) ---------------------------------------------------
    def _run_host(self, trajectory, orientations, reward, do_replays):
        cs, ori = trajectory[0], orientations[0]
        s = self._maze.get_closest_visitable_cont(cs)
        for i in range(1, len(trajectory)):
            t = Transition(cs=cs, ori=ori, s=s, cs1=trajectory[i], ori1=orientations[i],
                           s1=self._maze.get_closest_visitable_cont(trajectory[i]), r=reward[i])
            self._replay_strategy.append(self._alg.update(t))  # AttributeError if None, like rl.py:98
            cs, ori, s = t.cs1, t.ori1, t.s1
        if do_replays:
            self.do_replays()

    # -- device path ----------------------------------------------------------------------------------
    def _run_device(self, trajectory, orientations, reward, do_replays):
        from .. import sweep
        alg, rs = self._alg, self._replay_strategy
        traj = np.ascontiguousarray(np.asarray(trajectory, dtype=np.float64).reshape(-1, 2))
        rew = np.ascontiguousarray(np.asarray(reward, dtype=np.float64).reshape(-1))
        if len(traj) == 0:
            raise IndexError("list index out of range")  # trajectory[0] in the reference
        with_cand = alg.device_kind != _native.ALG_TD0
        mouse = alg._mouse if with_cand else None
        recs, trans = sweep.build_session_records(self._maze, traj, rew, mouse=mouse,
                                                  orientations=orientations, make_transitions=True)
        fresh = [Update(transition=t) for t in trans]
        for u in fresh:
            rs.append(u)
        n_online = len(recs)
        if do_replays:
            recs = np.concatenate([recs, self._replay_records()])
        trace = self._launch(recs, n_online)
        if trace is not None:  # what the learner's update returned for every online step (rl.py:190-194, 260-262)
            dv, rpe = trace
            for i, u in enumerate(fresh):
                u.dv, u.rpe = dv[i], rpe[i]

    def _replay_records(self):
        from .. import sweep
        rs = self._replay_strategy
        if rs._imaginary:
            # sampled transitions are parameter-independent: draw them on the host (same stdlib draws as the
            # reference), flatten them to records (the algorithm's mouse walks through them as in rl.py:178-189)
            # and run the updates in the same launch
            alg = self._alg
            mouse = alg._mouse if alg.device_kind != _native.ALG_TD0 else None
            return sweep.records_from_transitions(self._maze, rs.imaginary_transitions(alg), mouse=mouse)
        all_recs = np.zeros(len(rs._buffer), dtype=_native.STEP_REC_DTYPE)
        for i, u in enumerate(rs._buffer):
            all_recs[i] = sweep.record_from_transition(self._maze, u.transition)
        arrivals = [u.transition.cs1 for u in rs._buffer]
        orders = rs.next_orders()
        if orders.size == 0:
            return np.zeros(0, dtype=_native.STEP_REC_DTYPE)
        flat = orders.reshape(-1)
        if self._alg.device_kind != _native.ALG_TD0:
            # every replayed update of the look-ahead learners ends with mouse.move_to(t.cs1) (rl.py:189, 232): the
            # tables come from the launch, the learner's mouse (pose, heading, history) is walked through the same
            # arrival points here, so that a later run() on this experiment starts from the reference's pose
            pts = np.ascontiguousarray(np.asarray(arrivals, dtype=np.float64).reshape(-1, 2)[flat])
            _native.check(_native.lib().nm_mouse_move_through(self._alg._mouse._h, _native.np_ptr(pts), len(pts)))
        return all_recs[flat]

    def _replays_device(self):
        recs = self._replay_records()
        if len(recs):
            self._launch(recs, 0)

    def _launch(self, recs, n_online):
        from .. import sweep
        alg = self._alg
        v = alg._v_table
        alpha = float(np.asarray(alg._alpha).reshape(-1)[0])
        gamma = float(np.asarray(alg._discount_rate).reshape(-1)[0])
        traced = sweep.vtable_trace_single(self._maze, alg.device_kind, recs, n_online, alpha, gamma, v)
        if traced is None:  # more states than the trace kernel's table holds: tables only (Update.dv / rpe stay None)
            v[...] = sweep.vtable_single(self._maze, alg.device_kind, recs, n_online, alpha, gamma, v)
            return None
        v[...] = traced[0]  # the reference mutates the table in place
        return traced[1], traced[2]
