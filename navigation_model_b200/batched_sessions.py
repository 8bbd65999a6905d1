"""Session ingestion on the device: many raw recordings -> subsampled sessions ->
orientations / maze-adjusted positions / step streams, without the host touching a sample.

:class:`SessionBatch` is the batched counterpart of ``Session`` / ``SessionList`` (``analysis/data.py:7-369``): the
recordings of ``S`` sessions are stored back to back in HBM (session ``j`` owns ``counts[j]`` samples from index
``offsets[j]``) and every preparation step of the reference runs as a kernel over all sessions:

========================  =======================================  ===================================
method                    reference                                C-ABI
========================  =======================================  ===================================
``subsample(dt)``         ``Session.__init__`` ``data.py:39-74``   ``nm_sessions_subsample_device``
``adjust_to_maze(maze)``  ``adjust_positions_to_maze`` ``:372``    ``nm_positions_adjust_device``
``orientations(...)``     ``compute_orientations``                 ``nm_sessions_orientations_device``
                          ``metrics.py:653-694``
``streams(maze, ...)``    ``ConditioningExperiment.run``           ``nm_streams_build_device``
                          ``rl.py:85-102`` + ``rl.py:177-189``
========================  =======================================  ===================================

The result of :meth:`SessionBatch.streams` plugs into :class:`~navigation_model_b200.sweep.ConditioningSweep`
(:meth:`ConditioningSweep.use_streams`).  PyTorch tensors are buffers only; there is no CPU fallback.
"""
import numpy as np

from . import _native
from .sweep import DeviceMaze


class DeviceBuiltStreams:
    """``nm_streams`` handle whose records were built on the device (same duck type as ``DeviceStreams``)."""

    def __init__(self, dmaze, handle, counts, has_candidates, replays=None):
        self.dmaze = dmaze
        self._lib = _native.lib()
        self._h = handle
        self.online = np.maximum(np.asarray(counts, dtype=np.int64) - 1, 0)
        extra = np.zeros_like(self.online) if replays is None else np.asarray(replays, dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.online + extra)]).astype(np.int64)
        self.S = len(self.online)
        self.total = int(self.offsets[-1])
        self.has_candidates = bool(has_candidates)

    @property
    def handle(self):
        return self._h

    @property
    def host_bytes(self):
        return 0  # nothing is staged on the host

    def reupload(self):
        pass  # the records never left the device

    def download(self):
        """Records as a list of ``S`` numpy structured arrays (tests / inspection)."""
        import torch
        recs = np.zeros(self.total, dtype=_native.STEP_REC_DTYPE)
        if self.total:
            with torch.cuda.device(self.dmaze.torch_device):
                _native.check(self._lib.nm_streams_download(self._h, _native.np_ptr(recs), _native.cuda_stream_ptr()))
        return [recs[self.offsets[j]:self.offsets[j + 1]] for j in range(self.S)]

    def with_replays(self, n_times, seed=None):
        """New streams whose sessions carry ``n_times`` shuffled replay passes after their online records
        (``ShuffledReplays.offline_replays``, ``rl.py:419-424``): the same cumulative ``Generator.shuffle`` orders the
        host path draws (every session from ``default_rng(seed)``), gathered on the device
        (``nm_streams_with_replays``).  This object stays valid."""
        import torch
        from .sweep import replay_order
        if np.any(self.offsets[1:] - self.offsets[:-1] != self.online):
            raise RuntimeError("these streams already carry replay passes")
        orders = [replay_order(int(n), n_times, seed=seed) for n in self.online]
        counts = np.ascontiguousarray([len(o) for o in orders], dtype=np.int64)
        order = np.ascontiguousarray(np.concatenate(orders) if len(orders) else np.zeros(0), dtype=np.int64)
        with torch.cuda.device(self.dmaze.torch_device):
            h = _native.check_handle(self._lib.nm_streams_with_replays(
                self._h, _native.np_ptr(order), _native.np_ptr(counts), _native.cuda_stream_ptr()))
        return DeviceBuiltStreams(self.dmaze, h, self.online + 1, self.has_candidates, replays=counts)

    def close(self):
        h, self._h = self._h, None
        if h:
            self._lib.nm_streams_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SessionBatch:
    """``S`` recorded sessions resident in HBM.

    >>> batch = SessionBatch(timestamps, trajectories, rewards)        # ragged lists, one entry per session
    >>> sub = batch.subsample(0.1)                                     # Session(..., new_sampling_time=0.1)
    >>> sw = ConditioningSweep(maze, "positive", possible_actions=A8)
    >>> sw.use_streams(sub.streams(maze, A8))                          # no host round trip of the samples
    """

    def __init__(self, timestamps, trajectories, rewards=None, device=None):
        """
        :param timestamps: list of ``S`` 1-D arrays (or ``None`` for already sampled sessions)
        :param trajectories: list of ``S`` arrays ``[n_j, 2]``
        :param rewards: list of ``S`` 1-D arrays, or ``None``
        """
        import torch
        _native.require_device()
        if device is None:
            device = torch.cuda.current_device()
        self.torch_device = torch.device("cuda", int(device))
        S = len(trajectories)
        if S == 0:
            raise ValueError("need at least one session")
        trajs = [np.asarray(t, dtype=np.float64).reshape(-1, 2) for t in trajectories]
        lens = np.array([len(t) for t in trajs], dtype=np.int64)
        if timestamps is not None:
            times = [np.asarray(t, dtype=np.float64).reshape(-1) for t in timestamps]
            if len(times) != S or any(len(a) != n for a, n in zip(times, lens)):
                raise RuntimeError("Session: time and trajectory must have the same lenght")  # data.py:29-30
        if rewards is not None:
            rews = [np.asarray(r, dtype=np.float64).reshape(-1) for r in rewards]
            if len(rews) != S or any(len(a) != n for a, n in zip(rews, lens)):
                raise RuntimeError("Session: reward must have the same length as time and trajectory")  # :31-32
        self.offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        self.counts = lens.copy()
        up = lambda parts: torch.from_numpy(np.ascontiguousarray(np.concatenate(parts))).pin_memory().to(  # noqa: E731
            self.torch_device, non_blocking=True)
        self.traj = up(trajs)
        self.time = None if timestamps is None else up(times)
        self.reward = None if rewards is None else up(rews)
        self.sampling_time = None if timestamps is None else np.array(
            [np.mean(np.diff(t)) if len(t) > 1 else np.nan for t in times])
        self._off_t = torch.from_numpy(self.offsets).to(self.torch_device)
        self._cnt_t = torch.from_numpy(self.counts).to(self.torch_device)
        torch.cuda.current_stream(self.torch_device).synchronize()
        self._lib = _native.lib()
        self._dmaze = {}

    # ------------------------------------------------------------------------------------------------
    @classmethod
    def _from_parts(cls, src, traj, reward, counts_t, counts, sampling_time):
        new = cls.__new__(cls)
        new.torch_device = src.torch_device
        new.offsets = src.offsets
        new._off_t = src._off_t
        new.counts = counts
        new._cnt_t = counts_t
        new.traj = traj
        new.reward = reward
        new.time = None
        new.sampling_time = sampling_time
        new._lib = src._lib
        new._dmaze = src._dmaze
        return new

    @property
    def n_sessions(self):
        return len(self.counts)

    def _device_maze(self, maze):
        if isinstance(maze, DeviceMaze):
            return maze
        dm = self._dmaze.get(id(maze))
        if dm is None or dm.maze is not maze:
            dm = DeviceMaze(maze, self.torch_device.index)
            self._dmaze[id(maze)] = dm
        return dm

    # ------------------------------------------------------------------------------------------------
    def subsample(self, new_sampling_time):
        """``Session(timestamps, trajectory, new_sampling_time, reward)`` for every session (``data.py:39-74``):
        one sample per window of at least ``new_sampling_time`` seconds -- the first finite position and the first
        reward that is neither 0 nor NaN (else 0.0).  Returns a new batch (same offsets, new counts)."""
        import torch
        if self.time is None:
            raise RuntimeError("this batch has no timestamps (already subsampled)")
        S = self.n_sessions
        traj_out = torch.empty_like(self.traj)
        rew_out = None if self.reward is None else torch.empty_like(self.reward)
        counts_t = torch.empty(S, dtype=torch.int64, device=self.torch_device)
        with torch.cuda.device(self.torch_device):
            _native.check(self._lib.nm_sessions_subsample_device(
                _native.tensor_ptr(self.time), _native.tensor_ptr(self.traj), _native.tensor_ptr(self.reward),
                _native.tensor_ptr(self._off_t), S, float(new_sampling_time), _native.tensor_ptr(traj_out),
                _native.tensor_ptr(rew_out), _native.tensor_ptr(counts_t), _native.cuda_stream_ptr()))
        counts = counts_t.cpu().numpy()  # S int64: the only thing that returns to the host
        return SessionBatch._from_parts(self, traj_out, rew_out, counts_t, counts,
                                        np.full(S, float(new_sampling_time)))

    def adjust_to_maze(self, maze):
        """``adjust_positions_to_maze(trajectory, maze)`` for every session (``data.py:372-388``): new batch whose
        points outside the visitable tiles sit on the centre of the nearest visitable tile."""
        import torch
        dm = self._device_maze(maze)
        out = torch.empty_like(self.traj)
        with torch.cuda.device(self.torch_device):
            _native.check(self._lib.nm_positions_adjust_device(dm.handle, _native.tensor_ptr(self.traj),
                                                               int(self.traj.shape[0]), _native.tensor_ptr(out),
                                                               _native.cuda_stream_ptr()))
        new = SessionBatch._from_parts(self, out, self.reward, self._cnt_t, self.counts, self.sampling_time)
        new.time = self.time
        return new

    def orientations(self, maze=None, possible_actions=None, as_list=True):
        """``compute_orientations(trajectory, maze, possible_actions)`` for every session
        (``metrics.py:653-694``).  ``as_list``: list of ``S`` numpy arrays; else the flat device tensor (entries of
        session ``j`` at ``offsets[j] : offsets[j] + counts[j]``).  A session that never moves raises ``IndexError``
        like the reference."""
        import torch
        if (maze is None) != (possible_actions is None):
            raise ValueError("give both maze and possible_actions, or neither")  # the reference fails on None.cont2disc
        acts, dm = None, None
        if maze is not None:
            dm = self._device_maze(maze)
            acts = np.ascontiguousarray(np.asarray(possible_actions, dtype=np.float64).reshape(-1, 2))
            if len(acts) > _native.NM_MAX_CAND:
                raise ValueError(f"at most {_native.NM_MAX_CAND} actions are supported")
        S = self.n_sessions
        out = torch.empty(int(self.traj.shape[0]), dtype=torch.float64, device=self.torch_device)
        first = torch.empty(S, dtype=torch.int64, device=self.torch_device)
        with torch.cuda.device(self.torch_device):
            _native.check(self._lib.nm_sessions_orientations_device(
                None if dm is None else dm.handle, _native.tensor_ptr(self.traj), _native.tensor_ptr(self._off_t),
                _native.tensor_ptr(self._cnt_t), S, _native.np_ptr(acts), 0 if acts is None else len(acts),
                _native.tensor_ptr(out), _native.tensor_ptr(first), _native.cuda_stream_ptr()))
        if (first.cpu().numpy() < 0).any():
            raise IndexError("list index out of range")  # metrics.py:694 on a session without a single move
        if not as_list:
            return out
        flat = out.cpu().numpy()
        return [flat[o:o + c] for o, c in zip(self.offsets[:-1], self.counts)]

    def streams(self, maze, possible_actions=None, mouse_init=None):
        """Step streams of all sessions, built on the device (``nm_streams_build_device``): the transitions
        ``ConditioningExperiment.run`` extracts (``rl.py:85-102``) and, with ``possible_actions``, the candidate next
        states ``PositiveConditioning.update`` computes on first sight (``rl.py:177-189``) with the algorithm's mouse
        ``Mouse(trajectory[0], 0.0, possible_actions)`` (or ``mouse_init[j] = (x, y, ori)``)."""
        import torch
        if self.reward is None:
            raise RuntimeError("value learning needs a reward sequence")
        dm = self._device_maze(maze)
        acts = None
        if possible_actions is not None:
            acts = np.ascontiguousarray(np.asarray(possible_actions, dtype=np.float64).reshape(-1, 2))
            if len(acts) > _native.NM_MAX_CAND:
                raise ValueError(f"at most {_native.NM_MAX_CAND} actions are supported")
        init = None
        if mouse_init is not None:
            init = np.ascontiguousarray(np.asarray(mouse_init, dtype=np.float64).reshape(self.n_sessions, 3))
        src_off = np.ascontiguousarray(self.offsets[:-1])
        with torch.cuda.device(self.torch_device):
            h = _native.check_handle(self._lib.nm_streams_build_device(
                dm.handle, _native.tensor_ptr(self.traj), _native.tensor_ptr(self.reward), _native.np_ptr(src_off),
                _native.np_ptr(np.ascontiguousarray(self.counts)), self.n_sessions, _native.np_ptr(acts),
                0 if acts is None else len(acts), _native.np_ptr(init), _native.cuda_stream_ptr()))
        return DeviceBuiltStreams(dm, h, self.counts, acts is not None)

    # ------------------------------------------------------------------------------------------------
    def to_host(self):
        """``[(trajectory [n_j, 2], reward [n_j] or None)]`` per session (tests / inspection)."""
        traj = self.traj.cpu().numpy()
        rew = None if self.reward is None else self.reward.cpu().numpy()
        out = []
        for o, c in zip(self.offsets[:-1], self.counts):
            out.append((traj[o:o + c], None if rew is None else rew[o:o + c]))
        return out

    def sessions(self, maze=None, possible_actions=None):
        """Reference-style ``Session`` objects of the (sub)sampled data."""
        from .analysis.data import Session
        out = []
        for (traj, rew), dt in zip(self.to_host(), self.sampling_time):
            d = {"sampling_time": dt, "trajectory": list(traj)}
            if rew is not None:
                d["reward"] = list(rew)
            out.append(Session.from_dict(d))
        return out
