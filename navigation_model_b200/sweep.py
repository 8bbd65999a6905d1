"""Batched (parameter-grid) value learning and log-likelihood fitting on the GPU.

This is the host layer above the C-ABI for the fitness hot path:

* :func:`build_session_records` -- flatten one session into the parameter-independent step stream
  (native ``nm_stream_build``; restates ``ConditioningExperiment.run`` ``rl.py:85-102`` and the
  candidate computation of ``PositiveConditioning`` ``rl.py:177-182``);
* :class:`DeviceMaze`, :class:`DeviceStreams` -- handles of the tables resident in HBM;
* :class:`ConditioningSweep` -- the grid version of ``ConditioningExperiment``: V-tables and per-session
  log-likelihoods for ``P`` parameter sets x ``S`` sessions in one kernel launch, per-shard arg-min and
  the (tiny) cross-GPU gather of the best fits (``torch.distributed``; NCCL over NVLink on the GPU box).

PyTorch is used only for device buffers, streams and the process group.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _native
from .simulation.maze import Maze
from .simulation.mouse import Mouse

ALGORITHMS = {"td0": _native.ALG_TD0, "positive": _native.ALG_POSITIVE, "negative": _native.ALG_NEGATIVE,
              "qlearning": _native.ALG_POSITIVE}  # max-backup on deterministic after-states == Q-learning


def algorithm_code(alg):
    """Accept a name, an ``NM_ALG_*`` code, or a reference-style algorithm class / instance."""
    if isinstance(alg, str):
        return ALGORITHMS[alg.lower()]
    if isinstance(alg, (int, np.integer)):
        if int(alg) not in (0, 1, 2):
            raise ValueError(f"unknown algorithm code {alg}")
        return int(alg)
    kind = getattr(alg, "device_kind", None)
    if kind is None:
        raise ValueError(f"{alg!r} has no device implementation")
    return kind


# ----------------------------------------------------------------------------------- host handles
class HostMaze:
    """Host-only ``nm_maze`` handle (geometry queries + stream building; no GPU needed)."""

    def __init__(self, maze, device=-1, stream=None):
        if not isinstance(maze, Maze):
            raise TypeError("expected a navigation_model_b200 Maze")
        self.maze = maze
        tabs = maze.flat_tables()
        self.tables = tabs
        self.n_tiles = maze.size_x * maze.size_y
        self._lib = _native.lib()
        tiles = np.ascontiguousarray(tabs["tiles"])
        sub = None if tabs["subareas"] is None else np.ascontiguousarray(tabs["subareas"])
        self.device = device
        self._h = _native.check_handle(self._lib.nm_maze_upload(
            _native.np_ptr(tiles), _native.np_ptr(sub), maze.size_x, maze.size_y, float(maze.size_tile), int(device),
            stream))
        self.num_states = int(self._lib.nm_maze_num_states(self._h))

    @property
    def handle(self):
        return self._h

    def close(self):
        h, self._h = self._h, None
        if h:
            self._lib.nm_maze_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_host_maze_cache = {}


def host_maze(maze):
    key = id(maze)
    hm = _host_maze_cache.get(key)
    if hm is None or hm.maze is not maze:
        hm = HostMaze(maze)
        _host_maze_cache[key] = hm
    return hm


class DeviceMaze(HostMaze):
    """``nm_maze`` handle with the dense tables uploaded to HBM on ``device``."""

    def __init__(self, maze, device=None):
        import torch
        _native.require_device()
        if device is None:
            device = torch.cuda.current_device()
        self.torch_device = torch.device("cuda", int(device))
        with torch.cuda.device(self.torch_device):
            super().__init__(maze, device=int(device), stream=_native.cuda_stream_ptr())


def build_session_records(maze, trajectory, reward, mouse=None, orientations=None, make_transitions=False):
    """Flatten one session into ``nm_step_rec`` records (numpy structured array, ``T-1`` entries).

    :param mouse: the algorithm's own :class:`Mouse` (``PositiveConditioning(..., mouse, maze)``); it is
        driven through the session exactly as the reference drives it.  ``None``: no candidates (TD0
        value learning only -- a likelihood needs candidates).
    :param make_transitions: also return reference-style ``Transition`` objects for the replay buffer
    """
    from .simulation.rl import Transition
    lib = _native.lib()
    hm = host_maze(maze)
    traj = np.ascontiguousarray(np.asarray(trajectory, dtype=np.float64).reshape(-1, 2))
    rew = np.ascontiguousarray(np.asarray(reward, dtype=np.float64).reshape(-1))
    if len(traj) != len(rew):
        raise RuntimeError("trajectory and reward must have the same length")
    T = len(traj)
    recs = np.zeros(max(T - 1, 0), dtype=_native.STEP_REC_DTYPE)
    if T >= 1:
        mh = None if mouse is None else mouse._h
        n = _native.check(lib.nm_stream_build(hm.handle, _native.np_ptr(traj), _native.np_ptr(rew), T, mh,
                                              _native.np_ptr(recs) if T > 1 else None))
        assert n == T - 1
    if not make_transitions:
        return recs
    tos = hm.tables["tile_of_state"]
    Y = maze.size_y
    cells = [(int(t) // Y, int(t) % Y) for t in tos]
    trans = []
    for i, rec in enumerate(recs):
        t = Transition(cs=traj[i], ori=None if orientations is None else orientations[i], s=cells[rec["s"]],
                       cs1=traj[i + 1], ori1=None if orientations is None else orientations[i + 1],
                       s1=cells[rec["s1"]], r=rew[i + 1])
        if mouse is not None:
            t.extra["next_states"] = [cells[c] for c in rec["cand"][:rec["ncand"]]]
        t.extra["_rec"] = rec
        trans.append(t)
    return recs, trans


def record_from_transition(maze, t):
    """``nm_step_rec`` of a reference-style ``Transition`` (uses ``extra['next_states']`` if present)."""
    rec = t.extra.get("_rec")
    if rec is not None:
        return rec
    sot = host_maze(maze).tables["state_of_tile"]
    Y = maze.size_y
    out = np.zeros((), dtype=_native.STEP_REC_DTYPE)
    out["r"] = t.r
    out["s"] = sot[t.s[0] * Y + t.s[1]]
    out["s1"] = sot[t.s1[0] * Y + t.s1[1]]
    out["obs"] = _native.NM_OBS_NONE
    ns = t.extra.get("next_states", [])
    if len(ns) > _native.NM_MAX_CAND:
        raise ValueError(f"at most {_native.NM_MAX_CAND} candidate next states are supported")
    out["ncand"] = len(ns)
    for k, c in enumerate(ns):
        out["cand"][k] = sot[c[0] * Y + c[1]]
        if out["obs"] == _native.NM_OBS_NONE and out["cand"][k] == out["s1"]:
            out["obs"] = k
    return out


def records_from_transitions(maze, transitions, mouse=None):
    """``nm_step_rec`` records of independent ``Transition`` objects (``cs``, ``cs1``, ``s``, ``s1``, ``r``) -- each one
    a fresh transition as ``RLAlgorithm.update`` first sees it (no cached ``next_states``): with ``mouse`` (the
    algorithm's own) the candidate next states are computed by driving it through every transition like
    ``PositiveConditioning.update`` does (``rl.py:177-189``)."""
    n = len(transitions)
    recs = np.zeros(n, dtype=_native.STEP_REC_DTYPE)
    if n == 0:
        return recs
    cs = np.ascontiguousarray(np.array([t.cs for t in transitions], dtype=np.float64).reshape(n, 2))
    cs1 = np.ascontiguousarray(np.array([t.cs1 for t in transitions], dtype=np.float64).reshape(n, 2))
    s = np.ascontiguousarray(np.array([t.s for t in transitions], dtype=np.int32).reshape(n, 2))
    s1 = np.ascontiguousarray(np.array([t.s1 for t in transitions], dtype=np.int32).reshape(n, 2))
    r = np.ascontiguousarray(np.array([t.r for t in transitions], dtype=np.float64))
    got = _native.check(_native.lib().nm_stream_build_pairs(
        host_maze(maze).handle, _native.np_ptr(cs), _native.np_ptr(cs1), _native.np_ptr(s), _native.np_ptr(s1),
        _native.np_ptr(r), n, None if mouse is None else mouse._h, _native.np_ptr(recs)))
    assert got == n
    return recs


def replay_order(n, n_times, seed=None, rng=None):
    """Indices, in execution order, of ``n_times`` shuffled replay passes over a buffer of ``n`` updates
    (``ShuffledReplays.offline_replays``, ``rl.py:419-424``: cumulative in-place ``Generator.shuffle``)."""
    if rng is None:
        rng = np.random.default_rng(seed)
    idx = np.arange(int(n))
    out = []
    for _ in range(int(n_times)):
        rng.shuffle(idx)
        out.append(idx.copy())
    if not out:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate(out)


def replay_records(records, n_times, seed=None, rng=None):
    """Records of ``n_times`` shuffled replay passes over ``records`` in execution order."""
    return records[replay_order(len(records), n_times, seed=seed, rng=rng)]


class DeviceStreams:
    """``S`` session streams resident in HBM (``nm_streams`` handle)."""

    def __init__(self, dmaze, sessions, online=None, pinned=True, staging=None):
        """
        :param sessions: list of record arrays (online records followed by replay records)
        :param online: per session, the number of leading online records (default: all)
        :param staging: pinned uint8 tensor already holding the concatenated records (re-upload path)
        """
        import torch
        self.dmaze = dmaze
        self._lib = _native.lib()
        if len(sessions) == 0:
            raise ValueError("need at least one session")
        lens = np.array([len(s) for s in sessions], dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        self.online = lens.copy() if online is None else np.asarray(online, dtype=np.int64).copy()
        self.S = len(sessions)
        self.total = int(self.offsets[-1])
        if staging is None:
            sessions = [np.ascontiguousarray(s, dtype=_native.STEP_REC_DTYPE) for s in sessions]
            flat = np.concatenate(sessions) if self.total else np.zeros(0, dtype=_native.STEP_REC_DTYPE)
            staging = torch.from_numpy(flat.view(np.uint8).reshape(-1).copy())
            if pinned and self.total:
                staging = staging.pin_memory()
        # pinned staging buffer: kept alive by this object; the copy is enqueued and synchronised below
        self._host = staging
        with torch.cuda.device(dmaze.torch_device):
            self._h = _native.check_handle(self._lib.nm_streams_upload(
                C.c_void_p(self._host.data_ptr()), _native.np_ptr(self.offsets), _native.np_ptr(self.online), self.S,
                dmaze.num_states, dmaze.device, _native.cuda_stream_ptr()))
            torch.cuda.current_stream().synchronize()

    @property
    def handle(self):
        return self._h

    @property
    def host_bytes(self):
        return int(self._host.numel())

    def reupload(self):
        """Copy the pinned staging buffer to the device again (``nm_streams_update``; asynchronous on
        torch's current stream)."""
        import torch
        with torch.cuda.device(self.dmaze.torch_device):
            _native.check(self._lib.nm_streams_update(self._h, C.c_void_p(self._host.data_ptr()),
                                                      _native.cuda_stream_ptr()))

    def close(self):
        h, self._h = self._h, None
        if h:
            self._lib.nm_streams_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_param(x, P=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    if P is not None and a.size == 1 and P != 1:
        a = np.full(P, a[0])
    return a


def _params3(alpha, beta, gamma):
    """Broadcast scalars against the longest of the three arrays; ``ValueError`` on any other length mismatch."""
    alpha, beta, gamma = _as_param(alpha), _as_param(beta), _as_param(gamma)
    P = max(alpha.size, beta.size, gamma.size)
    alpha, beta, gamma = _as_param(alpha, P), _as_param(beta, P), _as_param(gamma, P)
    if not (alpha.size == beta.size == gamma.size == P):
        raise ValueError("alpha, beta and gamma must have the same length")
    return alpha, beta, gamma, P


def regroup_order(alpha, gamma, run=32):
    """Order that brings parameter sets with equal ``(alpha, gamma)`` together, so that every aligned run of ``run``
    consecutive sets shares them and the sweep keeps ONE table per run (the tables do not depend on beta).

    ``None`` when the arrays are already laid out that way (a grid with beta fastest), or when no such order exists
    without padding (some ``(alpha, gamma)`` pair occurs a number of times that is not a multiple of ``run``, other
    than the last one): the sweep then runs as given.  The sort is stable: sets of a pair keep their order."""
    P = alpha.size
    if P <= 1:
        return None

    def uniform_runs(a, g):
        n_full = (P // run) * run
        ok = True
        if n_full:
            A, G = a[:n_full].reshape(-1, run), g[:n_full].reshape(-1, run)
            ok = bool((A == A[:, :1]).all() and (G == G[:, :1]).all())
        if ok and n_full < P:
            ok = bool((a[n_full:] == a[n_full]).all() and (g[n_full:] == g[n_full]).all())
        return ok

    if uniform_runs(alpha, gamma):
        return None
    order = np.lexsort((gamma, alpha))  # stable; alpha major, gamma minor
    a, g = alpha[order], gamma[order]
    return order if uniform_runs(a, g) else None


def sharded_best_fit(base, ll_device, alpha, beta, gamma, group=None, v_init=None, regroup=True):
    """Grid search shared by the three sweeps: minimise the negative log-likelihood summed over sessions.

    The parameter sets are first brought into runs of 32 sharing ``(alpha, gamma)`` when a re-ordering achieves that
    (:func:`regroup_order`; the per-run-table kernels then apply whatever axis order the caller's grid has), the flat
    index of that order is sharded in contiguous blocks over the ranks of an initialised process group, and the
    per-shard ``(nll, index in the CALLER's order)`` pairs -- 16 bytes per rank -- are all-gathered.  Ties go to the
    lowest caller index, like ``numpy.argmin``.

    ``ll_device(alpha_t, beta_t, gamma_t, v_init_slice_or_None) -> ll[S, n]`` runs one shard on the device."""
    import torch
    import torch.distributed as dist
    alpha, beta, gamma, P = _params3(alpha, beta, gamma)
    per_unit = v_init is not None and np.asarray(v_init).ndim == 3
    order = regroup_order(alpha, gamma) if (regroup and not per_unit) else None
    if order is not None:
        alpha, beta, gamma = (np.ascontiguousarray(x[order]) for x in (alpha, beta, gamma))
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(P, rank, world)
    dev = base.torch_device
    if hi > lo:
        vi = None if v_init is None else (np.asarray(v_init)[lo:hi] if per_unit else v_init)
        ll = ll_device(base._to_dev(torch, alpha[lo:hi]), base._to_dev(torch, beta[lo:hi]),
                       base._to_dev(torch, gamma[lo:hi]), vi)
        if order is None:
            return reduce_best(base.argmin_device(ll), lo, group=group)
        nll = torch.empty(hi - lo, dtype=torch.float64, device=dev)
        best = base.argmin_device(ll, nll_out=nll)
        caller_idx = torch.from_numpy(order[lo:hi]).to(dev)
        tied = nll == best[0]
        local = best[1:].view(torch.int64)
        pick = torch.where(tied.any(), torch.where(tied, caller_idx, torch.iinfo(torch.int64).max).min(),
                           caller_idx[local.clamp(min=0)][0])
        local.copy_(torch.where(local >= 0, pick, local).reshape(1))
        return reduce_best(best, 0, group=group)
    best = torch.tensor([float("nan"), 0.0], dtype=torch.float64, device=dev)
    best[1:].view(torch.int64).fill_(-1)
    return reduce_best(best, lo, group=group)


def parameter_grid(*axes):
    """Cartesian product of 1-D axes -> tuple of flat ``float64[P]`` arrays (first axis slowest)."""
    mesh = np.meshgrid(*[np.asarray(a, dtype=np.float64) for a in axes], indexing="ij")
    return tuple(np.ascontiguousarray(m.reshape(-1)) for m in mesh)


def shard_bounds(P, rank, world, align=32):
    """Contiguous block ``[lo, hi)`` of the flat parameter index owned by ``rank``.

    Shards start on multiples of ``align`` = 32 parameter sets (when there are at least ``world`` such blocks), so that
    the runs of 32 sets sharing ``(alpha, gamma)`` that the sweeps look for stay aligned inside every shard; sizes then
    differ by less than two blocks."""
    P, world = int(P), int(world)
    blocks = -(-P // align)
    if blocks < world:  # tiny grids: split single parameter sets
        base, rem = divmod(P, world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)
    base, rem = divmod(blocks, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return min(lo * align, P), min(hi * align, P)


class ConditioningSweep:
    """Grid version of ``ConditioningExperiment``: many parameter sets x many sessions on one GPU.

    >>> sw = ConditioningSweep(maze, "positive", possible_actions=A8)
    >>> sw.add_session(trajectory, reward)            # any number of sessions
    >>> ll = sw.log_likelihood(alpha, beta, gamma)    # float64[S, P]
    """

    def __init__(self, maze, algorithm="td0", possible_actions=None, mouse_init=None, replay_n_times=0,
                 replay_seed=None, device=None):
        self.maze = maze
        self.alg = algorithm_code(algorithm)
        self.possible_actions = None if possible_actions is None else np.asarray(possible_actions, dtype=np.float64)
        self.mouse_init = mouse_init  # (x, y, ori) or None -> first trajectory point, orientation 0
        self.replay_n_times = int(replay_n_times)
        self.replay_seed = replay_seed
        self._device = device
        self._records = []
        self._online = []
        self._dmaze = None
        self._dstreams = None
        self._staging = None  # pinned host copy of the concatenated records
        self._external = False  # True: the streams were built on the device (use_streams)

    # -- sessions -------------------------------------------------------------------------------
    def add_session(self, trajectory, reward, session=None):
        """Append one session (or a reference-style ``Session`` via ``session=``)."""
        if session is not None:
            trajectory, reward = session.trajectory, session.reward
        traj = np.asarray(trajectory, dtype=np.float64).reshape(-1, 2)
        mouse = None
        if self.possible_actions is not None:
            init = self.mouse_init if self.mouse_init is not None else (traj[0, 0], traj[0, 1], 0.0)
            mouse = Mouse((init[0], init[1]), init[2], self.possible_actions, save_history=False)
        elif self.alg != _native.ALG_TD0:
            raise ValueError("PositiveConditioning / NegativeConditioning need possible_actions")
        recs = build_session_records(self.maze, traj, reward, mouse=mouse)
        n_on = len(recs)
        if self.replay_n_times > 0:
            seed = self.replay_seed
            recs = np.concatenate([recs, replay_records(recs, self.replay_n_times, seed=seed)])
        self.add_records(recs, n_on)

    def add_records(self, records, n_online=None):
        self._records.append(np.ascontiguousarray(records, dtype=_native.STEP_REC_DTYPE))
        self._online.append(len(records) if n_online is None else int(n_online))
        if self._external:
            raise RuntimeError("this sweep uses device-built streams; host sessions cannot be mixed in")
        self._dstreams = None
        self._staging = None

    def use_streams(self, streams):
        """Use step streams that already live on the device (``SessionBatch.streams``, built by
        ``nm_streams_build_device``) instead of host-built records; replaces every session added so far."""
        if streams.dmaze.maze is not self.maze:
            raise ValueError("the streams were built for another maze")
        if self.alg != _native.ALG_TD0 and not streams.has_candidates:
            raise ValueError("PositiveConditioning / NegativeConditioning need possible_actions")
        self._records, self._online, self._staging = [], [], None
        if self.replay_n_times > 0:  # the sweep's replay strategy applies to device-built sessions too
            streams = streams.with_replays(self.replay_n_times, self.replay_seed)
        self._dmaze = streams.dmaze
        self._dstreams = streams
        self._external = True

    @property
    def n_sessions(self):
        return self._dstreams.S if self._external else len(self._records)

    @property
    def updates_per_parameter_set(self):
        """Number of value updates one parameter set performs over all sessions (online + replays)."""
        return self._dstreams.total if self._external else int(sum(len(r) for r in self._records))

    # -- device state ----------------------------------------------------------------------------
    def _ensure_device(self):
        import torch
        _native.require_device()
        if self._dmaze is None:
            self._dmaze = DeviceMaze(self.maze, self._device)
        if self._dstreams is None:
            if not self._records:
                raise RuntimeError("no session added")
            self._dstreams = DeviceStreams(self._dmaze, self._records, self._online, staging=self._staging)
            self._staging = self._dstreams._host
        return torch

    def upload(self):
        """(Re-)upload the session streams from the pinned staging buffer; returns the bytes copied
        host->device."""
        if self._dstreams is not None:  # same layout: one async H2D copy into the existing allocation
            self._dstreams.reupload()
            return self._dstreams.host_bytes
        self._ensure_device()
        return self._dstreams.host_bytes

    @property
    def torch_device(self):
        self._ensure_device()
        return self._dmaze.torch_device

    def _v_init_tensor(self, torch, v_init, P):
        if v_init is None:
            return None, 0
        v = np.asarray(v_init, dtype=np.float64)
        X, Y = self.maze.shape
        if v.shape == (X, Y):
            return torch.from_numpy(np.ascontiguousarray(v).reshape(-1)).to(self.torch_device), 0
        if v.shape == (P, X, Y):
            return torch.from_numpy(np.ascontiguousarray(v).reshape(P, -1)).to(self.torch_device), 1
        raise ValueError(f"v_init must have shape {(X, Y)} or {(P, X, Y)}")

    # -- device-resident API (tensors in, tensors out; nothing synchronises) ---------------------------
    def v_tables_device(self, alpha_t, gamma_t, v_init_t=None, v_init_per_unit=0, out=None):
        torch = self._ensure_device()
        P = int(alpha_t.numel())
        n_tiles = self._dmaze.n_tiles
        if out is None:
            out = torch.empty((self.n_sessions, P, n_tiles), dtype=torch.float64, device=self.torch_device)
        with torch.cuda.device(self.torch_device):
            _native.check(self._dmaze._lib.nm_vtable_sweep(
                self.alg, self._dmaze.handle, self._dstreams.handle, _native.tensor_ptr(alpha_t),
                _native.tensor_ptr(gamma_t), P, _native.tensor_ptr(v_init_t), int(v_init_per_unit),
                _native.tensor_ptr(out), _native.cuda_stream_ptr()))
        return out

    def log_likelihood_device(self, alpha_t, beta_t, gamma_t, v_init_t=None, v_init_per_unit=0, ll_out=None,
                              v_out=None):
        torch = self._ensure_device()
        P = int(alpha_t.numel())
        if ll_out is None:
            ll_out = torch.empty((self.n_sessions, P), dtype=torch.float64, device=self.torch_device)
        with torch.cuda.device(self.torch_device):
            _native.check(self._dmaze._lib.nm_loglik_sweep(
                self.alg, self._dmaze.handle, self._dstreams.handle, _native.tensor_ptr(alpha_t),
                _native.tensor_ptr(beta_t), _native.tensor_ptr(gamma_t), P, _native.tensor_ptr(v_init_t),
                int(v_init_per_unit), _native.tensor_ptr(ll_out), _native.tensor_ptr(v_out), _native.cuda_stream_ptr()))
        return ll_out

    def last_sweep_mode(self):
        """Kernel the last :meth:`log_likelihood_device` on the current stream used: 1 = a table per warp shared by an
        aligned run of 32 consecutive parameter sets with equal ``(alpha, gamma)`` (order the grid with beta fastest,
        ``parameter_grid(gamma, alpha, beta)``, and give the beta axis a multiple of 32 points to get there); 2 = a
        table per parameter set in device memory (arbitrary sets, any maze size); 0 = a table per parameter set in
        shared memory (``NM_FIT_TABLES=shared``).  Synchronises the stream."""
        torch = self._ensure_device()
        mode = C.c_int(-1)
        with torch.cuda.device(self.torch_device):
            _native.check(self._dmaze._lib.nm_loglik_sweep_mode(_native.cuda_stream_ptr(), C.byref(mode)))
        return mode.value

    def argmin_device(self, ll_t, nll_out=None):
        """Per-shard best fit of ``ll_t`` ``[S, P]`` -> device tensor ``[2]`` int64-punned ``(nll, idx)``."""
        torch = self._ensure_device()
        S, P = int(ll_t.shape[0]), int(ll_t.shape[1])
        best = torch.empty(2, dtype=torch.float64, device=self.torch_device)
        with torch.cuda.device(self.torch_device):
            _native.check(self._dmaze._lib.nm_argmin_nll(_native.tensor_ptr(ll_t), S, P, _native.tensor_ptr(nll_out),
                                                         _native.tensor_ptr(best), _native.cuda_stream_ptr()))
        return best

    # -- numpy API ---------------------------------------------------------------------------------
    def _to_dev(self, torch, a):
        t = torch.from_numpy(a)
        return t.to(self.torch_device, non_blocking=False)

    def v_tables(self, alpha, gamma, v_init=None):
        """V-tables after every session: ``float64[S, P, X, Y]`` (each session starts from ``v_init``)."""
        torch = self._ensure_device()
        alpha = _as_param(alpha)
        gamma = _as_param(gamma, alpha.size)
        alpha = _as_param(alpha, gamma.size)
        P = alpha.size
        if gamma.size != P:
            raise ValueError("alpha and gamma must have the same length")
        vi, per = self._v_init_tensor(torch, v_init, P)
        out = self.v_tables_device(self._to_dev(torch, alpha), self._to_dev(torch, gamma), vi, per)
        X, Y = self.maze.shape
        return out.cpu().numpy().reshape(self.n_sessions, P, X, Y)

    def log_likelihood(self, alpha, beta, gamma, v_init=None, return_v=False, regroup=True):
        """Per-session log-likelihood of the observed choices: ``float64[S, P]``.

        ``regroup``: a value table depends on ``(alpha, gamma)`` only, so sets sharing them are brought side by side
        first when that yields uniform runs of 32 (:func:`regroup_order` -- any Cartesian grid whose beta axis has a
        multiple of 32 points, whatever its axis order); the kernel with one table per warp then applies, and the
        results come back in the caller's order (bit-identical either way)."""
        torch = self._ensure_device()
        alpha, beta, gamma, P = _params3(alpha, beta, gamma)
        if (self._external and not self._dstreams.has_candidates) or \
                any(r["ncand"].min(initial=1) == 0 for r in self._records if len(r)):
            raise ValueError("the log-likelihood needs candidate next states: pass possible_actions")
        vi, per = self._v_init_tensor(torch, v_init, P)
        order = regroup_order(alpha, gamma) if (regroup and not per) else None
        if order is not None:
            alpha, beta, gamma = (np.ascontiguousarray(x[order]) for x in (alpha, beta, gamma))
            back = np.empty_like(order)
            back[order] = np.arange(P)
        unsort = (lambda x: x) if order is None else (lambda x: np.ascontiguousarray(x[:, back]))
        v_out = None
        if return_v:
            v_out = torch.empty((self.n_sessions, P, self._dmaze.n_tiles), dtype=torch.float64,
                                device=self.torch_device)
        ll = self.log_likelihood_device(self._to_dev(torch, alpha), self._to_dev(torch, beta),
                                        self._to_dev(torch, gamma), vi, per, v_out=v_out)
        ll = unsort(ll.cpu().numpy())
        if return_v:
            X, Y = self.maze.shape
            return ll, unsort(v_out.cpu().numpy().reshape(self.n_sessions, P, X, Y))
        return ll

    def best_fit(self, alpha, beta, gamma, v_init=None, group=None, regroup=True):
        """Grid search: minimise the negative log-likelihood summed over sessions (:func:`sharded_best_fit`).

        Scalars broadcast against the longest array as in :meth:`log_likelihood`.  With an initialised
        ``torch.distributed`` process group the flat parameter index is sharded in contiguous blocks over the ranks
        (every rank holds all sessions) and the per-shard best fits ``(nll, index)`` -- 16 bytes per rank -- are
        all-gathered; every rank returns the global arg-min (lowest index on ties, like ``numpy.argmin``).

        :return: ``(best_nll, best_index)``
        """
        torch = self._ensure_device()

        def ll_device(at, bt, gt, vi_np):
            vi, per = self._v_init_tensor(torch, vi_np, int(at.numel()))
            return self.log_likelihood_device(at, bt, gt, vi, per)

        return sharded_best_fit(self, ll_device, alpha, beta, gamma, group=group, v_init=v_init, regroup=regroup)


def reduce_best(best, offset, group=None):
    """All-gather per-shard ``(nll, local idx)`` pairs and pick the global arg-min.

    ``best``: tensor ``float64[2]`` whose second element holds an int64 bit pattern (the kernel's
    ``{double, int64}`` output).  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist
    best = best.clone()
    idx_view = best[1:].view(torch.int64)
    idx_view += torch.where(idx_view >= 0, torch.tensor(int(offset), dtype=torch.int64, device=best.device),
                            torch.tensor(0, dtype=torch.int64, device=best.device))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        gathered = torch.empty(world * 2, dtype=torch.float64, device=best.device)
        dist.all_gather_into_tensor(gathered, best, group=group)
        pairs = gathered.view(world, 2).cpu()
    else:
        pairs = best.view(1, 2).cpu()
    vals = pairs[:, 0].numpy()
    idxs = pairs[:, 1:].contiguous().view(torch.int64).numpy().reshape(-1)
    best_v, best_i = float("nan"), -1
    for v, i in zip(vals, idxs):
        if i < 0 or np.isnan(v):
            continue
        if best_i < 0 or v < best_v or (v == best_v and i < best_i):
            best_v, best_i = float(v), int(i)
    return best_v, best_i


def vtable_single(maze, alg, records, n_online, alpha, gamma, v_init):
    """One parameter set, one stream -> dense ``float64[X, Y]`` table (used by ``ConditioningExperiment``)."""
    sw = ConditioningSweep(maze, alg)
    sw.add_records(records, n_online)
    return sw.v_tables([alpha], [gamma], v_init=v_init)[0, 0]


def vtable_trace_single(maze, alg, records, n_online, alpha, gamma, v_init):
    """One parameter set, one stream through ``nm_vtable_trace``: the dense ``float64[X, Y]`` table AND what the
    learner's ``update`` returns per record -- ``dv`` and ``|rpe|`` (``rl.py:190-194, 260-262``).  ``None`` when the maze
    has more states than the trace kernel's shared-memory table holds (the caller falls back to :func:`vtable_single`)."""
    import torch
    lib = _native.lib()
    hm = host_maze(maze)
    if hm.num_states > lib.nm_vtable_trace_max_states():
        return None
    sw = ConditioningSweep(maze, alg)
    sw.add_records(records, n_online)
    sw._ensure_device()
    dev = sw.torch_device
    X, Y = maze.shape
    n = len(records)
    vi = None if v_init is None else torch.from_numpy(np.ascontiguousarray(v_init, dtype=np.float64).reshape(-1)).to(dev)
    out = torch.empty(X * Y + 2 * n, dtype=torch.float64, device=dev)
    v_t, dv_t, rpe_t = out[: X * Y], out[X * Y: X * Y + n], out[X * Y + n:]
    with torch.cuda.device(dev):
        _native.check(lib.nm_vtable_trace(sw.alg, sw._dmaze.handle, sw._dstreams.handle, 0, float(alpha), float(gamma),
                                          _native.tensor_ptr(vi), _native.tensor_ptr(v_t), _native.tensor_ptr(dv_t),
                                          _native.tensor_ptr(rpe_t), _native.cuda_stream_ptr()))
    host = out.cpu().numpy()
    return host[: X * Y].reshape(X, Y), host[X * Y: X * Y + n], host[X * Y + n:]


class _TileActionSweep:
    """Shared host side of the two tile-action agents (model-based, Q-table): actions are tile displacements, the
    transition model ``next[s, a]`` is the maze's own geometry (``Maze.is_movement_possible``, ``maze.py:512-527``),
    the session streams only carry ``(s, s1, r)``."""

    def __init__(self, maze, discrete_actions, device=None, replay_n_times=0, replay_seed=None):
        self.maze = maze
        self.discrete_actions = np.asarray(discrete_actions, dtype=np.int64).reshape(-1, 2)
        if not 1 <= len(self.discrete_actions) <= _native.NM_MAX_CAND:
            raise ValueError(f"between 1 and {_native.NM_MAX_CAND} discrete actions are supported")
        self.next_table = np.ascontiguousarray(maze.transition_table(self.discrete_actions), dtype=np.int16)
        # session streams: s, s1, r only
        self._base = ConditioningSweep(maze, "td0", device=device, replay_n_times=replay_n_times,
                                       replay_seed=replay_seed)
        self.last_kernel_ms = None

    def add_session(self, trajectory, reward):
        self._base.add_session(trajectory, reward)

    def use_streams(self, streams):
        """Use step streams built on the device (``SessionBatch.streams``) instead of host-built records."""
        self._base.use_streams(streams)

    @property
    def n_sessions(self):
        return self._base.n_sessions

    @property
    def updates_per_parameter_set(self):
        return self._base.updates_per_parameter_set

    _params = staticmethod(_params3)
    regroup_order = staticmethod(regroup_order)

    def _timed(self, torch, launch):
        """Run ``launch()`` (one C-ABI call) between two CUDA events on the current stream."""
        with torch.cuda.device(self._base.torch_device):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _native.check(launch())
            e1.record()
            torch.cuda.current_stream().synchronize()
            self.last_kernel_ms = e0.elapsed_time(e1)

    def log_likelihood_device(self, alpha_t, beta_t, gamma_t):
        raise NotImplementedError

    _mode_symbol = None  # C-ABI diagnostic of the subclass

    def last_sweep_mode(self):
        """Mapping the last sweep on the current stream used: 0 = tables per parameter set, 1 = tables per aligned run
        of 32 consecutive parameter sets with equal ``(alpha, gamma)`` -- the tables do not depend on beta, so on a
        grid with beta fastest (``parameter_grid(gamma, alpha, beta)``, a multiple of 32 beta points) they are kept and
        updated once per run.  Decided on the device, bit-identical results.  Synchronises the stream."""
        b = self._base
        torch = b._ensure_device()
        mode = C.c_int(-1)
        with torch.cuda.device(b.torch_device):
            _native.check(getattr(b._dmaze._lib, self._mode_symbol)(_native.cuda_stream_ptr(), C.byref(mode)))
        return mode.value

    def best_fit(self, alpha, beta, gamma, group=None, regroup=True):
        """Grid search like :meth:`ConditioningSweep.best_fit` (:func:`sharded_best_fit`): sets re-ordered into runs
        of 32 sharing ``(alpha, gamma)`` when possible, the flat index sharded over the ranks of an initialised
        process group, 16 bytes per rank all-gathered.  ``(best_nll, best_index)`` in the caller's order."""
        self._base._ensure_device()
        return sharded_best_fit(self._base, lambda at, bt, gt, _vi: self.log_likelihood_device(at, bt, gt),
                                alpha, beta, gamma, group=group, regroup=regroup)


class ModelBasedSweep(_TileActionSweep):
    """Grid sweep of the model-based agent (EXTENSION -- the reference has no model-based agent; semantics in
    DESIGN.md section 2): learned reward model + ``n_sweeps`` value-iteration sweeps per observed step on the maze's
    transition table, softmax likelihood of the observed tile steps.  One warp per parameter set.

    >>> mb = ModelBasedSweep(maze, [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)])
    >>> mb.add_session(trajectory, reward)
    >>> ll = mb.log_likelihood(alpha, beta, gamma)      # float64[S, P]
    """

    _mode_symbol = "nm_modelbased_sweep_mode"

    def __init__(self, maze, discrete_actions, n_sweeps=1, device=None):
        super().__init__(maze, discrete_actions, device=device)
        self.n_sweeps = int(n_sweeps)

    def log_likelihood_device(self, alpha_t, beta_t, gamma_t, v_init_t=None, ll_out=None, v_out=None, r_out=None):
        b = self._base
        torch = b._ensure_device()
        dev = b.torch_device
        P = int(alpha_t.numel())
        nxt = torch.from_numpy(self.next_table).to(dev)
        if ll_out is None:
            ll_out = torch.empty((b.n_sessions, P), dtype=torch.float64, device=dev)
        self._timed(torch, lambda: b._dmaze._lib.nm_modelbased_sweep(
            b._dmaze.handle, b._dstreams.handle, _native.tensor_ptr(nxt), self.next_table.shape[1], self.n_sweeps,
            _native.tensor_ptr(alpha_t), _native.tensor_ptr(beta_t), _native.tensor_ptr(gamma_t), P,
            _native.tensor_ptr(v_init_t), _native.tensor_ptr(ll_out), _native.tensor_ptr(v_out),
            _native.tensor_ptr(r_out), _native.cuda_stream_ptr()))  # last_kernel_ms: device time of nm_mb_kernel
        return ll_out

    def log_likelihood(self, alpha, beta, gamma, v_init=None, return_tables=False, regroup=True):
        """``float64[S, P]`` log-likelihoods (and, optionally, the final ``V`` and ``Rhat`` tables ``[S, P, X, Y]``).
        ``regroup``: bring sets with equal ``(alpha, gamma)`` together first when that yields uniform runs of 32
        (:meth:`regroup_order`); results come back in the caller's order."""
        b = self._base
        torch = b._ensure_device()
        alpha, beta, gamma, P = self._params(alpha, beta, gamma)
        order = self.regroup_order(alpha, gamma) if regroup else None
        if order is not None:  # sets with equal (alpha, gamma) side by side: tables and sweeps once per run of 32
            alpha, beta, gamma = (np.ascontiguousarray(x[order]) for x in (alpha, beta, gamma))
            back = np.empty_like(order)
            back[order] = np.arange(P)
        dev = b.torch_device
        X, Y = self.maze.shape
        vi = None
        if v_init is not None:
            v = np.asarray(v_init, dtype=np.float64)
            if v.shape != (X, Y):
                raise ValueError(f"v_init must have shape {(X, Y)}")
            vi = torch.from_numpy(np.ascontiguousarray(v).reshape(-1)).to(dev)
        S = b.n_sessions
        v_out = r_out = None
        if return_tables:
            v_out = torch.empty((S, P, X * Y), dtype=torch.float64, device=dev)
            r_out = torch.empty((S, P, X * Y), dtype=torch.float64, device=dev)
        at, bt, gt = (b._to_dev(torch, x) for x in (alpha, beta, gamma))
        ll = self.log_likelihood_device(at, bt, gt, vi, v_out=v_out, r_out=r_out)
        unsort = (lambda x: x) if order is None else (lambda x: np.ascontiguousarray(x[:, back]))
        if return_tables:
            return (unsort(ll.cpu().numpy()), unsort(v_out.cpu().numpy().reshape(S, P, X, Y)),
                    unsort(r_out.cpu().numpy().reshape(S, P, X, Y)))
        return unsort(ll.cpu().numpy())


class QLearningSweep(_TileActionSweep):
    """Grid sweep of a tabular Q-learning agent with an explicit ``Q[s, a]`` table (EXTENSION -- the reference's
    learners keep state values only, ``rl.py:154-262``; semantics in DESIGN.md section 2): softmax likelihood of the
    observed tile steps under ``beta * Q[s, .]`` and the update ``Q[s,a] += alpha * (r + gamma * max_a' Q[s1,a'] -
    Q[s,a])``.  Eight lanes per parameter set, the table in shared memory.  ``replay_n_times`` shuffled replay passes
    (``ShuffledReplays``, ``rl.py:419-424``) update ``Q`` after the online pass.

    >>> ql = QLearningSweep(maze, [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)])
    >>> ql.add_session(trajectory, reward)
    >>> ll = ql.log_likelihood(alpha, beta, gamma)      # float64[S, P]
    """

    _mode_symbol = "nm_qtable_sweep_mode"

    def log_likelihood_device(self, alpha_t, beta_t, gamma_t, q_init_t=None, ll_out=None, q_out=None):
        b = self._base
        torch = b._ensure_device()
        dev = b.torch_device
        P = int(alpha_t.numel())
        nxt = torch.from_numpy(self.next_table).to(dev)
        if ll_out is None:
            ll_out = torch.empty((b.n_sessions, P), dtype=torch.float64, device=dev)
        self._timed(torch, lambda: b._dmaze._lib.nm_qtable_sweep(
            b._dmaze.handle, b._dstreams.handle, _native.tensor_ptr(nxt), self.next_table.shape[1],
            _native.tensor_ptr(alpha_t), _native.tensor_ptr(beta_t), _native.tensor_ptr(gamma_t), P,
            _native.tensor_ptr(q_init_t), _native.tensor_ptr(ll_out), _native.tensor_ptr(q_out),
            _native.cuda_stream_ptr()))  # last_kernel_ms: device time of nm_q_kernel
        return ll_out

    def log_likelihood(self, alpha, beta, gamma, q_init=None, return_tables=False, regroup=True):
        """``float64[S, P]`` log-likelihoods (and, optionally, the final tables ``Q[S, P, X, Y, A]``; entries of
        non-visitable tiles and unavailable actions keep their initial value).  ``regroup``: bring sets with equal
        ``(alpha, gamma)`` together first when that yields uniform runs of 32 (:meth:`regroup_order`); results come back
        in the caller's order."""
        b = self._base
        torch = b._ensure_device()
        alpha, beta, gamma, P = self._params(alpha, beta, gamma)
        order = self.regroup_order(alpha, gamma) if regroup else None
        if order is not None:
            alpha, beta, gamma = (np.ascontiguousarray(x[order]) for x in (alpha, beta, gamma))
            back = np.empty_like(order)
            back[order] = np.arange(P)
        dev = b.torch_device
        X, Y = self.maze.shape
        A = self.next_table.shape[1]
        qi = None
        if q_init is not None:
            q = np.asarray(q_init, dtype=np.float64)
            if q.shape != (X, Y, A):
                raise ValueError(f"q_init must have shape {(X, Y, A)}")
            qi = torch.from_numpy(np.ascontiguousarray(q).reshape(-1)).to(dev)
        S = b.n_sessions
        q_out = torch.empty((S, P, X * Y * A), dtype=torch.float64, device=dev) if return_tables else None
        at, bt, gt = (b._to_dev(torch, x) for x in (alpha, beta, gamma))
        ll = self.log_likelihood_device(at, bt, gt, qi, q_out=q_out)
        unsort = (lambda x: x) if order is None else (lambda x: np.ascontiguousarray(x[:, back]))
        if return_tables:
            return unsort(ll.cpu().numpy()), unsort(q_out.cpu().numpy().reshape(S, P, X, Y, A))
        return unsort(ll.cpu().numpy())
