"""Batched forward simulation: many mice (parameter sets x seeds) in one kernel launch.

Host layer above ``nm_simulate`` (``csrc/nm_sim.cu``): the grid version of the reference's
``StandardExperiment(BehaviouralModel(components, ...), maze, Mouse(...)).run(iterations)``
(``simulation/experiment.py:23-51``).  The model is described once (ordered component classes, the action
set, the maze); every free parameter of ``BehaviouralModel.parameters`` (``beta`` + the components'
descriptors, ``exploration_model.py:55-66``) becomes an array with one entry per mouse.

Randomness: either ``seeds`` (each mouse draws from ``numpy.random.default_rng(seed)``'s PCG64 stream,
generated on the device from the host-computed state) or ``draws[N, T]`` (replayed uniforms -- what the
parity tests use).  PyTorch is used for device buffers only.  No CPU fallback.
"""
import ctypes as C
import functools
import os
import threading
from dataclasses import dataclass

import numpy as np

from . import _native
from ._hostpool import pool as _pool
from .simulation import behavioural_components as bc
from .sweep import DeviceMaze

_NP_OF_TORCH = {"torch.float64": np.float64, "torch.float32": np.float32, "torch.int64": np.int64,
                "torch.int32": np.int32, "torch.int16": np.int16, "torch.uint8": np.uint8, "torch.int8": np.int8}


def pcg64_state_words(seeds):
    """``uint64[N, 4]`` = (state_hi, state_lo, inc_hi, inc_lo) of ``numpy.random.default_rng(seed)`` for every
    seed (or of already-constructed ``Generator`` objects)."""
    out = np.empty((len(seeds), 4), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, s in enumerate(seeds):
        gen = s if isinstance(s, np.random.Generator) else np.random.default_rng(None if s is None else int(s))
        bg = gen.bit_generator
        if type(bg).__name__ != "PCG64":
            raise ValueError("only PCG64 generators can be continued on the device")
        st = bg.state["state"]
        out[i] = (st["state"] >> 64, st["state"] & mask, st["inc"] >> 64, st["inc"] & mask)
    return out


def pcg64_buffered_u32(seeds):
    """``uint32[N, 2]`` = numpy's buffered half of a 64-bit output ``(has_uint32, uinteger)`` per generator
    (all zeros for fresh ``default_rng(seed)`` generators)."""
    out = np.zeros((len(seeds), 2), dtype=np.uint32)
    for i, s in enumerate(seeds):
        if isinstance(s, np.random.Generator):
            st = s.bit_generator.state
            out[i] = (st["has_uint32"], st["uinteger"])
    return out


def set_pcg64_state(gen, words, buffered=(0, 0)):
    """Write a ``(state_hi, state_lo, inc_hi, inc_lo)`` quadruple (and the buffered 32-bit half) back into a
    numpy ``Generator``."""
    st = gen.bit_generator.state
    st["state"] = {"state": (int(words[0]) << 64) | int(words[1]), "inc": (int(words[2]) << 64) | int(words[3])}
    st["has_uint32"], st["uinteger"] = int(buffered[0]), int(buffered[1])
    gen.bit_generator.state = st


@dataclass
class SimResult:
    """Outputs of :meth:`BatchedExperiment.run` (numpy arrays; ``None`` when not requested)."""
    final_pose: np.ndarray            # [N, 3]
    status: np.ndarray                # [N] 0 ok, t+1 when step t had no reachable endpoint
    it_visit: np.ndarray              # [N] it_perc_visited_maze(percentage)
    count_moving: np.ndarray          # [N] motion_analysis
    tile_counts: np.ndarray           # [N, 8] tile_analysis counts by tile code
    occupancy: np.ndarray = None      # [N, X, Y] occupancy_map
    occupancy_total: np.ndarray = None  # [X, Y] summed over mice
    history_position: np.ndarray = None     # [N, T+1, 2]
    history_orientation: np.ndarray = None  # [N, T+1]
    choices: np.ndarray = None        # [N, T] chosen action index
    pcg_state: np.ndarray = None      # [N, 4] generator state after the run
    model_state: np.ndarray = None    # [N, 5] prev_pos.x, prev_pos.y, moving, familiarity, ment_state
    pcg_buffered: np.ndarray = None   # [N, 2] numpy's buffered 32-bit half (has_uint32, uinteger)
    # fused trajectory metrics (analysis/metrics.py), each None unless requested in run()
    zone_latency: np.ndarray = None         # [N] latency_enter_shock_zone
    zone_occupancy: np.ndarray = None       # [N] positions inside the shock zone (occupancy_shock_zone)
    zone_entries: np.ndarray = None         # [N] shock_zone_entries
    corridor_switches: np.ndarray = None    # [N] arms
    static_intervals: np.ndarray = None     # [N] compute_static_intervals(hist_time_moving(moving_bool_list))
    subarea_counts: np.ndarray = None       # [N, 16] positions by subarea index
    orientations_histogram: np.ndarray = None       # [N, n_bins (+1 with distance=True)] hist_orientations(_filtered)
    orientations_histogram_bins: np.ndarray = None  # [n_bins + 1]
    n_positions: int = 0                    # history length, n_steps + 1

    def as_results(self, maze):
        """The fused metrics under the product names of the reference's ``@metric`` functions (``analysis/metrics.py``),
        one row per mouse: what ``BatchedMultiObjectiveFitness`` / ``MultiObjectiveFitness`` objectives key on."""
        res = {"count_moving": self.count_moving, "it_2_visit_perc_maze": self.it_visit,
               "moving_dynamics_histogram": self.mov_bouts(), "tiles_histogram": self.tiles_histogram(maze)}
        if self.occupancy is not None:
            res["occupancy"] = self.occupancy
        if self.static_intervals is not None:
            res["static_intervals"] = self.static_intervals
        if self.subarea_counts is not None:
            res["subareas_histogram"] = self.subareas_histogram(maze)
        if self.zone_entries is not None:
            res.update(latency_enter_shock_zone=self.zone_latency, occupancy_shock_zone=self.zone_occupancy,
                       shock_zone_entries=self.zone_entries)
        if self.corridor_switches is not None:
            res["corridors_switches"] = self.corridor_switches
        if self.orientations_histogram is not None:
            res["orientations_histogram"] = self.orientations_histogram
            if self.orientations_histogram.shape[1] >= 7:
                res["rel_ori_signature_histogram"] = self.rel_ori_signature()
        return res

    def mov_bouts(self):
        """``moving_dynamics_histogram`` of every mouse, ``[N, 2]`` = (steps in moving bouts, steps in not-moving bouts)
        (``metrics.py:589-616``): the moved flags are True ``count_moving`` times out of ``n_steps + 1``."""
        return np.stack([self.count_moving, self.n_positions - self.count_moving], axis=1)

    def subareas_histogram(self, maze):
        """``[N, max - min + 1]`` like ``metrics.subareas`` (``:470-485``)."""
        ids = maze.subareas
        return self.subarea_counts[:, min(ids):max(ids) + 1]

    def tiles_histogram(self, maze):
        """``[N, 3]`` visits per tile of type (CORNER, WALL, CENTRE) (``histogram_tiles``, ``:251-277``)."""
        T, totals = maze.MazeTile, maze.description()
        return np.stack([self.tile_counts[:, int(t)] / totals[t] if totals[t] > 0 else np.zeros(len(self.tile_counts))
                         for t in (T.CORNER, T.WALL, T.CENTRE)], axis=1)

    def rel_ori_signature(self):
        """``[N, 3]`` contrasts of the eight-bin turn histogram (``relative_orientation_signatures``, ``:362-380``)."""
        h = self.orientations_histogram
        lateral = np.median(h[:, [0, 1, 3, 4]], axis=1)
        return np.stack([h[:, 2] - lateral, h[:, 2] - h[:, 5], h[:, 6] - h[:, 5]], axis=1)

    def tile_analysis(self, maze, n_positions):
        """``(perc, ratio, count)`` dicts keyed by tile type with one array entry per mouse
        (``analysis/metrics.py:219-248``)."""
        totals = maze.description()
        count = {t: self.tile_counts[:, int(t)] for t in maze.MazeTile}
        perc = {t: count[t] * 100 / n_positions for t in count}
        ratio = {t: (count[t] / totals[t] if totals[t] > 0 else np.zeros(len(self.tile_counts))) for t in count}
        return perc, ratio, count


class BatchedExperiment:
    """``N`` independent mice with the same model structure and per-mouse parameters.

    >>> exp = BatchedExperiment(maze, actions, discrete_actions,
    ...                         [AnxietyComponent, BiomechanicalCostComponent, ExperienceComponent,
    ...                          BiomechPersistenceComponent, ValueTableComponent], v_table=V)
    >>> res = exp.run(params, n_steps=10_000, init_pos=p0, init_ori=0.0, init_disc=(1, 1), seeds=range(N))
    """

    def __init__(self, maze, possible_actions, discrete_actions, components, v_table=None, active=None,
                 device=None, v_table_normalised=False, model="behavioural"):
        """
        :param model: ``"behavioural"`` (softmax over the components, ``BehaviouralModel``) or ``"random"``
            (uniform ``rng.choice`` among the reachable endpoints, ``RandomModel``; ``components`` must be empty)
        :param possible_actions: ``[A, 2]`` relative displacements of the Mouse (continuous units)
        :param discrete_actions: ``[A, 2]`` the action list the biomechanical components are built from
        :param components: ordered component classes (or instances, whose ``active`` flag is honoured)
        :param v_table: raw value table(s) for ``ValueTableComponent``: ``[X, Y]`` shared or ``[N, X, Y]``
        """
        if model not in ("behavioural", "random"):
            raise ValueError("model must be 'behavioural' or 'random'")
        self.model_kind = 0 if model == "behavioural" else 1
        if self.model_kind == 1 and len(components):
            raise ValueError("the random model has no components")
        self.maze = maze
        self.actions = np.ascontiguousarray(np.asarray(possible_actions, dtype=np.float64).reshape(-1, 2))
        self.discrete_actions = np.asarray(discrete_actions)
        if len(self.actions) != len(self.discrete_actions):
            raise ValueError("possible_actions and discrete_actions must have the same length")
        if not 1 <= len(self.actions) <= _native.NM_MAX_CAND:
            raise ValueError(f"between 1 and {_native.NM_MAX_CAND} actions are supported")
        self.zero_action = np.ascontiguousarray(np.all(self.discrete_actions == (0, 0), axis=1).astype(np.uint8))
        self.kinds, self.active, self.classes = [], [], []
        for i, c in enumerate(components):
            cls = c if isinstance(c, type) else type(c)
            kind = getattr(cls, "device_kind", None)
            if kind is None:
                raise ValueError(f"{cls.__name__} has no device implementation (host plugin components run through "
                                 "StandardExperiment)")
            self.kinds.append(kind)
            self.classes.append(cls)
            self.active.append(bool(active[i]) if active is not None else (True if isinstance(c, type) else c.active))
        if len(self.kinds) > _native.NM_MAX_COMPONENTS:
            raise ValueError("too many components")
        if len(set(self.kinds)) != len(self.kinds):
            raise ValueError("each component kind may appear once")
        self.v_table = None if v_table is None else np.asarray(v_table, dtype=np.float64)
        if _native.COMP_VTABLE in self.kinds and self.v_table is None:
            raise ValueError("ValueTableComponent needs v_table")
        self._device = device
        self._dmaze = None
        self.last_kernel_ms = None
        self._run_lock = threading.RLock()  # one run at a time: the staging buffers below belong to the object
        self._arenas = []                  # pinned upload arenas [uint8 tensor, bytes used], kept across runs
        self._pinned = None                # pinned staging buffer of the per-mouse results, kept across runs
        self._workers = None               # run_pipelined: two shallow copies with their own staging buffers + streams
        # True: v_table already holds ValueTableComponent._v_table (normalised once at construction)
        self._normalise = (lambda t: np.asarray(t, dtype=np.float64)) if v_table_normalised else bc.normalise_v_table

    @property
    def parameter_names(self):
        """Names of the per-mouse parameters, in ``BehaviouralModel.parameters`` order."""
        if self.model_kind == 1:
            return []  # RandomModel.parameters == {}
        names = ["beta"]
        for cls in self.classes:
            names += [d.name for d in cls.get_descriptors()]
        return names

    def parameter_bounds(self):
        """``{name: (min, max, default)}`` from the component descriptors (the sweep's search space)."""
        out = {"beta": (0.0, np.inf, None)}
        for cls in self.classes:
            for d in cls.get_descriptors():
                out[d.name] = (d.minv, d.maxv, d.default)
        return out

    def _validate(self, params, N):
        cols = {}
        bounds = self.parameter_bounds()
        for name in self.parameter_names:
            if name not in params:
                raise RuntimeError(f"Value for {name} has not been specified")
            a = np.asarray(params[name], dtype=np.float64).reshape(-1)
            if a.size == 1 and N != 1:
                a = np.full(N, a[0])
            if a.size != N:
                raise ValueError(f"parameter {name}: expected {N} values, got {a.size}")
            cols[name] = a

        def out_of_range(name):  # first offending index of a column, or -1 (NaN counts as out of range)
            lo, hi, _ = bounds[name]
            bad = ~((cols[name] >= lo) & (cols[name] <= hi))
            return int(np.flatnonzero(bad)[0]) if bad.any() else -1

        names = [n for n in bounds if n != "beta"]
        # a million mice x 18 columns: the comparisons run column by column on a few threads (numpy drops the GIL)
        first = list(_pool().map(out_of_range, names)) if N > (1 << 16) else [out_of_range(n) for n in names]
        for name, i in zip(names, first):
            if i >= 0:
                lo, hi, _ = bounds[name]
                raise RuntimeError(f"Wrong value set for property '{name}' ({cols[name][i]}), must be in range "
                                   f"[{lo}; {hi}]")
        return cols

    # -- host -> device: large arrays go through pinned memory kept across runs ----------------------------------------
    def _stage_reset(self):
        """Start of a run: the pinned arenas of the previous run are free again (every run ends synchronised)."""
        for ar in self._arenas:
            ar[1] = 0

    def _stage_slot(self, torch, nbytes):
        need = (nbytes + 63) & ~63
        for ar in self._arenas:
            if ar[0].numel() - ar[1] >= need:
                off = ar[1]
                ar[1] += need
                return ar[0][off:off + nbytes]
        self._arenas.append([torch.empty(max(need, 64 << 20), dtype=torch.uint8, pin_memory=True), need])
        return self._arenas[-1][0][:nbytes]

    def _upload_f64(self, torch, a, dev):
        """``a`` (any dtype / strides) as a float64 device tensor.  A pageable ``.to(device)`` is a synchronous copy at
        ~10 GB/s; above 1 MB the array is written into a pinned arena in row blocks over a few threads (conversion and
        transposition happen in that pass) and leaves by an asynchronous copy that overlaps the next array's pass."""
        a = np.asarray(a)
        if a.nbytes < (1 << 20) or a.ndim == 0:
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        slot = self._stage_slot(torch, a.size * 8)
        dst = slot.numpy().view(np.float64).reshape(a.shape)
        rows = max(1, (4 << 20) // max(1, 8 * a.size // a.shape[0]))
        spans = [(lo, min(lo + rows, a.shape[0])) for lo in range(0, a.shape[0], rows)]
        if len(spans) == 1:
            np.copyto(dst, a)
        else:
            list(_pool().map(lambda sp: np.copyto(dst[sp[0]:sp[1]], a[sp[0]:sp[1]]), spans))
        return slot.view(torch.float64).view(a.shape).to(dev, non_blocking=True)

    def _fetch(self, torch, tensors):
        """Device tensors -> fresh numpy arrays through a pinned staging buffer that is kept across runs (pageable
        ``.cpu()`` copies run at a few GB/s and synchronise one by one).  ``{id(tensor): array}``.

        The staging buffer is reused by the next run, so the results are copied out of it: into ONE fresh block (the
        returned arrays are views of it), in slices spread over a few threads -- a single-threaded copy of the 240 MB
        a million mice return, page faults of the fresh block included, took 60 ms of a 280 ms run."""
        sizes = [(t.numel() * t.element_size() + 63) & ~63 for t in tensors]
        total = sum(sizes)
        pin = self._pinned
        if pin is None or pin.numel() < total:
            pin = self._pinned = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True)
        offs, off = [], 0
        for t, nbytes in zip(tensors, sizes):
            pin[off:off + t.numel() * t.element_size()].view(t.dtype).view(t.shape).copy_(t, non_blocking=True)
            offs.append(off)
            off += nbytes
        torch.cuda.current_stream().synchronize()
        src = pin.numpy()
        if getattr(self, "_fetch_views", False):
            # a run_pipelined worker: the caller copies the rows straight from the staging buffer into the population's
            # arrays before this worker runs again (one host copy instead of two)
            out = {}
            for t, o in zip(tensors, offs):
                a = src[o:o + t.numel() * t.element_size()].view(_NP_OF_TORCH[str(t.dtype)])
                out[id(t)] = a.reshape(tuple(t.shape))
            return out
        block = np.empty(total, dtype=np.uint8)
        chunk = 8 << 20
        if total <= 2 * chunk:
            block[:] = src[:total]
        else:
            spans = [(lo, min(lo + chunk, total)) for lo in range(0, total, chunk)]
            list(_pool().map(lambda sp: np.copyto(block[sp[0]:sp[1]], src[sp[0]:sp[1]]), spans))
        out = {}
        for t, o in zip(tensors, offs):
            a = block[o:o + t.numel() * t.element_size()].view(_NP_OF_TORCH[str(t.dtype)])
            out[id(t)] = a.reshape(tuple(t.shape))
        return out

    def _run(self, params, n_steps, init_pos, init_ori=0.0, init_disc=None, seeds=None, draws=None,
            visit_percentage=80, history=False, occupancy=False, occupancy_total=False, generators=None,
            device_seed=None, n_mice=None, shock_zone=None, corridors=None, static_intervals=False, subareas=False,
            orientation_histogram=None, init_visits=None, init_model_state=None, mouse_index_base=0):
        """Simulate ``n_steps`` iterations for every mouse.

        :param params: ``{name: array[N] or scalar}`` for every name in :attr:`parameter_names`
        :param init_pos: ``[2]`` or ``[N, 2]``; ``init_ori``: scalar or ``[N]``
        :param init_disc: ``init_pos_discrete`` handed to ``BehaviouralModel`` (default: tile of ``init_pos``)
        :param seeds: ``[N]`` seeds of ``default_rng`` -- or ``draws[N, n_steps]`` replayed uniforms -- or
            ``device_seed`` (with ``n_mice``): every mouse seeds its own PCG64 stream on the device from
            ``(device_seed, mouse index)``; the fast path for millions of mice (not numpy's seeding)
        :param history: also return the full position / orientation / choice histories
        :param shock_zone: ``{"subareas": int or ints}`` or ``{"tile_lines": int}`` -- fuse ``latency_enter_shock_zone``,
            ``occupancy_shock_zone`` and ``shock_zone_entries`` (``analysis/metrics.py:488-586``)
        :param corridors: pair of subarea indices (the reference's ``arms`` uses ``(1, 7)``) -- fuse the switch count
        :param static_intervals: fuse ``compute_static_intervals`` of the moved flags
        :param subareas: fuse the per-subarea visit counts (``metrics.subareas``)
        :param orientation_histogram: ``{}`` or a dict with any of ``n_bins``, ``max_lim``, ``possible_actions``,
            ``distance``, ``tiles_filter`` (tile types: ``hist_orientations_filtered``) -- fuse ``hist_orientations``
            of the mouse's own orientation history
        :param init_visits: ``[X, Y]`` or ``[N, X, Y]`` non-negative integers -- the ``ExperienceComponent``'s visit map
            left by an earlier run (``StandardExperiment.run`` called again on the same objects).  The kernel's one
            visit map starts from it: ``occupancy`` and the metrics counted from that map (``tile_counts``,
            ``it_visit``, zone / subarea occupancy) then include the seeded visits
        :param init_model_state: ``[3]`` or ``[N, 3]`` -- ``BehaviouralModel._moving`` (0 / 1), the component's
            ``_familiarity`` and ``_ment_state`` (0 EXPLO / 1 HOME) left by that run
        :param mouse_index_base: with ``device_seed``: mouse ``n`` of this call is mouse ``mouse_index_base + n`` of a
            larger population and draws that mouse's stream (:meth:`run_sharded`)
        """
        import torch
        _native.require_device()
        if device_seed is not None:
            if seeds is not None or draws is not None or generators is not None or n_mice is None:
                raise ValueError("device_seed needs n_mice and excludes seeds / draws / generators")
            N = int(n_mice)
        elif (seeds is None) == (draws is None) and generators is None:
            raise ValueError("give exactly one of seeds, draws and device_seed")
        if device_seed is not None:
            pass
        elif draws is not None:
            draws = np.asarray(draws, dtype=np.float64)
            N = draws.shape[0]
            if draws.shape != (N, n_steps):
                raise ValueError("draws must have shape [N, n_steps]")
        elif generators is not None:
            N = len(generators)
        else:
            seeds = list(seeds)
            N = len(seeds)
        cols = self._validate(params, N)
        if self._dmaze is None:
            self._dmaze = DeviceMaze(self.maze, self._device)
        dev = self._dmaze.torch_device
        X, Y = self.maze.shape
        nT = X * Y
        T = int(n_steps)

        pose = np.empty((N, 3))
        pose[:, :2] = np.asarray(init_pos, dtype=np.float64).reshape(-1, 2)
        pose[:, 2] = np.asarray(init_ori, dtype=np.float64).reshape(-1)
        # the kernel evaluates cos / sin / atan2 with the host C library's algorithm (csrc/nm_libm.h), whose table-driven
        # range covers |heading| < 105414350 and finite coordinates; anything else is refused here, not approximated
        if not np.isfinite(pose).all() or (np.abs(pose[:, 2]) >= 105414350.0).any():
            raise ValueError("initial poses must be finite, with |orientation| < 105414350 rad")
        if init_disc is None:
            prev = self.maze.cont2disc_array(pose[:, :2]).astype(np.float64)
        else:
            prev = np.empty((N, 2))
            prev[:] = np.asarray(init_disc, dtype=np.float64).reshape(-1, 2)

        keep = []  # device tensors must outlive the launch

        self._stage_reset()

        def dev_f64(a):
            t = self._upload_f64(torch, a, dev)
            keep.append(t)
            return t

        def dev_cols(*cols_):
            """``[N, k]`` parameter block from ``k`` per-mouse columns: the columns are uploaded as they are and
            interleaved on the device (a host ``np.stack`` of a million rows costs more than the copy)."""
            t = torch.stack([self._upload_f64(torch, c, dev) for c in cols_], dim=1)
            keep.append(t)
            return t

        def out(shape, dtype, zero=False):
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)
            keep.append(t)
            return t

        args = _native.SimArgs()
        args.n_mice, args.n_steps, args.n_actions = N, T, len(self.actions)
        args.h_actions = self.actions.ctypes.data
        args.h_zero_action = self.zero_action.ctypes.data
        args.n_components = len(self.kinds)
        for i, (k, act) in enumerate(zip(self.kinds, self.active)):
            args.component_kind[i] = k
            args.component_active[i] = int(act)
        args.model_kind = self.model_kind
        args.d_beta = dev_f64(cols["beta"] if self.model_kind == 0 else np.zeros(N)).data_ptr()
        args.d_init_pose = dev_f64(pose).data_ptr()
        args.d_init_prev = dev_f64(prev).data_ptr()
        if _native.COMP_ANXIETY in self.kinds:
            p4 = cols["p4"] if "p4" in cols else np.zeros(N)
            args.d_anxiety = dev_cols(cols["W_anx"], cols["p1"], cols["p2"], cols["p3"], p4).data_ptr()
        if _native.COMP_BCOST in self.kinds:
            args.d_bcost_weight = dev_f64(cols["W_bcost"]).data_ptr()
            # the tables are computed straight into pinned memory and leave from there
            slot = self._stage_slot(torch, N * len(self.actions) * 8)
            bc.bcost_tables(self.discrete_actions, cols["k"], cols["gamma"],
                            out=slot.numpy().view(np.float64).reshape(N, len(self.actions)))
            keep.append(slot.view(torch.float64).view(N, len(self.actions)).to(dev, non_blocking=True))
            args.d_bcost_table = keep[-1].data_ptr()
        if _native.COMP_EXPERIENCE in self.kinds:
            args.d_experience = dev_cols(*[cols[n] for n in ("W_exp", "xi", "kappa_home", "kappa_explo",
                                                              "theta_explo", "theta_home")]).data_ptr()
        if _native.COMP_BPERSISTENCE in self.kinds:
            args.d_bpersistence = dev_cols(*[cols[n] for n in ("W_bper", "bp", "Wm", "Wnm")]).data_ptr()
        if _native.COMP_VTABLE in self.kinds:
            args.d_vtable_weight = dev_f64(cols["W_values"]).data_ptr()
            v = self.v_table
            if v.shape == (X, Y):
                args.d_vtable = dev_f64(self._normalise(v).reshape(-1)).data_ptr()
                args.vtable_per_mouse = 0
            elif v.shape == (N, X, Y):
                args.d_vtable = dev_f64(np.stack([self._normalise(t) for t in v]).reshape(N, -1)).data_ptr()
                args.vtable_per_mouse = 1
            else:
                raise ValueError(f"v_table must have shape {(X, Y)} or {(N, X, Y)}")
        if device_seed is not None:
            args.device_seed = int(device_seed) & ((1 << 64) - 1)
            args.use_device_seed = 1
            args.mouse_index_base = int(mouse_index_base)
        elif draws is not None:
            args.d_draws = dev_f64(draws.T).data_ptr()  # time-major [T, N]: coalesced per step
        else:
            gens = generators if generators is not None else seeds
            words = pcg64_state_words(gens)
            t = torch.from_numpy(words.view(np.int64)).to(dev)
            keep.append(t)
            args.d_pcg_state = t.data_ptr()
            t32 = torch.from_numpy(pcg64_buffered_u32(gens).view(np.int32)).to(dev)
            keep.append(t32)
            args.d_pcg_u32 = t32.data_ptr()
        args.visit_percentage = float(visit_percentage)
        if init_visits is not None:
            iv = np.asarray(init_visits)
            if iv.shape not in ((X, Y), (N, X, Y)) or (iv < 0).any() or (iv != np.floor(iv)).any() or iv.max() >= 2**32:
                raise ValueError(f"init_visits must hold non-negative integers of shape {(X, Y)} or {(N, X, Y)}")
            iv = np.array(np.broadcast_to(iv.astype(np.uint32), (N, X, Y)), order="C")  # a writable copy
            t = torch.from_numpy(iv.view(np.int32)).to(dev)
            keep.append(t)
            args.d_init_visits, args.init_visits_max = t.data_ptr(), int(iv.max()) if iv.size else 0
        if init_model_state is not None:
            ms = np.empty((N, 3))
            ms[:] = np.asarray(init_model_state, dtype=np.float64).reshape(-1, 3)
            args.d_init_model_state = dev_f64(ms).data_ptr()

        # visit maps: shared memory while the maze fits (about 50 x 50 tiles), else a device-memory scratch the kernel
        # indexes [tile][launched thread] (same step body, same results); NM_SIM_MAP=global forces the latter (tests)
        where = self._dmaze._lib.nm_simulate_fits(int(X), int(Y), int(n_steps) + int(args.init_visits_max))
        if where == 0:
            raise ValueError(f"a maze of {X} x {Y} tiles is too large for the simulation kernel (its tile and "
                             "grid-line tables must fit shared memory: about 66 000 tiles)")
        if where == 2 or os.environ.get("NM_SIM_MAP", "").startswith("g"):
            cols = int(self._dmaze._lib.nm_simulate_scratch_columns(int(N)))
            scratch = torch.empty((X * Y, cols), dtype=torch.int32, device=dev)
            keep.append(scratch)
            args.d_visit_scratch, args.visit_scratch_columns = scratch.data_ptr(), cols

        o_pose = out((N, 3), torch.float64)
        o_pcg = out((N, 4), torch.int64)
        o_p32 = out((N, 2), torch.int32, zero=True)
        o_it = out((N,), torch.int32)
        o_mov = out((N,), torch.int32)
        o_tc = out((N, 8), torch.int32)
        o_st = out((N,), torch.int32)
        o_ms = out((N, 5), torch.float64)
        args.d_final_pose, args.d_pcg_state_out = o_pose.data_ptr(), o_pcg.data_ptr()
        args.d_pcg_u32_out = o_p32.data_ptr()
        args.d_it_visit, args.d_count_moving = o_it.data_ptr(), o_mov.data_ptr()
        args.d_tile_counts, args.d_status, args.d_model_state = o_tc.data_ptr(), o_st.data_ptr(), o_ms.data_ptr()
        o_hp = o_ho = o_hc = o_occ = o_tot = None
        if history:
            o_hp, o_ho = out((T + 1, N, 2), torch.float64, zero=True), out((T + 1, N), torch.float64, zero=True)
            o_hc = out((T, N), torch.uint8, zero=True)
            args.d_hist_pos, args.d_hist_ori, args.d_hist_choice = o_hp.data_ptr(), o_ho.data_ptr(), o_hc.data_ptr()
        if occupancy:
            o_occ = out((N, nT), torch.int32)
            args.d_occupancy = o_occ.data_ptr()
        if occupancy_total:
            o_tot = out((nT,), torch.int64, zero=True)
            args.d_occupancy_total = o_tot.data_ptr()
        o_zone = o_static = o_sub = o_oh = edges = None
        if shock_zone is not None or corridors is not None:
            flags = torch.from_numpy(tile_flags(self.maze, shock_zone, corridors)).to(dev)
            keep.append(flags)
            args.d_tile_flags = flags.data_ptr()
            o_zone = out((N, 4), torch.int32)
            args.d_zone_metrics = o_zone.data_ptr()
        if static_intervals:
            o_static = out((N,), torch.int32)
            args.d_static_intervals = o_static.data_ptr()
        if subareas:
            if max(self.maze.subareas) >= _native.NM_MAX_SUBAREAS:
                raise ValueError(f"subarea indices must be below {_native.NM_MAX_SUBAREAS}")
            o_sub = out((N, _native.NM_MAX_SUBAREAS), torch.int32)
            args.d_subarea_counts = o_sub.data_ptr()
        if orientation_histogram is not None:
            from .analysis import metrics as M
            oh = dict(orientation_histogram)
            filt = oh.pop("tiles_filter", None)
            distance = bool(oh.pop("distance", False))
            bins, lo, hi = M.orientation_bins("hist_orientations", oh.pop("n_bins", None), oh.pop("max_lim", None),
                                              oh.pop("possible_actions", None))
            if oh:
                raise TypeError(f"unknown orientation_histogram options {sorted(oh)}")
            edges = np.ascontiguousarray(M.orientation_bin_edges(bins, lo, hi), dtype=np.float64)
            if len(edges) - 1 > _native.NM_MAX_ORI_BINS:
                raise ValueError(f"at most {_native.NM_MAX_ORI_BINS} orientation bins")
            if distance and len(edges) != 7:
                raise ValueError("hist_orientations(distance=True) is defined for six-bin histograms")
            args.n_ori_bins, args.ori_distance = len(edges) - 1, int(distance)
            args.h_ori_edges, args.ori_min_lim, args.ori_max_lim = edges.ctypes.data, float(lo), float(hi)
            if filt is not None:
                args.ori_tile_filter = sum(1 << int(t) for t in filt)
                if args.ori_tile_filter == 0:
                    raise ValueError("tiles_filter must name at least one tile type")
            o_oh = out((N, len(edges)), torch.int32)
            args.d_ori_hist = o_oh.data_ptr()

        with torch.cuda.device(dev):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _native.check(self._dmaze._lib.nm_simulate(self._dmaze.handle, C.byref(args), _native.cuda_stream_ptr()))
            e1.record()
            torch.cuda.current_stream().synchronize()
            self.last_kernel_ms = e0.elapsed_time(e1)  # device time of nm_sim_kernel (CUDA events)

        # per-mouse results: ONE pinned staging buffer (kept across runs), asynchronous copies, one synchronisation
        small = [t for t in (o_pose, o_st, o_it, o_mov, o_tc, o_pcg, o_ms, o_p32, o_zone, o_static, o_sub, o_oh, o_tot)
                 if t is not None]
        host = self._fetch(torch, small)

        def on_host(t):
            return host[id(t)] if id(t) in host else t.cpu().numpy()

        res = SimResult(final_pose=on_host(o_pose), status=on_host(o_st), it_visit=on_host(o_it),
                        count_moving=on_host(o_mov), tile_counts=on_host(o_tc),
                        pcg_state=on_host(o_pcg).view(np.uint64), model_state=on_host(o_ms),
                        pcg_buffered=on_host(o_p32).view(np.uint32))
        if history:
            res.history_position = np.ascontiguousarray(o_hp.cpu().numpy().transpose(1, 0, 2))
            res.history_orientation = np.ascontiguousarray(o_ho.cpu().numpy().T)
            res.choices = np.ascontiguousarray(o_hc.cpu().numpy().T)
        if occupancy:
            res.occupancy = o_occ.cpu().numpy().reshape(N, X, Y)
        if occupancy_total:
            res.occupancy_total = on_host(o_tot).reshape(X, Y)
        res.n_positions = T + 1
        if o_zone is not None:
            z = on_host(o_zone)
            if shock_zone is not None:
                res.zone_latency, res.zone_occupancy, res.zone_entries = z[:, 0], z[:, 1], z[:, 2]  # column views
            if corridors is not None:
                res.corridor_switches = z[:, 3]
        if o_static is not None:
            res.static_intervals = on_host(o_static)
        if o_sub is not None:
            res.subarea_counts = on_host(o_sub)
        if o_oh is not None:
            h = on_host(o_oh)
            res.orientations_histogram = h if args.ori_distance else h[:, :-1]
            res.orientations_histogram_bins = edges
        return res

    @functools.wraps(_run)
    def run(self, *args, **kwargs):
        with self._run_lock:
            return self._run(*args, **kwargs)

    # -- one large population through two host threads / CUDA streams ----------------------------------------------------
    def _worker(self):
        """A shallow copy for one pipeline thread: same maze / model / device tables, its own staging buffers and lock."""
        import copy
        w = copy.copy(self)
        w._run_lock = threading.RLock()
        w._arenas, w._pinned, w.last_kernel_ms = [], None, None
        w._workers = None
        w._fetch_views = True
        return w

    def run_pipelined(self, params, n_steps, init_pos, init_ori=0.0, init_disc=None, seeds=None, draws=None,
                      generators=None, device_seed=None, n_mice=None, chunks=6, workers=3, **kwargs):
        """:meth:`run` for a large population cut into ``chunks`` contiguous blocks of whole CTAs (the first one split
        1 : 3, so that the device starts early) that travel through ``workers`` host threads, each with its own CUDA
        stream and pinned staging buffers: the parameter preparation and upload of one block and the read-back of
        another overlap the kernel of a third (one :meth:`run` does them one after the other: for a million mice x
        1 000 steps a third of its wall time is host work; measured 158 -> 116 ms with the defaults).  Same results as
        :meth:`run`: the mice are independent, and device-seeded streams follow the global mouse index
        (``mouse_index_base``).

        Per-mouse arguments are given for the whole population, as for :meth:`run_sharded`.  ``history=True`` is not
        supported here (the histories of a large population do not belong on the host)."""
        import torch
        from concurrent.futures import ThreadPoolExecutor
        if kwargs.get("history"):
            raise ValueError("run_pipelined does not return histories; use run()")
        base = int(kwargs.pop("mouse_index_base", 0))
        if device_seed is not None:
            N = int(n_mice)
        elif draws is not None:
            N = len(draws)
        elif generators is not None:
            N = len(generators)
        else:
            seeds = list(seeds)
            N = len(seeds)
        bounds = pipeline_bounds(N, chunks)
        if len(bounds) <= 1:
            return self.run(params, n_steps, init_pos, init_ori, init_disc, seeds=seeds, draws=draws,
                            generators=generators, device_seed=device_seed, n_mice=n_mice, mouse_index_base=base, **kwargs)
        _native.require_device()
        if self._dmaze is None:
            self._dmaze = DeviceMaze(self.maze, self._device)
        dev = self._dmaze.torch_device
        X, Y = self.maze.shape
        n_workers = max(1, min(int(workers), 4))
        with self._run_lock:
            if getattr(self, "_workers", None) is None:
                self._workers = []
            while len(self._workers) < n_workers:
                self._workers.append((self._worker(), torch.cuda.Stream(device=dev)))
            workers = self._workers[:n_workers]
            for w, _ in workers:  # whatever may have changed on the parent since the copies were made
                w.v_table, w._dmaze = self.v_table, self._dmaze
            merged, lock = {}, threading.Lock()
            kernel_ms = [0.0]

            def cut(a, lo, hi, ndim=1):
                if a is None:
                    return None
                b = np.asarray(a)
                return b[lo:hi] if b.ndim == ndim and b.shape[0] == N else a

            def place(name, arr, lo, hi):
                """Rows [lo, hi) of the population-wide array `name` (allocated by whichever block arrives first)."""
                with lock:
                    if name not in merged:
                        merged[name] = np.empty((N,) + arr.shape[1:], dtype=arr.dtype)
                    dst = merged[name]
                rows = max(1, (4 << 20) // max(1, arr[0].nbytes if len(arr) else 1))
                if len(arr) <= 2 * rows:
                    np.copyto(dst[lo:hi], arr)
                else:  # large blocks: slices over a few threads (the page faults of the fresh array included)
                    list(_pool().map(lambda r0: np.copyto(dst[lo + r0:min(lo + r0 + rows, hi)], arr[r0:r0 + rows]),
                                     range(0, len(arr), rows)))

            def one(idx):
                lo, hi = bounds[idx]
                w, stream = workers[idx % n_workers]
                kw = dict(kwargs)
                for key, ndim in (("init_visits", 3), ("init_model_state", 2)):
                    if kw.get(key) is not None:
                        kw[key] = cut(kw[key], lo, hi, ndim)
                with w._run_lock, torch.cuda.device(dev), torch.cuda.stream(stream):
                    whole = w.v_table
                    if whole is not None and whole.shape == (N, X, Y):
                        w.v_table = whole[lo:hi]
                    try:
                        res = w._run({k: cut(v, lo, hi) for k, v in params.items()}, n_steps, cut(init_pos, lo, hi, 2),
                                     cut(init_ori, lo, hi), cut(init_disc, lo, hi, 2),
                                     seeds=None if seeds is None else seeds[lo:hi],
                                     draws=None if draws is None else np.asarray(draws)[lo:hi],
                                     generators=None if generators is None else list(generators)[lo:hi],
                                     device_seed=device_seed, n_mice=(hi - lo) if device_seed is not None else None,
                                     mouse_index_base=base + lo if device_seed is not None else 0, **kw)
                    finally:
                        w.v_table = whole
                shared = {}
                for name, val in vars(res).items():
                    if isinstance(val, np.ndarray) and val.ndim >= 1 and val.shape[0] == hi - lo and \
                            name not in ("occupancy_total", "orientations_histogram_bins"):
                        place(name, val, lo, hi)
                    else:  # population-wide values: copied out of the worker's staging buffer
                        shared[name] = val.copy() if isinstance(val, np.ndarray) else val
                with lock:
                    kernel_ms[0] += w.last_kernel_ms or 0.0
                return shared

            with ThreadPoolExecutor(max_workers=n_workers) as pool:
                # block i goes to worker i % n_workers: each worker handles its blocks in order, on its own stream
                futures = [pool.submit(lambda ids: [one(i) for i in ids], list(range(k, len(bounds), n_workers)))
                           for k in range(n_workers)]
                shared = [sh for f in futures for sh in f.result()]
        out = dict(shared[0])
        if shared[0].get("occupancy_total") is not None:
            out["occupancy_total"] = np.sum([sh["occupancy_total"] for sh in shared], axis=0)
        out.update(merged)
        self.last_kernel_ms = kernel_ms[0]  # sum of the blocks' kernel times (blocks of the two streams overlap)
        return SimResult(**out)

    # -- a population sharded over the GPUs of a box -------------------------------------------------
    def run_sharded(self, params, n_steps, init_pos, init_ori=0.0, init_disc=None, seeds=None, draws=None,
                    generators=None, device_seed=None, n_mice=None, group=None, **kwargs):
        """:meth:`run` for a population split over the ranks of an initialised ``torch.distributed`` process group (one
        process per GPU): the flat mouse index is cut into contiguous blocks (:func:`shard_bounds_mice`), every rank
        simulates its block -- no data-path collective; with ``device_seed`` the streams follow the GLOBAL mouse index,
        so the population does not depend on the number of ranks -- and the population totals are summed over the
        ranks by ONE small all-reduce (NCCL; gloo for CPU tensors): :func:`population_totals`.

        Per-mouse arguments (``params`` columns, ``init_*``, ``seeds`` / ``draws`` / ``generators``, a per-mouse
        ``v_table``, ``init_visits`` / ``init_model_state``) are given for the WHOLE population and sliced here.
        Without a process group this is :meth:`run` plus the totals.

        :return: ``(result of the local block, (lo, hi), totals)`` -- ``totals`` as :func:`population_totals` documents
        """
        import torch
        import torch.distributed as dist
        if device_seed is not None:
            N = int(n_mice)
        elif draws is not None:
            N = len(draws)
        elif generators is not None:
            N = len(generators)
        else:
            seeds = list(seeds)
            N = len(seeds)
        rank, world = 0, 1
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        lo, hi = shard_bounds_mice(N, rank, world)

        def cut(a, ndim=1):
            """Rows [lo, hi) of a per-mouse array (``ndim`` axes, the first of length N); anything else -- scalars,
            values shared by the population -- passes through."""
            if a is None:
                return None
            b = np.asarray(a)
            return b[lo:hi] if b.ndim == ndim and b.shape[0] == N else a

        X, Y = self.maze.shape
        local = {k: cut(v) for k, v in params.items()}
        for key, ndim in (("init_visits", 3), ("init_model_state", 2)):
            if kwargs.get(key) is not None:
                kwargs[key] = cut(kwargs[key], ndim)
        with self._run_lock:  # the per-mouse value tables are swapped for this rank's rows while the block runs
            whole_table = self.v_table
            if whole_table is not None and whole_table.shape == (N, X, Y):
                self.v_table = whole_table[lo:hi]
            try:
                if hi > lo:
                    res = self.run(local, n_steps, cut(init_pos, 2), cut(init_ori), cut(init_disc, 2),
                                   seeds=None if seeds is None else seeds[lo:hi],
                                   draws=None if draws is None else np.asarray(draws)[lo:hi],
                                   generators=None if generators is None else list(generators)[lo:hi],
                                   device_seed=device_seed, n_mice=(hi - lo) if device_seed is not None else None,
                                   mouse_index_base=lo if device_seed is not None else 0, **kwargs)
                else:
                    res = None  # more ranks than mice: this rank only takes part in the all-reduce
            finally:
                self.v_table = whole_table
        vec = pack_population_totals(res, X * Y)
        if world > 1:
            on_gpu = dist.get_backend(group) == "nccl"
            t = torch.from_numpy(vec)
            t = t.to(torch.device("cuda", torch.cuda.current_device())) if on_gpu else t
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            vec = t.cpu().numpy()
        n_bins = None if res is None or res.orientations_histogram is None else res.orientations_histogram.shape[1]
        return res, (lo, hi), unpack_population_totals(vec, (X, Y), n_bins)


def pipeline_bounds(N, chunks, align=128):
    """Blocks ``[(lo, hi), ...]`` of :meth:`BatchedExperiment.run_pipelined`: ``chunks`` contiguous blocks of whole CTAs
    (:func:`shard_bounds_mice`), the first one split 1 : 3 -- the device starts working after a quarter of a block's
    preparation, and from then on the host threads prepare blocks while another block runs."""
    chunks = max(1, min(int(chunks), (N + align - 1) // align))
    bounds = [shard_bounds_mice(N, c, chunks, align) for c in range(chunks)]
    bounds = [(lo, hi) for lo, hi in bounds if hi > lo]
    if len(bounds) > 1 and bounds[0][1] - bounds[0][0] >= 4 * align:
        lo, hi = bounds[0]
        mid = lo + (((hi - lo) // 4 + align - 1) // align) * align
        bounds[0:1] = [(lo, mid), (mid, hi)]
    return bounds


def shard_bounds_mice(N, rank, world, align=128):
    """Contiguous block ``[lo, hi)`` of the flat mouse index owned by ``rank``: blocks of ``align`` = 128 mice (one CTA
    of the simulation kernel) dealt out evenly, the last block ragged."""
    from .sweep import shard_bounds
    return shard_bounds(N, rank, world, align=align)


def pack_population_totals(res, n_tiles):
    """Sums over the mice of one shard as ONE int64 vector (what the all-reduce carries): mice, mice without error,
    positions, ``count_moving``, zone occupancy / entries, corridor switches, not-moving runs | ``tile_counts[8]`` |
    ``subarea_counts[16]`` | orientation histogram ``[NM_MAX_ORI_BINS + 1]`` | ``occupancy_total[n_tiles]`` (zeros for
    what the run did not compute; fixed layout, so that ranks without mice pack the same vector)."""
    head = np.zeros(8, dtype=np.int64)
    tiles, subs = np.zeros(8, dtype=np.int64), np.zeros(_native.NM_MAX_SUBAREAS, dtype=np.int64)
    hist, occ = np.zeros(_native.NM_MAX_ORI_BINS + 1, dtype=np.int64), np.zeros(n_tiles, dtype=np.int64)
    if res is not None:
        def total(a):
            return 0 if a is None else int(np.sum(a, dtype=np.int64))
        head[:] = [len(res.status), int((res.status == 0).sum()), len(res.status) * int(res.n_positions),
                   total(res.count_moving), total(res.zone_occupancy), total(res.zone_entries),
                   total(res.corridor_switches), total(res.static_intervals)]
        tiles[:] = np.sum(res.tile_counts, axis=0, dtype=np.int64)
        if res.subarea_counts is not None:
            subs[:] = np.sum(res.subarea_counts, axis=0, dtype=np.int64)
        if res.orientations_histogram is not None:
            hist[:res.orientations_histogram.shape[1]] = np.sum(res.orientations_histogram, axis=0, dtype=np.int64)
        if res.occupancy_total is not None:
            occ[:] = np.asarray(res.occupancy_total, dtype=np.int64).reshape(-1)
        elif res.occupancy is not None:
            occ[:] = np.sum(res.occupancy, axis=0, dtype=np.int64).reshape(-1)
    return np.concatenate([head, tiles, subs, hist, occ])


def unpack_population_totals(vec, shape, n_bins=None):
    """``dict`` view of :func:`pack_population_totals` (after the all-reduce: totals of the whole population);
    ``n_bins`` trims the orientation histogram to the width the run used."""
    vec = np.asarray(vec, dtype=np.int64)
    names = ("n_mice", "mice_ok", "n_positions", "count_moving", "zone_occupancy", "zone_entries", "corridor_switches",
             "static_intervals")
    out = {k: int(v) for k, v in zip(names, vec[:8])}
    o = 8
    out["tile_counts"] = vec[o:o + 8].copy()
    o += 8
    out["subarea_counts"] = vec[o:o + _native.NM_MAX_SUBAREAS].copy()
    o += _native.NM_MAX_SUBAREAS
    width = _native.NM_MAX_ORI_BINS + 1
    out["orientations_histogram"] = vec[o:o + (width if n_bins is None else n_bins)].copy()
    o += width
    out["occupancy_total"] = vec[o:].reshape(shape).copy()
    return out


def tile_flags(maze, shock_zone=None, corridors=None):
    """``uint8[X*Y]`` of ``NM_TILE_*`` bits for the fused shock-zone / corridor metrics (x-major).

    ``shock_zone``: ``{"subareas": int or sequence}`` -- inside = subarea in the set, leaving = any other tile -- or
    ``{"tile_lines": n}`` -- inside = ``y <= n - 1 and x <= 2``, leaving = ``y > n - 1 or x >= 6``
    (``analysis/metrics.py:488-586``).  ``corridors``: the two subarea indices ``arms`` counts passages between."""
    X, Y = maze.shape
    flags = np.zeros((X, Y), dtype=np.uint8)
    xs, ys = np.meshgrid(np.arange(X), np.arange(Y), indexing="ij")
    sub = None
    if (shock_zone is not None and "subareas" in shock_zone) or corridors is not None:
        sub = np.array([[maze.get_subarea(x, y) for y in range(Y)] for x in range(X)])
    if shock_zone is not None:
        if set(shock_zone) == {"subareas"}:
            z = shock_zone["subareas"]
            inside = np.isin(sub, [z] if isinstance(z, (int, np.integer)) else list(z))
            leaving = ~inside
        elif set(shock_zone) == {"tile_lines"}:
            low = ys <= int(shock_zone["tile_lines"]) - 1
            inside, leaving = low & (xs <= 2), ~low | (xs >= 6)
        else:
            raise ValueError("shock_zone must be {'subareas': ...} or {'tile_lines': n}")
        flags |= np.where(inside, _native.TILE_ZONE_IN, 0).astype(np.uint8)
        flags |= np.where(leaving, _native.TILE_ZONE_OUT, 0).astype(np.uint8)
    if corridors is not None:
        a, b = corridors
        flags |= np.where(sub == a, _native.TILE_CORRIDOR_A, 0).astype(np.uint8)
        flags |= np.where(sub == b, _native.TILE_CORRIDOR_B, 0).astype(np.uint8)
    return np.ascontiguousarray(flags.reshape(-1))


def movements_possible(maze, segments, device=None):
    """Vectorised ``Maze.is_movement_possible_cont`` on the GPU: ``segments[n, 4] = (x1, y1, x2, y2)``."""
    import torch
    _native.require_device()
    dm = DeviceMaze(maze, device)
    seg = torch.from_numpy(np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)).to(dm.torch_device)
    res = torch.empty(seg.shape[0], dtype=torch.uint8, device=dm.torch_device)
    with torch.cuda.device(dm.torch_device):
        _native.check(dm._lib.nm_maze_movements_possible_device(dm.handle, _native.tensor_ptr(seg), seg.shape[0],
                                                                _native.tensor_ptr(res), _native.cuda_stream_ptr()))
    return res.cpu().numpy().astype(bool)

