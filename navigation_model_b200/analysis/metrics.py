"""Trajectory metrics on the hot path + the Metric / Analyzer pipeline they plug into.

Host-side mirror of the part of the reference's ``analysis/metrics.py`` that sits next to the
simulation hot path: the dependency-resolving ``Metric`` / ``@metric`` /
``Analyzer`` machinery (``:8-186``) and ``discretize_positions`` (``:192-201``), ``occupancy_map``
(``:204-216``), ``tile_analysis`` (``:219-248``), ``motion_analysis`` (``:425-442``),
``it_perc_visited_maze`` (``:697-731``), plus ``compute_orientations`` (``:653-694``) which ``Session``
needs.  For a single history these run on the host (vectorised numpy); for batches of simulated mice
the same quantities come fused out of the simulation kernel (``batched_sim.SimResult``).  The remaining
metric functions of the reference (orientation histograms, shock-zone metrics, ...) are out of scope
(SURVEY section 8f, rank 2).
"""
import inspect

import numpy as np


class Metric:
    """A callable with named requirements and named products (``:8-89``)."""

    def __init__(self, fun, requirements=None, products=None, defaults=None):
        self._fun = fun
        self._requirements = list(requirements) if requirements is not None else []
        self._product_labels = products if products is not None else []
        self._defaults = defaults if defaults is not None else {}
        self._products = None

    def _inputs(self, available):
        picked = {}
        for r in self._requirements:
            if r in self._defaults:
                picked[r] = self._defaults[r]
            if r in available:
                picked[r] = available[r]
        return picked

    def __call__(self, **kwargs):
        res = self._fun(**self._inputs(kwargs))
        parts = res if type(res) == tuple else (res,)
        self._products = dict(zip(self._product_labels, parts))
        return res

    def is_callable(self, available):
        return all(r in available or r in self._defaults for r in self._requirements)

    @property
    def requirements(self):
        return self._requirements

    @property
    def products(self):
        return self._products


def metric(*products):
    """Decorator: take requirements / defaults from the signature, products from the arguments (``:92-113``)."""

    def deco(fun):
        sig = inspect.signature(fun)
        fun.requirements = list(sig.parameters.keys())
        fun.products = products
        fun.defaults = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
        return fun

    return deco


class Analyzer:
    """Runs metrics as soon as their requirements become available (``:116-186``)."""

    def __init__(self, *metrics):
        self._metrics = []
        for m in metrics:
            if type(m) is Metric:
                self._metrics.append(m)
            else:
                self._metrics.append(Metric(m, m.requirements, m.products, getattr(m, "defaults", None)))
        self._last = None

    def analize(self, **current_measures):
        pending = list(range(len(self._metrics)))
        while pending:
            ready = next((i for i in pending if self._metrics[i].is_callable(current_measures.keys())), None)
            if ready is None:
                wanted = set()
                for m in self._metrics:
                    wanted.update(m.requirements)
                raise RuntimeError("Analysis failed: one among these measures cannot be produced: "
                                   f"{wanted - set(current_measures.keys())}")
            self._metrics[ready](**current_measures)
            current_measures.update(self._metrics[ready].products)
            pending.remove(ready)
        self._last = current_measures
        return current_measures

    def __call__(self, **current_measures):
        return self.analize(**current_measures)

    @property
    def last(self):
        return self._last


# ------------------------------------------------------------------------------------ metric functions
@metric("discrete_positions")
def discretize_positions(continuous_positions, maze):
    """Continuous positions -> list of tile coordinates (``:192-201``)."""
    return maze.cont2disc_list(continuous_positions)


def _as_cells(discrete_positions):
    return np.asarray(discrete_positions, dtype=np.int64).reshape(-1, 2)


@metric("occupancy")
def occupancy_map(discrete_positions, maze):
    """Number of visits of every tile (``:204-216``)."""
    occ = np.zeros(maze.shape)
    cells = _as_cells(discrete_positions)
    if len(cells):
        np.add.at(occ, (cells[:, 0], cells[:, 1]), 1)
    return occ


@metric("tiles_perc", "tiles_ratio", "tiles_count")
def tile_analysis(discrete_positions, maze, tiles=None):
    """Percentage / ratio / count of visits by tile type (``:219-248``)."""
    if tiles is None:
        tiles = list(maze.MazeTile)
    n_positions = len(discrete_positions)
    count = {t: 0 for t in tiles}
    for cell in discrete_positions:
        count[maze.get_tile(cell)] += 1  # KeyError for a type that was not asked for, like the reference
    perc = {t: count[t] * 100 / n_positions for t in tiles}
    totals = maze.description()
    ratio = {t: count[t] / totals[t] if totals[t] > 0 else 0 for t in tiles}
    return perc, ratio, count


@metric("count_moving", "moving_bool_list")
def motion_analysis(discrete_positions):
    """Number of tile changes and the per-step moved flags (``:425-442``)."""
    n = len(discrete_positions)
    flags = [False] * n
    for i in range(1, n):
        if discrete_positions[i] != discrete_positions[i - 1]:
            flags[i] = True
    return sum(flags), flags


@metric("it_2_visit_perc_maze")
def it_perc_visited_maze(discrete_positions, maze, percentage=80):
    """First index at which ``percentage`` % of the visitable tiles have been seen (``len`` if never;
    ``:697-731``)."""
    n_visitable = maze.get_number_visitable_tiles()
    seen = set()
    for idx, cell in enumerate(discrete_positions):
        key = tuple(cell)
        if key not in seen:
            seen.add(key)
            perc = (len(seen) * 100) / n_visitable
            if perc >= percentage:
                return idx
            if perc > 100:
                raise RuntimeError(f"perc_visit_maze: {perc} > 100 !!!")
    return len(discrete_positions)


def get_endpoints_without_mouse(cont_pos, cont_ori, possible_actions, maze):
    """Tiles under ``R(cont_ori) * a + cont_pos`` for every action (``:638-650``)."""
    c, s = np.cos(cont_ori), np.sin(cont_ori)
    rot = np.array([[c, -s], [s, c]])
    ends = np.zeros((len(possible_actions), 2))
    for i, act in enumerate(possible_actions):
        ends[i] = cont_pos + np.matmul(rot, act)
    return maze.cont2disc_list(ends)


@metric("absolute_orientations")
def compute_orientations(continuous_positions, maze=None, possible_actions=None):
    """Heading of every step from consecutive positions; steps that do not move (or, with a maze and an
    action set, land on a tile no action reaches) repeat the previous heading (``:653-694``)."""
    out = []
    prev_ori = 0
    theta = None
    n = len(continuous_positions)
    for i in range(n - 1):
        p1, p2 = continuous_positions[i], continuous_positions[i + 1]
        moved = not np.allclose(p1, p2)
        if moved and not (maze is None and possible_actions is None):
            moved = maze.cont2disc(p2) in get_endpoints_without_mouse(p1, prev_ori, possible_actions, maze)
        if moved:
            theta = np.arctan2(p2[1] - p1[1], p2[0] - p1[0])
            if theta > np.pi:
                theta = np.pi * 2 - theta
            elif theta < -np.pi:
                theta = -np.pi * 2 - theta
            out.append(theta)
        elif i > 0 and len(out) != 0:
            out.append(out[-1])
        if theta is not None:
            prev_ori = theta
    return [out[0]] * (n - len(out)) + out
