"""Trajectory metrics on the hot path + the Metric / Analyzer pipeline they plug into.

Host-side mirror of the part of the reference's ``analysis/metrics.py`` that sits next to the
simulation hot path: the dependency-resolving ``Metric`` / ``@metric`` /
``Analyzer`` machinery (``:8-186``) and ``discretize_positions`` (``:192-201``), ``occupancy_map``
(``:204-216``), ``tile_analysis`` (``:219-248``), ``motion_analysis`` (``:425-442``),
``it_perc_visited_maze`` (``:697-731``), ``compute_orientations`` (``:653-694``) which ``Session`` needs,
and the remaining trajectory metrics real objectives compare: ``histogram_tiles``
(``:251-277``), ``hist_orientations`` / ``relative_orientation_signatures`` / ``hist_orientations_filtered``
(``:280-422``), ``hist_time_moving`` (``:445-467``), ``subareas`` (``:470-485``), the shock-zone metrics
(``:488-586``), ``mov_bouts`` / ``compute_static_intervals`` (``:589-635``) and ``arms`` (``:735-769``).
For a single history these run on the host (numpy); for batches of simulated mice the same quantities come
fused out of the simulation kernel (``batched_sim.SimResult``).
"""
import inspect
import math

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


# ------------------------------------------------------------------------------ SURVEY 8f rank 2 metrics
@metric("tiles_histogram")
def histogram_tiles(discrete_positions, maze, tiles=None):
    """Visits per tile of each type, as the vector (CORNER, WALL, CENTRE) (``:251-277``)."""
    _, ratio, _ = tile_analysis(discrete_positions, maze, tiles)
    T = maze.MazeTile
    return np.array([ratio[T.CORNER], ratio[T.WALL], ratio[T.CENTRE]])


def orientation_bins(caller, n_bins=None, max_lim=None, possible_actions=None):
    """``(bins, min_lim, max_lim)`` of the relative-orientation histograms (``:280-303``): either ``n_bins`` (default
    8) even bins ending at ``max_lim`` (default 1.125 pi), or uneven bins whose edges are the midpoints between
    the directions of ``possible_actions`` (wrapping around)."""
    if n_bins is not None and possible_actions is not None:
        raise RuntimeError(f"Cannot use both n_bins and possible_actions in {caller}")
    if possible_actions is not None:
        acts = np.array(possible_actions)
        ang = np.unique(np.sort(np.arctan2(acts[:, 1], acts[:, 0])))
        bins = [(ang[-1] - 2 * np.pi + ang[0]) / 2]
        bins += [(ang[i] + ang[i + 1]) / 2 for i in range(len(ang) - 1)]
        bins.append((ang[0] + 2 * np.pi + ang[-1]) / 2)
        return bins, bins[0], bins[-1]
    if max_lim is None:
        max_lim = np.pi * 1.125
    return (8 if n_bins is None else n_bins), max_lim - 2 * np.pi, max_lim


_hist_orientation_commons = orientation_bins  # the reference's (private) name


def orientation_bin_edges(bins, min_lim, max_lim):
    """The edges ``np.histogram(x, bins=bins, range=(min_lim, max_lim))`` uses (what the device histogram searches)."""
    return np.histogram([], bins=bins, range=(min_lim, max_lim))[1]


def _wrapped_turns(absolute_orientations, min_lim, max_lim):
    """``ori[t+1] - ori[t]`` brought into ``[min_lim, max_lim]`` by one turn (``:329-334``)."""
    ori = np.asarray(absolute_orientations, dtype=np.float64)
    rel = ori[1:] - ori[:-1]
    return np.where(rel > max_lim, rel - np.pi * 2, np.where(rel < min_lim, rel - (-np.pi * 2), rel))


def _tile_changed(discrete_positions):
    n = len(discrete_positions)
    return np.array([discrete_positions[t + 1] != discrete_positions[t] for t in range(n - 1)], dtype=bool)


@metric("orientations_histogram", "relative_orientations", "orientations_histogram_bins")
def hist_orientations(absolute_orientations, discrete_positions, n_bins=None, max_lim=None, possible_actions=None,
                      distance=False):
    """Histogram of the turns ``ori[t+1] - ori[t]`` over the steps that change tile (``:306-359``).

    ``distance=True`` (six-bin histograms): straight-ahead turns (bin 2) that jump more than 1.5 tiles are taken out
    of the histogram and counted in one extra trailing entry."""
    if len(absolute_orientations) != len(discrete_positions):
        raise RuntimeError(f"Wrong input sizes for hist_orientations {len(absolute_orientations)} != "
                           f"{len(discrete_positions)}")
    bins, min_lim, max_lim = orientation_bins("hist_orientations", n_bins, max_lim, possible_actions)
    if len(absolute_orientations) < 2:
        rel, moved = np.zeros(0), np.zeros(0, dtype=bool)
    else:
        rel, moved = _wrapped_turns(absolute_orientations, min_lim, max_lim), _tile_changed(discrete_positions)
    keep = moved.copy()
    two_steps = 0
    if distance:
        edges = orientation_bin_edges(bins, min_lim, max_lim)
        if len(edges) != 7:
            raise ValueError("hist_orientations(distance=True) is defined for six-bin histograms")
        for t in np.flatnonzero(moved):
            if np.histogram(rel[t], bins=bins, range=(min_lim, max_lim))[0][2] == 1 and \
                    math.dist(discrete_positions[t], discrete_positions[t + 1]) > 1.5:
                keep[t] = False
                two_steps += 1
    relative = [float(r) for r in rel[keep]]
    hist, edges = np.histogram(relative, bins=bins, range=(min_lim, max_lim))
    if distance:
        return np.concatenate((hist, np.array([two_steps])), axis=0), relative, edges
    return hist, relative, edges


@metric("rel_ori_signature", "rel_ori_signature_histogram")
def relative_orientation_signatures(orientations_histogram):
    """Three contrasts of the eight-bin turn histogram -- forward vs lateral, forward vs backward, backward vs
    the 2-step forward bin -- as a tuple and as an array (``:362-380``)."""
    h = orientations_histogram
    lateral = np.median([h[0], h[1], h[3], h[4]])
    sig = (h[2] - lateral, h[2] - h[5], h[6] - h[5])
    return sig, np.array(sig)


@metric("orientations_histogram", "relative_orientations", "orientations_histogram_bins")
def hist_orientations_filtered(absolute_orientations, discrete_positions, maze, orientation_tiles_filter, n_bins=None,
                               max_lim=None, possible_actions=None):
    """:func:`hist_orientations` restricted to the steps that START on a tile whose type is in
    ``orientation_tiles_filter`` (``:383-422``)."""
    if len(absolute_orientations) != len(discrete_positions):
        raise RuntimeError(f"Wrong input sizes for hist_orientations_filtered {len(absolute_orientations)} != "
                           f"{len(discrete_positions)}")
    bins, min_lim, max_lim = orientation_bins("hist_orientations_filtered", n_bins, max_lim, possible_actions)
    n = len(absolute_orientations)
    if n < 2:
        relative = []
    else:
        rel, moved = _wrapped_turns(absolute_orientations, min_lim, max_lim), _tile_changed(discrete_positions)
        on_type = np.array([maze.get_tile(discrete_positions[t]) in orientation_tiles_filter for t in range(n - 1)],
                           dtype=bool)
        relative = [float(r) for r in rel[moved & on_type]]
    hist, edges = np.histogram(relative, bins=bins, range=(min_lim, max_lim))
    return hist, relative, edges


@metric("histogram_time_moving")
def hist_time_moving(moving_bool_list):
    """Run-length encoding of the moved flags: ``[(n_steps, 1 moving / 0 not moving), ...]`` (``:445-467``)."""
    runs = []
    current, length = moving_bool_list[0], 1
    for flag in moving_bool_list[1:]:
        if flag == current:
            length += 1
        else:
            runs.append((length, int(current)))
            current, length = (not current), 1
    runs.append((length, int(current)))
    return runs


@metric("subareas_histogram")
def subareas(discrete_positions, maze):
    """Visits per subarea, one bin per index between the smallest and largest subarea of the maze (``:470-485``)."""
    ids = maze.subareas
    lo, hi = min(ids), max(ids)
    return np.histogram(maze.get_subareas(discrete_positions), bins=hi - lo + 1, range=[lo, hi])[0]


def _shock_zone_membership(discrete_positions, maze, subareas_shock_zone, tile_lines):
    """Per position: inside the shock zone / in the region that ends a visit.  ``None`` when neither or both of the
    zone descriptions are given (the reference then falls through)."""
    if subareas_shock_zone is not None and tile_lines is None:
        zones = [subareas_shock_zone] if type(subareas_shock_zone) == int else subareas_shock_zone
        inside = np.array([s in zones for s in maze.get_subareas(discrete_positions)], dtype=bool)
        return inside, ~inside
    if subareas_shock_zone is None and tile_lines is not None:
        cells = _as_cells(discrete_positions)
        low = cells[:, 1] <= tile_lines - 1
        return low & (cells[:, 0] <= 2), ~low | (cells[:, 0] >= 6)
    return None


@metric("latency_enter_shock_zone")
def latency_enter_shock_zone(discrete_positions, maze, subareas_shock_zone=None, tile_lines=None):
    """Index of the first position inside the shock zone, ``len - 1`` if it is never entered (``:488-516``).  The zone
    is a set of subareas, or the tiles with ``y <= tile_lines - 1`` and ``x <= 2``."""
    member = _shock_zone_membership(discrete_positions, maze, subareas_shock_zone, tile_lines)
    if member is not None and member[0].any():
        return int(np.argmax(member[0]))
    return len(discrete_positions) - 1


@metric("occupancy_shock_zone")
def occupancy_shock_zone(discrete_positions, tile_lines=5):
    """Number of positions with ``y <= tile_lines - 1`` and ``x <= 2`` (``:519-534``)."""
    cells = _as_cells(discrete_positions)
    return int(np.count_nonzero((cells[:, 1] <= tile_lines - 1) & (cells[:, 0] <= 2)))


@metric("shock_zone_entries")
def shock_zone_entries(discrete_positions, maze, subareas_shock_zone=None, tile_lines=None):
    """Number of visits to the shock zone (``:537-586``): a visit starts on a position inside the zone and ends when
    the mouse leaves it (subarea description) or reaches ``y > tile_lines - 1`` / ``x >= 6`` (tile-line description)."""
    member = _shock_zone_membership(discrete_positions, maze, subareas_shock_zone, tile_lines)
    if member is None:
        raise UnboundLocalError("shock_zone_entries needs exactly one of subareas_shock_zone and tile_lines")
    entries, inside = 0, False
    for enters, leaves in zip(*member):
        if enters and not inside:
            entries, inside = entries + 1, True
        elif leaves and inside:
            inside = False
    return entries


@metric("bouts_mov", "bouts_stop", "moving_dynamics_histogram")
def mov_bouts(histogram_time_moving):
    """Total steps spent in moving and in not-moving bouts (``:589-616``)."""
    mov = sum(n for n, kind in histogram_time_moving if kind == 1)
    stop = sum(n for n, kind in histogram_time_moving if kind != 1)
    return mov, stop, np.array([mov, stop])


@metric("static_intervals")
def compute_static_intervals(histogram_time_moving):
    """Number of not-moving bouts (``:621-635``)."""
    return sum(1 for _, kind in histogram_time_moving if kind == 0)


@metric("corridors_switches")
def arms(discrete_positions, maze):
    """Number of passages between the two corridors, subareas 1 and 7 (``:735-769``)."""
    side, switches = None, 0
    for cell in discrete_positions:
        sub = maze.get_subarea(cell[0], cell[1])
        if sub in (1, 7):
            if side is not None and side != sub:
                switches += 1
            side = sub
    return switches
