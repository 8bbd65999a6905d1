"""Behavioural components: the free-parameter namespace of the model and the per-candidate values the
softmax policy sums.

Host-side mirror of the reference's ``simulation/behavioural_components.py`` (``:10-544``): the same
parameter system (``ParameterDescriptor`` / ``Parameter`` / ``@parameter`` / ``@weight`` with range
validation raising ``RuntimeError``, ``:10-215``), the same ``Component`` interface (``update(**kw)``,
``values``, ``weighted_values``, ``parameters``, ``name``, ``get_descriptors()``, ``:220-313``) and the six
components (``:315-544``).  The scalar ``update`` methods are the per-step plugin interface (what the
reference's tests drive); the batched simulation does not call them -- it flattens the components into
per-mouse parameter arrays (:meth:`Component.device_kind`, :func:`bcost_tables`, :func:`normalise_v_table`)
and evaluates them inside the CUDA kernel (``csrc/nm_sim.cu``).
"""
from dataclasses import dataclass
from enum import IntEnum

import numpy as np

from .. import _native
from .maze import GridTiles, StandardTiles


# ------------------------------------------------------------------------------ parameter system
@dataclass(frozen=True)
class ParameterDescriptor:
    """Name, bounds and default of a parameter -- no value, so it can live on the class (``:10-69``)."""
    name: str
    minv: float
    maxv: float
    default: float

    def __str__(self):
        return f"[{self.name}; range: ({self.minv}; {self.maxv}); default: {self.default}]"


class Parameter:
    """A descriptor bound to a current, range-checked value; created per component instance (``:72-139``)."""

    __slots__ = ("_descriptor", "_val")

    def __init__(self, descriptor, **kwargs):
        self._descriptor = descriptor
        self.val = kwargs.get(descriptor.name, descriptor.default)

    def __getattr__(self, item):  # name / minv / maxv are those of the descriptor
        if item in ("name", "minv", "maxv"):
            return getattr(self._descriptor, item)
        raise AttributeError(item)

    @property
    def val(self):
        return self._val

    @val.setter
    def val(self, value):
        lo, hi = self._descriptor.minv, self._descriptor.maxv
        if not lo <= value <= hi:
            raise RuntimeError(
                f"Wrong value set for property '{self._descriptor.name}' ({value}), must be in range [{lo}; {hi}]")
        self._val = value

    def __str__(self):
        return f"[{self.name} = {self.val}; range: ({self.minv}; {self.maxv})]"


def parameter(name, doc, minv=-np.inf, maxv=np.inf, default=0.0, isweight=False):
    """Class decorator: declare a free parameter of a component (``:142-202``)."""

    def deco(cls):
        if not issubclass(cls, Component):
            raise RuntimeError("Trying to apply a parameter decorator to a class not inheriting from Component")
        if isweight and hasattr(cls, "_weight_idx"):
            raise RuntimeError("A component cannot have more than one weight")
        descs = cls.__dict__.get("_descriptors")
        if descs is None:
            descs = list(getattr(cls, "_descriptors", []))
            cls._descriptors = descs
            header = "\n\n    Parameters for this component:"
        else:
            header = ""
        if name in [d.name for d in descs]:
            raise RuntimeError(f"Component cannot have the same parameter {name} twice")
        descs.append(ParameterDescriptor(name, minv, maxv, default))
        cls.__doc__ = (cls.__doc__ or "") + header + "\n        " + name + ": " + doc
        idx = len(descs) - 1

        def getter(self):
            return self._parameters[idx].val

        def setter(self, value):
            self._parameters[idx].val = value

        setattr(cls, name, property(getter, setter, doc=doc))
        if isweight:
            cls._weight_idx = idx
        return cls

    return deco


def weight(name, doc="weight of the component", minv=0.0, maxv=10.0, default=1.0):
    """Class decorator: declare the (single) weight of a component (``:205-215``)."""
    return parameter(name, doc=doc, minv=minv, maxv=maxv, default=default, isweight=True)


# noinspection PyUnresolvedReferences,PyProtectedMember
class Component:
    """Something that assigns a value to every candidate next position (``:220-313``)."""

    #: NM_COMP_* code when the simulation kernel implements the component, else None (host plugin)
    device_kind = None

    def __init__(self, debug=False, active=True, **kwargs):
        self._val = 0
        self.active = active
        self.debug = debug
        descs = type(self)._descriptors
        if "_parameters" in kwargs:
            self._parameters = kwargs["_parameters"]
            if len(self._parameters) != len(descs):
                raise RuntimeError(f"Wrong number of parameters ({len(self._parameters)}). Required: {len(descs)}.")
        else:
            self._parameters = []
            for d in descs:
                try:
                    self._parameters.append(Parameter(d, **kwargs))
                except KeyError:
                    raise RuntimeError(f"Value for {d.name} has not been specified")

    @property
    def weight(self):
        return self._parameters[type(self)._weight_idx].val

    @property
    def values(self):
        return self._val if self.active else 0

    @property
    def weighted_values(self):
        return self._val * self.weight if self.active else 0

    @property
    def parameters(self):
        return {p.name: p.val for p in self._parameters}

    def update(self, **kwargs):
        raise NotImplementedError

    @property
    def name(self):
        return type(self)._name

    @classmethod
    def get_descriptors(cls):
        return cls._descriptors


# -------------------------------------------------------------------------------------- components
def _anxiety_table(c, w, m, d=None, v=-1.0):
    """Value by tile type (``:335-341, 366-371``)."""
    tab = {StandardTiles.TileType.WALL: w, StandardTiles.TileType.CORNER: c, StandardTiles.TileType.CENTRE: m,
           StandardTiles.TileType.VOID: v}
    if d is not None:
        tab[GridTiles.TileType.DECISION_POINT] = d
    return tab


@weight("W_anx")
@parameter("p1", doc="value of a CORNER tile")
@parameter("p2", doc="WALL value as a fraction of the CORNER value", minv=0.0, maxv=1.0)
@parameter("p3", doc="CENTRE value as a fraction of the WALL value", minv=0.0, maxv=1.0)
@parameter("p4", doc="value of a DECISION_POINT tile")
class AnxietyComponent_Grid(Component):
    """Values the type of the tile under each candidate (grid mazes with decision points, ``:315-346``)."""

    _name = "anx"
    device_kind = _native.COMP_ANXIETY

    def __init__(self, **params):
        super().__init__(**params)
        self._maze_weights = _anxiety_table(c=self.p1, w=self.p2 * self.p1, m=self.p3 * self.p2 * self.p1, d=self.p4)

    def update(self, next_tiles, **kwargs):
        self._val = np.array([self._maze_weights[t] for t in next_tiles])


@weight("W_anx")
@parameter("p1", doc="value of a CORNER tile")
@parameter("p2", doc="WALL value as a fraction of the CORNER value", minv=0.0, maxv=1.0)
@parameter("p3", doc="CENTRE value as a fraction of the WALL value", minv=0.0, maxv=1.0)
class AnxietyComponent(Component):
    """Values the type of the tile under each candidate (``:348-376``)."""

    _name = "anx"
    device_kind = _native.COMP_ANXIETY

    def __init__(self, **params):
        super().__init__(**params)
        self._maze_weights = _anxiety_table(c=self.p1, w=self.p2 * self.p1, m=self.p3 * self.p2 * self.p1)

    def update(self, next_tiles, **kwargs):
        self._val = np.array([self._maze_weights[t] for t in next_tiles])


def bcost_tables(discrete_actions, k, gamma, out=None):
    """Cached per-action values of :class:`BiomechanicalCostComponent` for arrays of ``k`` / ``gamma``
    (``:393-406``): von Mises pdf of the action's direction (0 for the null action), times ``gamma`` for
    actions whose Chebyshev length is >= 2.  Returns ``float64[N, A]`` (written into ``out`` when given, e.g. pinned
    memory on its way to the device)."""
    acts = np.array(discrete_actions)
    k = np.atleast_1d(np.asarray(k, dtype=np.float64))
    gamma = np.atleast_1d(np.asarray(gamma, dtype=np.float64))
    angles = np.arctan2(acts[:, 1], acts[:, 0])
    null = np.all(acts == (0, 0), axis=1)
    far = np.linalg.norm(acts, ord=np.inf, axis=1) >= 2
    if not np.all(np.isfinite(k) & (k >= 0)):
        from scipy.stats import vonmises
        res = vonmises.pdf(angles[None, :], k[:, None])
        res[:, null] = 0.0
        res[:, far] *= gamma[:, None]  # the other columns are multiplied by 1.0 in the reference: unchanged bits
        if out is None:
            return res
        out[...] = res
        return out
    # scipy's own formula (vonmises_gen._pdf: exp(kappa * cosm1(x)) / (2 pi i0e(kappa))) without the generic
    # rv_continuous.pdf machinery: same bits (tests/test_host_api.py), ~100x faster for a million mice.  Elementwise,
    # so large batches are cut into row blocks and spread over a few threads (numpy releases the GIL in its loops).
    import scipy.special as sc
    cm1 = sc.cosm1(angles)[None, :]
    if out is None:
        out = np.empty((len(k), len(acts)))
    assert out.shape == (len(k), len(acts)) and out.dtype == np.float64

    def block(lo, hi):
        kk = k[lo:hi]
        np.multiply(kk[:, None], cm1, out=out[lo:hi])
        np.exp(out[lo:hi], out=out[lo:hi])
        out[lo:hi] /= (2 * np.pi * sc.i0e(kk))[:, None]
        out[lo:hi, null] = 0.0
        out[lo:hi, far] *= gamma[lo:hi, None]

    rows = 1 << 15
    if len(k) <= 2 * rows:
        block(0, len(k))
    else:
        from .._hostpool import pool
        spans = [(lo, min(lo + rows, len(k))) for lo in range(0, len(k), rows)]
        list(pool().map(lambda sp: block(*sp), spans))
    return out


@weight("W_bcost")
@parameter("k", doc="concentration of the von Mises heading preference")
@parameter("gamma", doc="factor applied to actions two or more tiles long")
class BiomechanicalCostComponent(Component):
    """Values the direction of the movement; contributes only on the step after a stay (``:379-416``)."""

    _name = "bcost"
    device_kind = _native.COMP_BCOST

    def __init__(self, discrete_actions, **params):
        super().__init__(**params)
        self._discrete_actions = np.array(discrete_actions)
        self._cachedval = bcost_tables(self._discrete_actions, self.k, self.gamma)[0]

    def update(self, moving, is_visitable, **kwargs):
        self._val = self._cachedval[is_visitable] if moving else np.zeros(np.sum(is_visitable))


@weight("W_exp")
@parameter("xi", doc="familiarity averaging rate", default=0.006, minv=0.006, maxv=0.6)
@parameter("kappa_home", doc="visit-count decay in the HOME state", default=0.01, minv=0.01, maxv=5.0)
@parameter("kappa_explo", doc="visit-count decay in the EXPLO state", default=0.01, minv=0.01, maxv=5.0)
@parameter("theta_explo", doc="familiarity above which HOME switches to EXPLO", default=80.0, minv=80.0, maxv=120.0)
@parameter("theta_home", doc="familiarity below which EXPLO switches to HOME", default=30.0, minv=30.0, maxv=70.0)
class ExperienceComponent(Component):
    """Visit counts, a familiarity average and a HOME / EXPLO hysteresis (``:419-482``)."""

    _name = "exp"
    device_kind = _native.COMP_EXPERIENCE

    class State(IntEnum):
        EXPLO = 0
        HOME = 1

    def __init__(self, maze_shape, maze_map_exp=None, **params):
        super().__init__(**params)
        self._ment_state = self.State.EXPLO
        self._maze_map_exp = np.array(maze_map_exp) if maze_map_exp is not None else np.zeros(maze_shape)
        self._familiarity = 0.0

    def update(self, x, y, next_positions, **kwargs):
        visits = self._maze_map_exp
        visits[x, y] += 1
        self._familiarity = (1 - self.xi) * self._familiarity + self.xi * visits[x, y]
        if self._ment_state == self.State.HOME and self._familiarity >= self.theta_explo:
            self._ment_state = self.State.EXPLO
        elif self._ment_state == self.State.EXPLO and self._familiarity < self.theta_home:
            self._ment_state = self.State.HOME
        self._val = np.zeros(len(next_positions))
        exploring = self._ment_state == self.State.EXPLO
        for i, pos in enumerate(next_positions):
            n = visits[pos[0], pos[1]]
            self._val[i] = np.exp(-self.kappa_explo * n) if exploring else 1.0 - np.exp(-self.kappa_home * n)


@weight("W_bper")
@parameter("bp", doc="base persistence value")
@parameter("Wm", doc="factor of the null action after a stay")
@parameter("Wnm", doc="factor of the moving actions after a move")
class BiomechPersistenceComponent(Component):
    """Tendency to keep moving / keep still (``:485-516``)."""

    _name = "bper"
    device_kind = _native.COMP_BPERSISTENCE

    def __init__(self, discrete_actions, **params):
        super().__init__(**params)
        self._discrete_actions = np.array(discrete_actions)
        null = np.all(self._discrete_actions == (0, 0), axis=1)
        n = len(self._discrete_actions)
        self._cachedval_mov = np.full(n, self.bp, dtype=np.float64)
        self._cachedval_mov[null] = self.Wm * self.bp
        self._cachedval_not_mov = np.full(n, self.Wnm * self.bp, dtype=np.float64)
        self._cachedval_not_mov[null] = self.bp

    def update(self, moving, is_visitable, **kwargs):
        self._val = (self._cachedval_mov if moving else self._cachedval_not_mov)[is_visitable]


def normalise_v_table(v_table):
    """``(V - min) / (max - min) - 1`` (zeros for an all-zero table), ``:530-535``."""
    v_table = np.asarray(v_table, dtype=np.float64)
    hi, lo = np.max(v_table), np.min(v_table)
    if hi == lo == 0:
        return np.zeros(v_table.shape)
    return (v_table - lo) / (hi - lo) - 1


@weight("W_values")
class ValueTableComponent(Component):
    """Looks the candidates up in a value table learned beforehand, normalised to [-1, 0] (``:519-544``)."""

    _name = "v_table"
    device_kind = _native.COMP_VTABLE

    def __init__(self, v_table, **params):
        super().__init__(**params)
        self._v_table = normalise_v_table(v_table)

    def update(self, next_positions, **kwargs):
        self._val = np.zeros(len(next_positions))
        for i, pos in enumerate(next_positions):
            self._val[i] = self._v_table[pos[0], pos[1]]
