"""Discretised maze: dense tile tables + the continuous<->discrete geometry.

Host-side mirror of the reference's ``navigation_model.simulation.maze`` (``maze.py:7-627``; the
plotting half ``:631-736`` is out of scope).  The public surface (class / method names, argument
overloads ``f(x, y)`` vs ``f((x, y))``, return types, ``RuntimeError`` when subareas are missing) is the
reference's; the storage is not: the layout is a dense ``uint8[X, Y]`` array (``x`` = text *row*,
``y`` = column, ``maze.py:126-143``) because that is exactly what is uploaded to HBM / shared memory
by :meth:`Maze.flat_tables` (C-ABI ``nm_maze_upload``), and the list variants are vectorised.

Numerical contract:

* ``cont2disc`` is Python float floor-division ``int(x // size_tile)`` (``maze.py:393-394``) -- the
  mathematical floor of the *exact* quotient, which differs from ``floor(x / size_tile)`` on tile
  boundaries.  The device code (``csrc/nm_geometry.cuh``) restates CPython's ``float_divmod``.
* the ray / AABB walk of ``is_movement_possible_cont`` keeps the reference's operation order
  (swap so that ``x1 < x2`` before the slope, reuse of swapped values in the horizontal pass, both
  tiles adjacent to each crossed grid line tested), ``maze.py:563-611``.
"""
from enum import IntEnum

import numpy as np


class TileSystem:
    """Interface of a tile alphabet (``maze.py:7-23``)."""

    @staticmethod
    def char_to_mazetile(ch):
        raise NotImplementedError()

    @staticmethod
    def mazetile_tochar(mt):
        raise NotImplementedError()

    @staticmethod
    def is_visitable(t):
        raise NotImplementedError()


def _make_tile_system(name, letters):
    """Build a tile system from ``{char: (NAME, code)}``; VOID (code 0, char 'V') is implicit."""
    members = {"VOID": 0}
    for _, (nm, code) in letters.items():
        members[nm] = code
    tile_type = IntEnum("TileType", members)
    by_char = {ch: tile_type[nm] for ch, (nm, _) in letters.items()}
    by_tile = {tile_type[nm]: ch for ch, (nm, _) in letters.items()}
    by_tile[tile_type.VOID] = "V"

    class _System(TileSystem):
        TileType = tile_type

        @staticmethod
        def char_to_mazetile(ch):
            return by_char.get(ch, tile_type.VOID)

        @staticmethod
        def mazetile_tochar(mt):
            return by_tile.get(mt)

        @staticmethod
        def is_visitable(t):
            return t != tile_type.VOID

    _System.__name__ = _System.__qualname__ = name
    tile_type.__qualname__ = name + ".TileType"
    return _System


# codes follow maze.py:28-32 and :66-71
StandardTiles = _make_tile_system("StandardTiles", {"W": ("WALL", 1), "C": ("CORNER", 2), "M": ("CENTRE", 3)})
GridTiles = _make_tile_system("GridTiles", {"W": ("WALL", 1), "C": ("CORNER", 2), "M": ("CENTRE", 3),
                                            "D": ("DECISION_POINT", 4)})


def _rows(text):
    """Split a layout string into rows the way maze.py:133-143 does (a trailing newline adds no row,
    an empty line adds an empty row)."""
    rows = text.split("\n")
    if rows and rows[-1] == "":
        rows.pop()
    return rows


class Maze:
    """A maze of square tiles of side ``size_tile`` (reference ``maze.py:107-627``)."""

    def __init__(self, layout, subareas, size_tile, tile_system=StandardTiles):
        self._tile_system = tile_system
        self._size_tile = size_tile
        rows = _rows(layout)
        self._sizeX = len(rows)
        self._sizeY = len(rows[0]) if rows else 0
        if any(len(r) != self._sizeY for r in rows):
            raise RuntimeError("Maze: all layout rows must have the same length")
        self._tiles = np.array([[int(tile_system.char_to_mazetile(c)) for c in r] for r in rows],
                               dtype=np.uint8).reshape(self._sizeX, self._sizeY)
        self._visitable = self._tiles != int(self.MazeTile.VOID)
        self._sub = None
        if subareas is not None and subareas != "":
            srows = _rows(subareas)
            if len(srows) > 0:
                self._sub = [[int(c) for c in r] for r in srows]
        # centres of the visitable tiles in x-major order: nearest-visitable lookup table
        # (the reference builds a scipy KDTree over the same list, maze.py:117-122)
        self._vis_xy = np.argwhere(self._visitable)
        st = self._size_tile
        self._vis_centres = st * (self._vis_xy + 1) - st / 2

    @classmethod
    def from_file(cls, filename, size_tile, tile_system=StandardTiles):
        """Load ``layout [SUB subareas]`` from a text file (``maze.py:169-192``)."""
        with open(filename) as fh:
            text = fh.read()
        if len(text) == 0:
            return None
        parts = text.split("SUB")
        sub = ""
        if len(parts) > 1:
            sub = parts[1][1:] if parts[1][:1] == "\n" else parts[1]
        return cls(parts[0], sub, size_tile=size_tile, tile_system=tile_system)

    # ------------------------------------------------------------------ properties
    # noinspection PyPep8Naming
    @property
    def MazeTile(self):
        return self._tile_system.TileType

    @property
    def size_tile(self):
        return self._size_tile

    @property
    def shape(self):
        return self._sizeX, self._sizeY

    @property
    def size_x(self):
        return self._sizeX

    @property
    def size_y(self):
        return self._sizeY

    def description(self):
        """``{TileType: count}`` over the whole layout (``maze.py:217-227``)."""
        counts = np.bincount(self._tiles.ravel(), minlength=max(int(t) for t in self.MazeTile) + 1)
        return {t: int(counts[int(t)]) for t in self.MazeTile}

    @property
    def subareas(self):
        self.assert_subareas()
        return {self._sub[x][y] for x in range(self.size_x) for y in range(self.size_y)}

    def assert_subareas(self):
        if self._sub is None:
            raise RuntimeError("Requesting access to subareas, but no subareas present")

    # ------------------------------------------------------------------ dense tables (device upload)
    def flat_tables(self):
        """Dense, x-major (``index = x * size_y + y``) tables for the device.

        :return: dict with ``tiles`` u8[X*Y], ``visitable`` u8[X*Y], ``subareas`` u8[X*Y] or None,
                 ``state_of_tile`` i32[X*Y] (compact index of visitable tiles in x-major order, -1 for
                 VOID) and ``tile_of_state`` i32[NV]
        """
        vis = self._visitable.ravel()
        state_of_tile = np.full(vis.size, -1, dtype=np.int32)
        tile_of_state = np.flatnonzero(vis).astype(np.int32)
        state_of_tile[tile_of_state] = np.arange(tile_of_state.size, dtype=np.int32)
        sub = None
        if self._sub is not None:
            sub = np.zeros((self._sizeX, self._sizeY), dtype=np.uint8)
            for x, row in enumerate(self._sub[:self._sizeX]):
                sub[x, :min(len(row), self._sizeY)] = row[:self._sizeY]
            sub = sub.ravel()
        return {"tiles": self._tiles.ravel().copy(), "visitable": vis.astype(np.uint8),
                "subareas": sub, "state_of_tile": state_of_tile, "tile_of_state": tile_of_state}

    @property
    def tile_array(self):
        """``uint8[X, Y]`` tile codes (read-only view)."""
        v = self._tiles.view()
        v.flags.writeable = False
        return v

    @property
    def visitable_mask(self):
        """``bool[X, Y]`` visitable mask (read-only view)."""
        v = self._visitable.view()
        v.flags.writeable = False
        return v

    # ------------------------------------------------------------------ getters
    @staticmethod
    def _xy(x, y):
        if y is None:
            return x[0], x[1]
        return x, y

    def get_tile_centre(self, x, y=None):
        x, y = self._xy(x, y)
        st = self.size_tile
        return st * (x + 1) - st / 2, st * (y + 1) - st / 2

    def get_tile(self, x, y=None):
        x, y = self._xy(x, y)
        return self.MazeTile(int(self._tiles[x, y]))  # negative indices wrap like the list-of-lists

    def get_tiles(self, lpos):
        return [self.get_tile(p) for p in lpos]

    def get_subarea(self, x, y=None):
        self.assert_subareas()
        x, y = self._xy(x, y)
        return self._sub[x][y]

    def get_subareas(self, lpos):
        return [self.get_subarea(p) for p in lpos]

    def get_coordinates_of_type(self, tile_type):
        return [(int(x), int(y)) for x, y in np.argwhere(self._tiles == int(tile_type))]

    def get_visitable_coordinates(self):
        return [(int(x), int(y)) for x, y in self._vis_xy]

    def get_non_visitable_coordinates(self):
        return [(int(x), int(y)) for x, y in np.argwhere(~self._visitable)]

    def get_number_visitable_tiles(self):
        return int(self._visitable.sum())

    def get_closest_visitable(self, x, y=None):
        x, y = self._xy(x, y)
        if self.is_visitable(x, y):
            return x, y
        return self.get_closest_visitable_cont(self.disc2cont(x, y))

    # ------------------------------------------------------------------ continuous coordinates
    def cont2disc(self, x, y=None):
        x, y = self._xy(x, y)
        st = self.size_tile
        return int(x // st), int(y // st)

    def cont2disc_list(self, lpos):
        return [self.cont2disc(p) for p in lpos]

    def cont2disc_array(self, pos):
        """Vectorised :meth:`cont2disc`: ``float64[N, 2] -> int64[N, 2]`` (numpy ``floor_divide`` on
        float64 is the same divmod as Python's ``//``)."""
        pos = np.asarray(pos, dtype=np.float64)
        return np.floor_divide(pos, self.size_tile).astype(np.int64)

    def disc2cont(self, x, y=None):
        return self.get_tile_centre(x, y)

    def disc2cont_list(self, lpos):
        return [self.disc2cont(p) for p in lpos]

    def get_tile_cont(self, x, y=None):
        return self.get_tile(self.cont2disc(x, y))

    def get_tiles_cont(self, lpos):
        return [self.get_tile_cont(p) for p in lpos]

    def get_subarea_cont(self, x, y=None):
        return self.get_subarea(self.cont2disc(x, y))

    def get_subareas_cont(self, lpos):
        return [self.get_subarea_cont(p) for p in lpos]

    def get_closest_visitable_cont(self, x, y=None):
        """Tile containing the point if visitable, else the visitable tile with the nearest centre
        (``maze.py:469-484``; exact ties resolve to the first tile in x-major order)."""
        x, y = self._xy(x, y)
        if self.is_visitable_cont(x, y):
            return self.cont2disc(x, y)
        d = self._vis_centres - np.array([x, y], dtype=np.float64)
        i = int(np.argmin(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]))
        return self.cont2disc(self._vis_centres[i])

    def closest_visitable_array(self, pos):
        """Vectorised :meth:`get_closest_visitable_cont`: ``float64[N, 2] -> int64[N, 2]``."""
        pos = np.asarray(pos, dtype=np.float64).reshape(-1, 2)
        disc = self.cont2disc_array(pos)
        ok = self.visitable_array(disc)
        for i in np.flatnonzero(~ok):
            disc[i] = self.get_closest_visitable_cont(pos[i, 0], pos[i, 1])
        return disc

    # ------------------------------------------------------------------ checks
    def is_visitable(self, x, y=None):
        x, y = self._xy(x, y)
        if x < 0 or x >= self.size_x or y < 0 or y >= self.size_y:
            return False
        return bool(self._visitable[x, y])

    def are_visitable(self, lpos):
        return [self.is_visitable(p) for p in lpos]

    def visitable_array(self, disc):
        """Vectorised :meth:`is_visitable`: ``int[N, 2] -> bool[N]``."""
        disc = np.asarray(disc)
        inside = (disc[:, 0] >= 0) & (disc[:, 0] < self.size_x) & (disc[:, 1] >= 0) & (disc[:, 1] < self.size_y)
        out = np.zeros(len(disc), dtype=bool)
        out[inside] = self._visitable[disc[inside, 0], disc[inside, 1]]
        return out

    def is_movement_possible(self, x1, y1, x2=None, y2=None):
        if y2 is None:
            (x1, y1), (x2, y2) = (x1[0], x1[1]), (y1[0], y1[1])
        return self.is_movement_possible_cont(*self.disc2cont(x1, y1), *self.disc2cont(x2, y2))

    def are_movements_possible(self, x, y, lpos=None):
        if lpos is None:
            lpos = y
            x, y = x[0], x[1]
        return [self.is_movement_possible(x, y, p[0], p[1]) for p in lpos]

    def is_visitable_cont(self, x, y=None):
        return self.is_visitable(self.cont2disc(x, y))

    def are_visitable_cont(self, lpos):
        return [self.is_visitable_cont(p) for p in lpos]

    def _free_at(self, x, y):
        st = self.size_tile
        return self.is_visitable(int(x // st), int(y // st))

    def is_movement_possible_cont(self, x1, y1, x2=None, y2=None):
        """Segment test: both end tiles visitable and, at every crossed vertical / horizontal grid line,
        the tiles on both sides of the crossing point visitable (``maze.py:563-611``)."""
        if y2 is None:
            (x1, y1), (x2, y2) = (x1[0], x1[1]), (y1[0], y1[1])
        st = self.size_tile
        if not (self._free_at(x1, y1) and self._free_at(x2, y2)):
            return False
        if x1 != x2:  # vertical grid lines x = (i+1)*st between the two abscissae
            if x1 > x2:
                x1, x2, y1, y2 = x2, x1, y2, y1
            slope = (y2 - y1) / (x2 - x1)
            for i in range(int(x1 // st), int(x2 // st)):
                gx = (i + 1) * st
                gy = slope * (gx - x1) + y1
                if not (self._free_at(gx - st, gy) and self._free_at(gx, gy)):
                    return False
        if y1 != y2:  # horizontal grid lines, on the (possibly swapped) endpoints
            if y1 > y2:
                x1, x2, y1, y2 = x2, x1, y2, y1
            slope = (x2 - x1) / (y2 - y1)
            for j in range(int(y1 // st), int(y2 // st)):
                gy = (j + 1) * st
                gx = slope * (gy - y1) + x1
                if not (self._free_at(gx, gy - st) and self._free_at(gx, gy)):
                    return False
        return True

    def are_movements_possible_cont(self, x, y, lpos=None):
        if lpos is None:
            lpos = y
            x, y = x[0], x[1]
        return [self.is_movement_possible_cont(x, y, p[0], p[1]) for p in lpos]

    # ------------------------------------------------------------------ text output
    def print_layout(self):
        for row in self._tiles:
            print("".join(self._tile_system.mazetile_tochar(self.MazeTile(int(t))) + " " for t in row))

    def print_subareas(self):
        self.assert_subareas()
        for row in self._sub:
            print("".join(str(a) + " " for a in row))

    # ------------------------------------------------------------------ discrete transition table
    def transition_table(self, discrete_actions):
        """``next[S, A]`` compact-state transition table for discrete (tile-step) agents.

        ``discrete_actions`` are integer tile displacements ``(dx, dy)``; entry ``[s, a]`` is the compact
        state reached from visitable state ``s`` with action ``a`` if :meth:`is_movement_possible`
        (``maze.py:512-527``) allows it, else ``-1``.
        """
        tabs = self.flat_tables()
        acts = np.asarray(discrete_actions, dtype=np.int64).reshape(-1, 2)
        out = np.full((len(tabs["tile_of_state"]), len(acts)), -1, dtype=np.int16)
        for s, tile in enumerate(tabs["tile_of_state"]):
            x, y = divmod(int(tile), self.size_y)
            for a, (dx, dy) in enumerate(acts):
                nx, ny = x + int(dx), y + int(dy)
                if self.is_visitable(nx, ny) and self.is_movement_possible(x, y, nx, ny):
                    out[s, a] = tabs["state_of_tile"][nx * self.size_y + ny]
        return out
