"""Seeded synthetic workloads of SURVEY section 8 / BASELINE.json (the reference ships no mouse data: every benchmark
session is synthetic): the 9 x 9 test maze of the reference (``test/data/maze_layout.txt``,
``test/test_metrics.py:115``), the 8-action set "A8", the all-visitable 20 x 20 maze of configs[3] and the
random-walk session recipe.  Inputs only -- used by ``bench.py`` and ``tools/``; the test oracle carries its own
restatement of the same recipe and ``tests/test_host_cpu.py`` holds the two to identical bits."""
import numpy as np

from .simulation.maze import Maze

#: 8 relative actions ("A8"), in units of size_tile
A8_UNITS = np.array([(0, 0), (1, 0), (2, 0), (1, 1), (1, -1), (0, 1), (0, -1), (-1, 0)], dtype=np.float64)
#: the eight tile displacements of the tile-action agents (model-based, Q-table)
MOVES8 = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]

M9_LAYOUT = "\n".join(["CWWWWWWWC", "WMMMMMMMW", "CWWWWWWWW", "VVVVVVVWW", "VVVVVVVWW", "VVVVVVVWW",
                       "CWWWWWWWW", "WMMMMMMMW", "CWWWWWWWC"])
M9_SUBAREAS = "\n".join(["111222333", "111222333", "111222333", "000000044", "000000044", "000000044",
                         "777666555", "777666555", "777666555"])
M9_SIZE_TILE = 0.111
M20_SIZE_TILE = 0.05


def maze_m9():
    """The reference's 9 x 9 test maze (81 tiles, 60 visitable)."""
    return Maze(M9_LAYOUT, M9_SUBAREAS, size_tile=M9_SIZE_TILE)


def maze_m20():
    """All-visitable 20 x 20 maze: the 400-state maze of BASELINE configs[3]."""
    return Maze("\n".join(["M" * 20] * 20), "", size_tile=M20_SIZE_TILE)


def a8(size_tile=M9_SIZE_TILE):
    return A8_UNITS * size_tile


def synthetic_session(maze, session_id, T=5000, start=(1, 1), goal=(7, 4), jitter=0.04):
    """The seeded benchmark session of SURVEY section 8c: 9-neighbourhood (stay included, ``dx`` slowest) random walk
    over visitable tiles from ``default_rng(session_id)`` -- per step the next tile by ``rng.integers`` first, then the
    jitter ``rng.uniform(-j, j, 2)`` of the current point; reward 1.0 when the jittered point lies on ``goal``.
    Returns ``(trajectory[T, 2], reward[T])``."""
    rng = np.random.default_rng(session_id)
    tile = tuple(start)
    traj = np.zeros((T, 2))
    rew = np.zeros(T)
    steps = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1)]
    X, Y = maze.shape
    for t in range(T):
        nbrs = [(tile[0] + dx, tile[1] + dy) for dx, dy in steps
                if 0 <= tile[0] + dx < X and 0 <= tile[1] + dy < Y and maze.is_visitable(tile[0] + dx, tile[1] + dy)]
        nxt = nbrs[rng.integers(len(nbrs))]
        cx, cy = maze.get_tile_centre(tile[0], tile[1])
        jx, jy = rng.uniform(-jitter, jitter, 2)
        traj[t] = (cx + jx, cy + jy)
        rew[t] = 1.0 if tuple(maze.cont2disc(traj[t, 0], traj[t, 1])) == tuple(goal) else 0.0
        tile = nxt
    return traj, rew
