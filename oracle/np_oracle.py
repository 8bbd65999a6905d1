"""CPU oracle (numpy / plain Python) of the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker.  The product (``navigation_model_b200``) never
imports anything from ``oracle/``.

It restates the reference's algorithm for the path, function by function, citing the reference
``file:line`` each one follows (paths relative to ``/root/reference/src/navigation_model`` unless they
start with ``src/``).  It is deliberately independent of the product code: it has its own maze
geometry, mouse pose arithmetic and stream extraction.

Pinning:

* V-tables of ``TD0`` / ``PositiveConditioning`` / ``NegativeConditioning`` with and without
  ``ShuffledReplays``, the extracted transitions, softmax step probabilities / choices of
  ``BehaviouralModel.decision_making``, ``StandardExperiment.run`` trajectories under replayed draws
  and the occupancy metrics are PINNED: ``tests/golden/make_golden.py`` runs the *unmodified* reference
  (imported from ``/root/reference`` through ``oracle/refharness.py``) and commits its outputs as
  fixtures; ``tests/test_oracle_golden.py`` checks this oracle against them bit for bit.
* The **log-likelihood** is an extension with no reference implementation: **parity unpinned**.  It is
  composed from pinned primitives -- the value recurrences above and the softmax of
  ``simulation/exploration_model.py:84-93`` -- evaluated in log space.
"""
import math

import numpy as np

ALG_TD0, ALG_POSITIVE, ALG_NEGATIVE = 0, 1, 2
OBS_NONE = 255


# ------------------------------------------------------------------------------------------ maze
class OracleMaze:
    """Tile grid + geometry (simulation/maze.py:107-627), dense numpy layout.

    ``tiles[x, y]``: 0 = VOID, 1 = WALL, 2 = CORNER, 3 = CENTRE, 4 = DECISION_POINT
    (maze.py:28-32, 66-71); x = text row, y = column (maze.py:126-143).
    """

    CODES = {"W": 1, "C": 2, "M": 3, "D": 4}

    def __init__(self, layout, size_tile, grid_tiles=False):
        rows = layout.split("\n")
        if rows and rows[-1] == "":
            rows.pop()
        codes = dict(self.CODES)
        if not grid_tiles:
            codes.pop("D")  # StandardTiles maps unknown chars (incl. 'D') to VOID, maze.py:35-42
        self.tiles = np.array([[codes.get(c, 0) for c in r] for r in rows], dtype=np.uint8)
        self.X, self.Y = self.tiles.shape
        self.st = size_tile
        self.vis = self.tiles != 0
        self.vis_cells = [(int(x), int(y)) for x, y in np.argwhere(self.vis)]  # x-major
        self.state_of = {c: i for i, c in enumerate(self.vis_cells)}
        self.centres = np.array([self.centre(c) for c in self.vis_cells])

    @classmethod
    def from_file(cls, path, size_tile, grid_tiles=False):
        with open(path) as fh:
            text = fh.read()
        return cls(text.split("SUB")[0], size_tile, grid_tiles)  # maze.py:183-190

    def centre(self, c):  # maze.py:262-263
        return self.st * (c[0] + 1) - self.st / 2, self.st * (c[1] + 1) - self.st / 2

    def cont2disc(self, x, y):  # maze.py:393-394  (Python float floor-division)
        return int(x // self.st), int(y // self.st)

    def visitable(self, c):  # maze.py:499-501
        x, y = c
        if x < 0 or x >= self.X or y < 0 or y >= self.Y:
            return False
        return bool(self.vis[x, y])

    def closest_visitable(self, x, y):  # maze.py:469-484 (KD-tree nearest centre -> exhaustive search)
        c = self.cont2disc(x, y)
        if self.visitable(c):
            return c
        d = (self.centres[:, 0] - x) ** 2 + (self.centres[:, 1] - y) ** 2
        cx, cy = self.centres[int(np.argmin(d))]
        return self.cont2disc(cx, cy)

    def move_possible(self, x1, y1, x2, y2):  # maze.py:563-611
        st = self.st
        if not self.visitable(self.cont2disc(x1, y1)) or not self.visitable(self.cont2disc(x2, y2)):
            return False
        if x1 != x2:
            if x1 > x2:
                x1, x2 = x2, x1
                y1, y2 = y2, y1
            m = (y2 - y1) / (x2 - x1)
            for dx in range(int(x1 // st), int(x2 // st)):
                x = (dx + 1) * st
                y = m * (x - x1) + y1
                if not self.visitable(self.cont2disc(x - st, y)):
                    return False
                if not self.visitable(self.cont2disc(x, y)):
                    return False
        if y1 != y2:
            if y1 > y2:
                x1, x2 = x2, x1
                y1, y2 = y2, y1
            m = (x2 - x1) / (y2 - y1)
            for dy in range(int(y1 // st), int(y2 // st)):
                y = (dy + 1) * st
                x = m * (y - y1) + x1
                if not self.visitable(self.cont2disc(x, y - st)):
                    return False
                if not self.visitable(self.cont2disc(x, y)):
                    return False
        return True


# ----------------------------------------------------------------------------------------- mouse
class OracleMouse:
    """Pose arithmetic of the C++ Mouse (src/_navigation_model/mouse.cpp:64-87)."""

    def __init__(self, pos, ori, actions):
        self.x, self.y = float(pos[0]), float(pos[1])
        self.ori = float(ori)
        self.actions = np.asarray(actions, dtype=np.float64).reshape(-1, 2)

    def endpoints(self):  # mouse.cpp:64-73
        c, s = math.cos(self.ori), math.sin(self.ori)
        return [((c * ax + (-s) * ay) + self.x, (s * ax + c * ay) + self.y) for ax, ay in self.actions]

    def move_to(self, x, y):  # mouse.cpp:75-81; Eigen isApprox, precision 1e-12
        dx, dy = x - self.x, y - self.y
        if (dx * dx + dy * dy) <= (1e-12 * 1e-12) * min(x * x + y * y, self.x * self.x + self.y * self.y):
            return False
        self.ori = math.atan2(y - self.y, x - self.x)
        self.x, self.y = x, y
        return True


# -------------------------------------------------------------------------- stream extraction
def extract_transitions(maze, trajectory, reward, mouse=None):
    """The parameter-independent part of ``ConditioningExperiment.run`` (simulation/rl.py:85-102) and of
    ``PositiveConditioning.update`` (rl.py:177-182, 189).

    :return: dict of plain arrays: ``s, s1`` (compact state ids, x-major over visitable tiles), ``r``,
             ``ncand``, ``cand[T-1, 8]``, ``obs`` (first k with cand[k] == s1, else 255)
    """
    T = len(trajectory)
    if T != len(reward):
        raise RuntimeError("trajectory and reward must have the same length")  # rl.py:82-83
    n = T - 1
    out = {"s": np.zeros(n, np.int32), "s1": np.zeros(n, np.int32), "r": np.zeros(n, np.float64),
           "ncand": np.zeros(n, np.int32), "cand": np.zeros((n, 8), np.int32),
           "obs": np.full(n, OBS_NONE, np.int32)}
    s = maze.closest_visitable(trajectory[0][0], trajectory[0][1])  # rl.py:87
    for i in range(1, T):
        cs, cs1 = trajectory[i - 1], trajectory[i]
        s1 = maze.closest_visitable(cs1[0], cs1[1])  # rl.py:94
        out["s"][i - 1] = maze.state_of[s]
        out["s1"][i - 1] = maze.state_of[s1]
        out["r"][i - 1] = reward[i]  # rl.py:95
        if mouse is not None:
            mouse.move_to(cs[0], cs[1])  # rl.py:178
            cells = [maze.cont2disc(ex, ey) for ex, ey in mouse.endpoints()]  # rl.py:179-180
            nxt = [c for c in cells if maze.visitable(c)]  # rl.py:181-182
            if len(nxt) == 0:
                raise ValueError("zero-size array to reduction operation maximum which has no identity")
            if len(nxt) > 8:
                raise ValueError("oracle supports at most 8 candidates")
            out["ncand"][i - 1] = len(nxt)
            for k, c in enumerate(nxt):
                out["cand"][i - 1, k] = maze.state_of[c]
                if out["obs"][i - 1] == OBS_NONE and c == s1:
                    out["obs"][i - 1] = k
            mouse.move_to(cs1[0], cs1[1])  # rl.py:189
        s = s1
    return out


def replay_orders(n, n_times, seed):
    """Index orders of ``ShuffledReplays.offline_replays`` (rl.py:419-424): ``n_times`` cumulative in-place
    ``Generator.shuffle`` calls; shuffling ``arange(n)`` draws the same permutation as shuffling the
    buffer list."""
    rng = np.random.default_rng(seed)
    idx = np.arange(n)
    orders = []
    for _ in range(n_times):
        rng.shuffle(idx)
        orders.append(idx.copy())
    return orders


# ------------------------------------------------------------------------- value recurrences
def np_sum(values):
    """``numpy.sum`` of a contiguous float64 vector of length <= 8, spelled out: sequential from 0.0 for
    n < 8, the 8-accumulator pairwise tree for n == 8."""
    n = len(values)
    if n < 8:
        acc = 0.0
        for v in values:
            acc = acc + v
        return acc
    assert n == 8
    e = values
    return ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]))


def step_update(alg, V, s, s1, r, cand, alpha, gamma):
    """One value update on the compact table ``V`` (python floats, left-to-right like the reference)."""
    if alg == ALG_TD0:  # rl.py:257-259
        rpe = r + gamma * V[s1] - V[s]
    elif alg == ALG_POSITIVE:  # rl.py:184-186
        rpe = r + gamma * max(V[c] for c in cand) - V[s]
    else:  # rl.py:227-229
        rpe = r + gamma * min(V[c] for c in cand) - V[s]
    V[s] += alpha * rpe


def step_logp(V, cand, obs, beta):
    """Log-probability of the observed candidate under ``softmax(beta * V[cand])``
    (simulation/exploration_model.py:84-93, in log space).  EXTENSION -- parity unpinned."""
    val = [beta * V[c] for c in cand]  # :87
    m = max(val)  # :92
    val = [v - m for v in val]
    e = [math.exp(v) for v in val]  # :93
    return val[obs] - math.log(np_sum(e))


def run_session(alg, tr, alpha, gamma, n_states, beta=None, v0=None, orders=()):
    """Online pass (+ replay passes) of one parameter set over one session.

    :param tr: output of :func:`extract_transitions`
    :param beta: None -> value learning only; else accumulate the log-likelihood over online steps
    :param orders: replay index orders (see :func:`replay_orders`)
    :return: ``(V[n_states], ll)``
    """
    V = [0.0] * n_states if v0 is None else [float(v) for v in v0]
    ll = 0.0
    n = len(tr["s"])
    s_, s1_, r_, k_, c_, o_ = (tr["s"].tolist(), tr["s1"].tolist(), tr["r"].tolist(), tr["ncand"].tolist(),
                               tr["cand"].tolist(), tr["obs"].tolist())
    for i in range(n):
        cand = c_[i][:k_[i]]
        if beta is not None and o_[i] != OBS_NONE:
            ll = ll + step_logp(V, cand, o_[i], beta)
        step_update(alg, V, s_[i], s1_[i], r_[i], cand, alpha, gamma)
    for order in orders:
        for i in order.tolist():
            step_update(alg, V, s_[i], s1_[i], r_[i], c_[i][:k_[i]], alpha, gamma)
    return np.array(V), ll


def dense_table(maze, V, v0_dense=None):
    """Scatter a compact table to ``[X, Y]`` (untouched tiles keep their initial value / 0)."""
    out = np.zeros((maze.X, maze.Y)) if v0_dense is None else np.array(v0_dense, dtype=np.float64)
    for i, (x, y) in enumerate(maze.vis_cells):
        out[x, y] = V[i]
    return out


# --------------------------------------------------------------------------- synthetic sessions
def synthetic_session(maze, session_id, T=5000, start=(1, 1), goal=(7, 4), jitter=0.04):
    """The seeded benchmark session of SURVEY section 8c / BASELINE.md section 4.

    9-neighbourhood (stay included, ``dx`` slowest) random walk over visitable tiles: for every step draw
    the next tile with ``rng.integers`` first, then the jitter ``rng.uniform(-j, j, 2)`` of the current
    point; reward 1.0 when the jittered point lies on ``goal``.  Known answer for ``session_id=0``,
    T=5000, ``TD0(0.1, 0.9)`` on the 9x9 test maze: ``V.max() = 0.5949675648481678``,
    ``V[7,4] = 0.48326347528794783``, 108 rewarded steps, 60 distinct tiles.
    """
    rng = np.random.default_rng(session_id)
    tile = tuple(start)
    traj = np.zeros((T, 2))
    rew = np.zeros(T)
    steps = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1)]
    for t in range(T):
        nbrs = [(tile[0] + dx, tile[1] + dy) for dx, dy in steps if maze.visitable((tile[0] + dx, tile[1] + dy))]
        nxt = nbrs[rng.integers(len(nbrs))]
        cx, cy = maze.centre(tile)
        jx, jy = rng.uniform(-jitter, jitter, 2)
        traj[t] = (cx + jx, cy + jy)
        rew[t] = 1.0 if maze.cont2disc(traj[t, 0], traj[t, 1]) == tuple(goal) else 0.0
        tile = nxt
    return traj, rew


#: 8 relative actions of SURVEY section 8 ("A8"), in units of size_tile
A8_UNITS = np.array([(0, 0), (1, 0), (2, 0), (1, 1), (1, -1), (0, 1), (0, -1), (-1, 0)], dtype=np.float64)

#: the 9x9 test maze of the reference (test/data/maze_layout.txt), restated so that GPU-box runs need no
#: access to /root/reference
M9_LAYOUT = "\n".join(["CWWWWWWWC", "WMMMMMMMW", "CWWWWWWWW", "VVVVVVVWW", "VVVVVVVWW", "VVVVVVVWW",
                       "CWWWWWWWW", "WMMMMMMMW", "CWWWWWWWC"])
M9_SUBAREAS = "\n".join(["111222333", "111222333", "111222333", "000000044", "000000044", "000000044",
                         "777666555", "777666555", "777666555"])
M9_SIZE_TILE = 0.111


# ==================================================================================== simulation
# Restatement of StandardExperiment.run (simulation/experiment.py:23-51) with BehaviouralModel
# (simulation/exploration_model.py:68-109), the five behavioural components
# (simulation/behavioural_components.py:348-544) and the C++ Mouse.  Pinned by
# tests/golden/sim_golden.npz (histories, probabilities and metrics produced by the reference).
COMP_ANXIETY, COMP_BCOST, COMP_EXPERIENCE, COMP_BPERSISTENCE, COMP_VTABLE = 1, 2, 3, 4, 5


def bcost_table(discrete_actions, k, gamma):
    """BiomechanicalCostComponent.__init__ cached values (behavioural_components.py:393-406)."""
    from scipy.stats import vonmises
    acts = np.array(discrete_actions)
    angles = np.arctan2(acts[:, 1], acts[:, 0])
    cached = vonmises(k).pdf(angles)
    cached[np.all(acts == (0, 0), axis=1)] = 0.0
    dist = np.linalg.norm(acts, ord=np.inf, axis=1)
    gammas = np.ones(len(acts)) * gamma
    gammas[dist < 2] = 1
    return cached * gammas


def normalise_vtable(v_table):
    """ValueTableComponent.__init__ (behavioural_components.py:530-535)."""
    v_table = np.asarray(v_table, dtype=np.float64)
    max_v, min_v = np.max(v_table), np.min(v_table)
    if max_v == min_v == 0:
        return np.zeros(v_table.shape)
    return (v_table - min_v) / (max_v - min_v) - 1


def simulate(maze, actions, discrete_actions, init_pos, init_ori, init_disc, beta, components, draws,
             record_probs=False):
    """One mouse, ``len(draws)`` steps, with replayed uniform draws.

    :param components: ordered list of ``(kind, params)``: ANXIETY ``dict(W_anx,p1,p2,p3[,p4])``, BCOST
        ``dict(W_bcost, table)`` (see :func:`bcost_table`), EXPERIENCE ``dict(W_exp, xi, kappa_home,
        kappa_explo, theta_explo, theta_home)``, BPERSISTENCE ``dict(W_bper, bp, Wm, Wnm)``, VTABLE
        ``dict(W_values, table)`` (normalised, see :func:`normalise_vtable`)
    :return: dict(hist_pos [T+1,2], hist_ori [T+1], choice [T] action indices, probs, status)
    """
    mouse = OracleMouse(init_pos, init_ori, actions)
    zero_act = np.all(np.array(discrete_actions) == (0, 0), axis=1)
    T = len(draws)
    hist_pos = [(mouse.x, mouse.y)]
    hist_ori = [mouse.ori]
    choice = []
    probs = []
    prev = np.array(init_disc, dtype=np.float64)  # exploration_model.py:47
    moving = False  # :48
    exp_map = np.zeros((maze.X, maze.Y))  # behavioural_components.py:445
    fam, home = 0.0, False  # :440, 447
    status = 0
    for t in range(T):
        ends = mouse.endpoints()  # experiment.py:36
        vis = [maze.move_possible(mouse.x, mouse.y, ex, ey) for ex, ey in ends]  # :37
        idxs = [i for i, v in enumerate(vis) if v]
        if not idxs:
            status = t + 1  # np.max([]) raises in the reference
            break
        cand = [ends[i] for i in idxs]  # :38
        cells = [maze.cont2disc(ex, ey) for ex, ey in cand]  # :39
        x, y = maze.cont2disc(mouse.x, mouse.y)  # :41
        val_fin = np.zeros(len(cand))  # exploration_model.py:84
        for kind, p in components:
            if kind == COMP_ANXIETY:  # behavioural_components.py:364-374
                w = {2: p["p1"], 1: p["p2"] * p["p1"], 3: p["p3"] * p["p2"] * p["p1"], 0: -1.0, 4: p.get("p4", 0.0)}
                v = np.array([w[int(maze.tiles[c])] for c in cells])
                val_fin = val_fin + v * p["W_anx"]
            elif kind == COMP_BCOST:  # :408-413
                v = p["table"][vis] if moving else np.zeros(np.sum(vis))
                val_fin = val_fin + v * p["W_bcost"]
            elif kind == COMP_EXPERIENCE:  # :449-479
                exp_map[x, y] += 1
                fam = (1 - p["xi"]) * fam + p["xi"] * exp_map[x, y]
                if home and fam >= p["theta_explo"]:
                    home = False
                elif (not home) and fam < p["theta_home"]:
                    home = True
                v = np.zeros(len(cells))
                for i, c in enumerate(cells):
                    if not home:
                        v[i] = np.exp(-p["kappa_explo"] * exp_map[c[0], c[1]])
                    else:
                        v[i] = 1.0 - np.exp(-p["kappa_home"] * exp_map[c[0], c[1]])
                val_fin = val_fin + v * p["W_exp"]
            elif kind == COMP_BPERSISTENCE:  # :502-513
                mov = np.array([p["bp"]] * len(zero_act))
                mov[zero_act] = p["Wm"] * p["bp"]
                nmov = np.array([p["Wnm"] * p["bp"]] * len(zero_act))
                nmov[zero_act] = p["bp"]
                v = mov[vis] if moving else nmov[vis]
                val_fin = val_fin + v * p["W_bper"]
            elif kind == COMP_VTABLE:  # :537-541
                v = np.array([p["table"][c[0], c[1]] for c in cells])
                val_fin = val_fin + v * p["W_values"]
        val_fin = beta * val_fin  # exploration_model.py:87
        val_fin = val_fin - np.max(val_fin)  # :92
        p_soft = np.exp(val_fin) / np.sum(np.exp(val_fin))  # :93
        cdf = p_soft.cumsum()  # Generator.choice(a, p=p)
        cdf /= cdf[-1]
        k = int(cdf.searchsorted(draws[t], side="right"))
        new = cand[k]
        moving = bool(np.all(prev == np.array(new)))  # :101
        prev = np.array(new)
        mouse.move_to(new[0], new[1])  # experiment.py:51
        hist_pos.append((mouse.x, mouse.y))
        hist_ori.append(mouse.ori)
        choice.append(idxs[k])
        if record_probs:
            probs.append(p_soft)
    return {"hist_pos": np.array(hist_pos), "hist_ori": np.array(hist_ori), "choice": np.array(choice),
            "probs": probs, "status": status}


def metrics_of_history(maze, hist_pos, percentage=80):
    """discretize_positions / occupancy_map / it_perc_visited_maze / motion_analysis / tile_analysis counts
    (analysis/metrics.py:192-248, 425-442, 697-731) of one history."""
    disc = [maze.cont2disc(px, py) for px, py in hist_pos]  # :201
    occ = np.zeros((maze.X, maze.Y))
    for c in disc:  # :213-215
        occ[c[0], c[1]] += 1
    nvis = len(maze.vis_cells)
    seen, n_seen, it = [], 0, len(disc)  # :707-721
    for idx, c in enumerate(disc):
        if c not in seen:
            n_seen += 1
            seen.append(c)
            if (n_seen * 100) / nvis >= percentage:
                it = idx
                break
    count_moving = sum(1 for i in range(1, len(disc)) if disc[i] != disc[i - 1])  # :437-440
    tile_counts = np.zeros(8, dtype=np.int64)
    for c in disc:  # :236-239
        tile_counts[int(maze.tiles[c])] += 1
    return {"disc": disc, "occupancy": occ, "it_visit": it, "count_moving": count_moving, "tile_counts": tile_counts}


def pcg64_state_words(seed):
    """``(state_hi, state_lo, inc_hi, inc_lo)`` of ``np.random.default_rng(seed)`` (PCG64)."""
    st = np.random.default_rng(seed).bit_generator.state["state"]
    m = (1 << 64) - 1
    return np.array([st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)


# ============================================================================ model-based agent
# EXTENSION with no reference implementation: parity unpinned.  Composed from the
# reference's primitives: the maze geometry as the transition model (Maze.is_movement_possible,
# simulation/maze.py:512-527 -- what PositiveConditioning's look-ahead uses, rl.py:177-187), the
# `x += alpha * (target - x)` update form (rl.py:185-187) and the softmax of exploration_model.py:84-93.
def transition_table(maze, discrete_actions):
    """``next[S, A]``: compact state reached from state ``s`` by tile displacement ``a`` when the centre-to-
    centre movement is possible (maze.py:512-527), else -1."""
    acts = np.asarray(discrete_actions, dtype=np.int64).reshape(-1, 2)
    out = np.full((len(maze.vis_cells), len(acts)), -1, dtype=np.int64)
    for s, (x, y) in enumerate(maze.vis_cells):
        cx, cy = maze.centre((x, y))
        for a, (dx, dy) in enumerate(acts):
            c = (x + int(dx), y + int(dy))
            if maze.visitable(c) and maze.move_possible(cx, cy, *maze.centre(c)):
                out[s, a] = maze.state_of[c]
    return out


def run_model_based(tr, nxt, alpha, beta, gamma, n_sweeps=1, v0=None):
    """Learned reward model + synchronous value-iteration planning + softmax likelihood.

    Per online step (s -> s1, reward r):
      Q_a   = Rhat[next(s,a)] + gamma * V[next(s,a)]        for the available actions a, in action order
      logp  = softmax(beta * Q)[a_obs]  (log space; a_obs = first a with next(s,a) == s1, step skipped if none)
      Rhat[s1] += alpha * (r - Rhat[s1])
      n_sweeps x:  V'[u] = max_a (Rhat[next(u,a)] + gamma * V[next(u,a)])  for every state u;  V <- V'
    :return: (V[S], Rhat[S], ll)
    """
    S, A = nxt.shape
    V = np.zeros(S) if v0 is None else np.array(v0, dtype=np.float64)
    R = np.zeros(S)
    ll = 0.0
    avail = nxt >= 0
    safe = np.where(avail, nxt, 0)
    for s, s1, r in zip(tr["s"].tolist(), tr["s1"].tolist(), tr["r"].tolist()):
        acts = [a for a in range(A) if nxt[s, a] >= 0]
        if acts:
            q = [R[nxt[s, a]] + gamma * V[nxt[s, a]] for a in acts]
            obs = next((i for i, a in enumerate(acts) if nxt[s, a] == s1), None)
            if obs is not None:
                val = [beta * v for v in q]
                m = max(val)
                val = [v - m for v in val]
                ll = ll + (val[obs] - math.log(np_sum([math.exp(v) for v in val])))
        R[s1] += alpha * (r - R[s1])
        for _ in range(n_sweeps):
            q_all = np.where(avail, R[safe] + gamma * V[safe], -np.inf)
            best = q_all.max(axis=1)
            V = np.where(avail.any(axis=1), best, V)
    return V, R, ll


def run_q_learning(tr, nxt, alpha, beta, gamma, q0=None, orders=()):
    """Tabular Q-learning on an explicit Q[s, a] table + softmax likelihood (EXTENSION; the
    reference's learners keep state values only -- the update is PositiveConditioning's max-backup, rl.py:184-187,
    moved to state-action values; the softmax is exploration_model.py:84-93).

    Per step (s -> s1, reward r), a = first action with next(s, a) == s1 (steps without one are skipped):
      logp    = softmax(beta * Q[s, available actions])[a]        (log space, Q before the update; online pass only)
      Q[s, a] += alpha * (r + gamma * max_{a' available at s1} Q[s1, a'] - Q[s, a])     (bootstrap 0 if none)
    :param orders: replay index orders (see :func:`replay_orders`): update-only passes after the online pass
    :return: (Q[S, A], ll)
    """
    S, A = nxt.shape
    Q = np.zeros((S, A)) if q0 is None else np.array(q0, dtype=np.float64).reshape(S, A)
    ll = 0.0
    steps = list(zip(tr["s"].tolist(), tr["s1"].tolist(), tr["r"].tolist()))
    acts_of = [[a for a in range(A) if nxt[s, a] >= 0] for s in range(S)]

    def one(step, online):
        nonlocal ll
        s, s1, r = step
        acts = acts_of[s]
        obs = next((j for j, a in enumerate(acts) if nxt[s, a] == s1), None)
        if obs is None:
            return
        if online:
            val = [beta * float(Q[s, a]) for a in acts]
            m = max(val)
            val = [v - m for v in val]
            ll = ll + (val[obs] - math.log(np_sum([math.exp(v) for v in val])))
        boot = max(float(Q[s1, a]) for a in acts_of[s1]) if acts_of[s1] else 0.0
        a = acts[obs]
        q = float(Q[s, a])
        rpe = r + gamma * boot - q
        Q[s, a] = q + alpha * rpe

    for step in steps:
        one(step, True)
    for order in orders:
        for i in order.tolist():
            one(steps[i], False)
    return Q, ll


# ================================================================================= distances
# Restatement of analysis/statistics.py (pinned by the reference's test/test_statistics.py and test_fitness.py).
def normalized_distance(h1, h2):  # statistics.py:67-92
    h1, h2 = np.asarray(h1, dtype=np.float64), np.asarray(h2, dtype=np.float64)
    c1, c2 = np.sum(h1), np.sum(h2)
    if c1 == 0 and c2 != 0:
        nd = np.linalg.norm(h2 / c2)
    elif c2 == 0 and c1 != 0:
        nd = np.linalg.norm(h1 / c1)
    elif c2 == 0 and c1 == 0:
        nd = 0
    else:
        nd = np.linalg.norm(h1 / c1 - h2 / c2)
    return 1 if nd > 1 else nd


def manhattan_distance(x1, x2):  # statistics.py:95-103
    return np.linalg.norm(np.array(x1, dtype=np.float64) - np.array(x2, dtype=np.float64), ord=1)


def js_distance(c1, c2):  # statistics.py:10-54 on the count vectors h[0]
    from scipy.special import rel_entr
    p, q = np.asarray(c1, dtype=np.float64), np.asarray(c2, dtype=np.float64)
    p = 0 if np.sum(p) == 0 else p / np.sum(p)
    q = 0 if np.sum(q) == 0 else q / np.sum(q)
    m = (p + q) / 2.0
    js = np.sum(rel_entr(p, m)) + np.sum(rel_entr(q, m))
    d = np.sqrt(js / 2.0)
    return 1 if math.isinf(d) else d
