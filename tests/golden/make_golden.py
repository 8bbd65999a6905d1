"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [fit] [sim]

The reference package is imported from /root/reference/src through oracle/refharness.py (matplotlib /
seaborn stubs + an Eigen-free restatement of the C++ Mouse, which cannot be compiled offline).  Nothing
from this repo's product code or oracle is used to compute the stored values: inputs are built with
numpy and the reference's own classes, outputs are what the reference returns.

Fixtures (npz, a few hundred KB in total):
  fit_golden.npz  ConditioningExperiment.run (+ShuffledReplays) V-tables for TD0 / PositiveConditioning /
                  NegativeConditioning, the transitions the reference extracted, known-answer scalars
                  of the full T=5000 benchmark session
  sim_golden.npz  StandardExperiment.run trajectories with every random draw recorded, per-step softmax
                  probabilities, occupancy metrics (written by the `sim` stage)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refharness  # noqa: E402

if not refharness.install():
    raise SystemExit("reference not available at /root/reference")

from navigation_model.simulation.maze import Maze  # noqa: E402  (the reference)
from navigation_model.simulation.mouse import Mouse  # noqa: E402
from navigation_model.simulation.rl import (ConditioningExperiment, NegativeConditioning,  # noqa: E402
                                            PositiveConditioning, ShuffledReplays, TD0)

M9_FILE = os.path.join(refharness.REFERENCE_ROOT, "test", "data", "maze_layout.txt")
SIZE_TILE = 0.111
A8 = np.array([(0, 0), (1, 0), (2, 0), (1, 1), (1, -1), (0, 1), (0, -1), (-1, 0)], dtype=np.float64) * SIZE_TILE

# the 7x7 maze of test/test_maze.py:12-18 (has an interior VOID tile)
M7_LAYOUT = "CWWWWWC\nWMMMMMW\nWMWWWMW\nWMWVWMW\nWMWWWMW\nWMMMMMW\nCWWWWWC"


def synthetic_session(maze, session_id, T, start=(1, 1), goal=(7, 4), jitter=0.04, outside_every=0):
    """Seeded random-walk session (SURVEY 8c recipe): next tile drawn with integers() first, then the
    jitter of the current point.  ``outside_every`` > 0 pushes every n-th point far out of its tile so
    that get_closest_visitable_cont's nearest-centre branch is exercised."""
    rng = np.random.default_rng(session_id)
    tile = tuple(start)
    traj, rew = [], []
    steps = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1)]
    for t in range(T):
        nbrs = [(tile[0] + dx, tile[1] + dy) for dx, dy in steps if maze.is_visitable(tile[0] + dx, tile[1] + dy)]
        nxt = nbrs[rng.integers(len(nbrs))]
        p = np.array(maze.get_tile_centre(tile)) + rng.uniform(-jitter, jitter, 2)
        if outside_every and t % outside_every == outside_every - 1:
            p = p + np.array([0.35, -0.27])
        traj.append(p)
        rew.append(1.0 if maze.cont2disc(p) == tuple(goal) else 0.0)
        tile = nxt
    return np.array(traj), np.array(rew)


def run_reference(maze, alg_name, alpha, gamma, traj, rew, n_times, seed, v_init=None, extra_do_replays=False,
                  actions=None):
    mouse = Mouse(traj[0], 0.0, A8 if actions is None else actions)
    if alg_name == "td0":
        alg = TD0(alpha, gamma)
    elif alg_name == "positive":
        alg = PositiveConditioning(alpha, gamma, mouse, maze)
    else:
        alg = NegativeConditioning(alpha, gamma, mouse, maze)
    rep = ShuffledReplays(n_times=n_times, seed=seed)
    exp = ConditioningExperiment(maze, alg, rep, v_table=None if v_init is None else v_init.copy())
    exp.run(list(traj), [0.0] * len(traj), list(rew), do_replays=n_times > 0)
    if extra_do_replays:
        exp.do_replays()
    return np.array(exp.v_table), rep, mouse


def make_fit():
    out = {}
    maze = Maze.from_file(M9_FILE, size_tile=SIZE_TILE)
    params = [(0.1, 0.9), (0.5, 0.5), (1.0, 0.99), (0.013, 0.731)]
    out["params"] = np.array(params)
    out["A8"] = A8
    # --- case "m9": T = 1200 session on the 9x9 test maze
    T = 1200
    traj, rew = synthetic_session(maze, 3, T)
    neg_rew = -rew
    out["m9_traj"], out["m9_rew"] = traj, rew
    for alg in ("td0", "positive", "negative"):
        r = neg_rew if alg == "negative" else rew
        for n_times in (0, 3):
            tabs = []
            for (a, g) in params:
                V, rep, mouse = run_reference(maze, alg, a, g, traj, r, n_times, seed=11)
                tabs.append(V)
            out[f"m9_{alg}_rep{n_times}_V"] = np.array(tabs)
    # transitions as the reference extracted them (PositiveConditioning run, no replays)
    _, rep, mouse = run_reference(maze, "positive", 0.1, 0.9, traj, rew, 0, seed=None)
    trs = [u.transition for u in rep._buffer]
    out["m9_tr_s"] = np.array([t.s for t in trs])
    out["m9_tr_s1"] = np.array([t.s1 for t in trs])
    out["m9_tr_r"] = np.array([t.r for t in trs])
    out["m9_tr_ncand"] = np.array([len(t.extra["next_states"]) for t in trs])
    cand = np.full((len(trs), 8, 2), -1, dtype=np.int64)
    for i, t in enumerate(trs):
        for k, c in enumerate(t.extra["next_states"]):
            cand[i, k] = c
    out["m9_tr_cand"] = cand
    out["m9_mouse_final"] = np.array([mouse.position[0], mouse.position[1], mouse.orientation])
    # continued learning: start from a given table (rl.py:63-64), then a second do_replays() call
    rng = np.random.default_rng(5)
    v0 = rng.random(maze.shape)
    out["m9_v0"] = v0
    out["m9_td0_v0_V"] = run_reference(maze, "td0", 0.2, 0.8, traj, rew, 2, seed=4, v_init=v0)[0]
    out["m9_positive_v0_V"] = run_reference(maze, "positive", 0.2, 0.8, traj, rew, 2, seed=4, v_init=v0,
                                            extra_do_replays=True)[0]
    # --- case "m7": 7x7 maze with an interior void and points pushed outside the maze
    maze7 = Maze(M7_LAYOUT, "", size_tile=0.2)
    a7 = A8 / SIZE_TILE * 0.2
    # (a) TD0 with every 7th point pushed outside the maze -> nearest-centre branch of maze.py:483-484
    traj7, rew7 = synthetic_session(maze7, 8, 600, start=(1, 1), goal=(5, 5), jitter=0.07, outside_every=7)
    out["m7_traj"], out["m7_rew"] = traj7, rew7
    out["m7_layout"] = np.array(M7_LAYOUT)
    out["m7_td0_rep2_V"] = np.array([run_reference(maze7, "td0", a, g, traj7, rew7, 2, seed=21, actions=a7)[0]
                                     for (a, g) in params])
    _, rep7, _ = run_reference(maze7, "td0", 0.1, 0.9, traj7, rew7, 0, seed=None, actions=a7)
    out["m7_tr_s"] = np.array([u.transition.s for u in rep7._buffer])
    out["m7_tr_s1"] = np.array([u.transition.s1 for u in rep7._buffer])
    # (b) look-ahead learners with jitter larger than half a tile (points spill into neighbouring tiles,
    #     the interior void and over the border)
    traj7b, rew7b = synthetic_session(maze7, 9, 600, start=(1, 1), goal=(5, 5), jitter=0.12)
    out["m7b_traj"], out["m7b_rew"] = traj7b, rew7b
    for alg in ("positive", "negative"):
        r = -rew7b if alg == "negative" else rew7b
        out[f"m7b_{alg}_rep2_V"] = np.array([run_reference(maze7, alg, a, g, traj7b, r, 2, seed=21, actions=a7)[0]
                                             for (a, g) in params])
    # --- known answers of the full benchmark session (session 0, T = 5000)
    trajK, rewK = synthetic_session(maze, 0, 5000)
    VK, _, _ = run_reference(maze, "td0", 0.1, 0.9, trajK, rewK, 0, seed=None)
    out["kat_td0_V"] = VK
    out["kat_rewarded_steps"] = np.array(int(rewK.sum()))
    out["kat_traj_head"] = trajK[:16]
    VP, _, _ = run_reference(maze, "positive", 0.1, 0.9, trajK, rewK, 1, seed=0)
    out["kat_positive_rep1_V"] = VP
    np.savez_compressed(os.path.join(HERE, "fit_golden.npz"), **out)
    print("fit_golden.npz:", {k: np.shape(v) for k, v in out.items()})


def make_imaginary():
    """ShuffledReplays(imaginary=True) (rl.py:388-418): transitions sampled from the maze with the stdlib `random`
    module, made replayable by seeding it right before the run.  Stores the V-tables and, for one run, the
    transitions the strategy sampled (recorded by wrapping the algorithm's update -- the reference is not modified)."""
    import random
    out = {}
    maze = Maze.from_file(M9_FILE, size_tile=SIZE_TILE)
    T = 400
    traj, rew = synthetic_session(maze, 6, T, goal=(1, 2))  # a goal next to the start: rewarded often
    out["traj"], out["rew"] = traj, rew
    params = [(0.1, 0.9), (0.6, 0.7)]
    out["params"] = np.array(params)
    out["random_seed"] = np.array(4242)
    for alg_name in ("td0", "positive", "negative"):
        r = -rew if alg_name == "negative" else rew
        tabs = []
        for (a, g) in params:
            mouse = Mouse(traj[0], 0.0, A8)
            alg = {"td0": lambda: TD0(a, g), "positive": lambda: PositiveConditioning(a, g, mouse, maze),
                   "negative": lambda: NegativeConditioning(a, g, mouse, maze)}[alg_name]()
            rep = ShuffledReplays(n_times=2, seed=1, imaginary=True, maze=maze, possible_actions=A8)
            exp = ConditioningExperiment(maze, alg, rep)
            random.seed(4242)
            exp.run(list(traj), [0.0] * len(traj), list(r), do_replays=True)
            exp.do_replays()  # a second round of imaginary replays continues the stdlib stream
            tabs.append(np.array(exp.v_table))
        out[f"{alg_name}_V"] = np.array(tabs)
    # the sampled transitions of one run (positive, first parameter set, first do_replays only)
    mouse = Mouse(traj[0], 0.0, A8)
    alg = PositiveConditioning(0.1, 0.9, mouse, maze)
    rep = ShuffledReplays(n_times=2, seed=1, imaginary=True, maze=maze, possible_actions=A8)
    exp = ConditioningExperiment(maze, alg, rep)
    random.seed(4242)
    exp.run(list(traj), [0.0] * len(traj), list(rew), do_replays=False)
    seen = []
    plain_update = alg.update

    def recording_update(t):
        u = plain_update(t)
        seen.append((t.s, t.s1, t.r, t.ori, [tuple(c) for c in t.extra["next_states"]]))
        return u

    alg.update = recording_update
    exp.do_replays()
    out["im_s"] = np.array([x[0] for x in seen])
    out["im_s1"] = np.array([x[1] for x in seen])
    out["im_r"] = np.array([x[2] for x in seen], dtype=np.float64)
    out["im_ori"] = np.array([x[3] for x in seen])
    out["im_ncand"] = np.array([len(x[4]) for x in seen])
    np.savez_compressed(os.path.join(HERE, "imaginary_golden.npz"), **out)
    print("imaginary_golden.npz:", {k: np.shape(v) for k, v in out.items()}, "rewarded:", int((out["im_r"] != 0).sum()))


if __name__ == "__main__":
    stages = sys.argv[1:] or ["fit", "sim"]
    if "imaginary" in stages:
        make_imaginary()
    if "fit" in stages:
        make_fit()
    if "sim" in stages:
        try:
            from make_golden_sim import make_sim
        except ImportError:
            make_sim = None
        if make_sim is not None:
            make_sim()
