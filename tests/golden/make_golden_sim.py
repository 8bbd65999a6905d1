"""Golden fixtures of the forward-simulation path, produced by the UNMODIFIED reference.

Called by make_golden.py (`python tests/golden/make_golden.py sim`).  For a handful of mice with
different parameter sets and seeds it runs the reference's

    StandardExperiment(BehaviouralModel([Anxiety, BiomechanicalCost, Experience, BiomechPersistence,
                                         ValueTable], init_disc, beta, seed), maze, Mouse).run(T)

and stores: the parameters, the uniform draws the model's generator produced (a twin generator seeded
identically: Generator.choice(p=...) consumes exactly one random() per call), the full history
(positions, orientations), the per-step softmax probabilities observed through a recording component,
and the reference's own metrics (occupancy_map, it_perc_visited_maze, motion_analysis, tile_analysis)
of that history.
"""
import os

import numpy as np

import refharness

refharness.install()

from navigation_model.analysis.metrics import (discretize_positions, it_perc_visited_maze, motion_analysis,  # noqa: E402
                                               occupancy_map, tile_analysis)
from navigation_model.simulation.behavioural_components import (AnxietyComponent, BiomechanicalCostComponent,  # noqa: E402
                                                                BiomechPersistenceComponent, ExperienceComponent,
                                                                ValueTableComponent)
from navigation_model.simulation.experiment import StandardExperiment  # noqa: E402
from navigation_model.simulation.exploration_model import BehaviouralModel  # noqa: E402
from navigation_model.simulation.maze import Maze  # noqa: E402
from navigation_model.simulation.mouse import Mouse  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
M9_FILE = os.path.join(refharness.REFERENCE_ROOT, "test", "data", "maze_layout.txt")
SIZE_TILE = 0.111
A8_UNITS = [(0, 0), (1, 0), (2, 0), (1, 1), (1, -1), (0, 1), (0, -1), (-1, 0)]

PARAM_NAMES = ["beta", "W_anx", "p1", "p2", "p3", "W_bcost", "k", "gamma", "W_exp", "xi", "kappa_home", "kappa_explo",
               "theta_explo", "theta_home", "W_bper", "bp", "Wm", "Wnm", "W_values"]


def draw_params(rng):
    """Uniform inside the descriptor bounds (behavioural_components.py:205, 317-318, 420-424)."""
    p = {
        "beta": rng.uniform(0.5, 6.0),
        "W_anx": rng.uniform(0, 10), "p1": rng.uniform(0, 1), "p2": rng.uniform(0, 1), "p3": rng.uniform(0, 1),
        "W_bcost": rng.uniform(0, 10), "k": rng.uniform(0.1, 4), "gamma": rng.uniform(0.2, 2),
        "W_exp": rng.uniform(0, 10), "xi": rng.uniform(0.006, 0.6), "kappa_home": rng.uniform(0.01, 5),
        "kappa_explo": rng.uniform(0.01, 5), "theta_explo": rng.uniform(80, 120), "theta_home": rng.uniform(30, 70),
        "W_bper": rng.uniform(0, 10), "bp": rng.uniform(0, 1), "Wm": rng.uniform(0, 2), "Wnm": rng.uniform(0, 2),
        "W_values": rng.uniform(0, 10),
    }
    return p


def make_sim():
    maze = Maze.from_file(M9_FILE, size_tile=SIZE_TILE)
    actions = np.array(A8_UNITS, dtype=np.float64) * SIZE_TILE
    T = 600
    n_mice = 6
    rng = np.random.default_rng(2024)
    v_table = rng.random(maze.shape) * 3.0 - 1.0
    out = {"param_names": np.array(PARAM_NAMES), "actions": actions, "actions_units": np.array(A8_UNITS),
           "v_table": v_table, "T": np.array(T)}
    params, seeds, init_pose, init_disc = [], [], [], []
    hist_pos, hist_ori, draws, occ, it80, it50, cmov, tcount = [], [], [], [], [], [], [], []
    starts = [(1, 1), (7, 4), (1, 6), (7, 7), (2, 8), (0, 0)]
    for m in range(n_mice):
        p = draw_params(rng)
        seed = 1000 + m
        start = starts[m]
        # start at the tile centre (mouse 0..2) or at a jittered point with a non-trivial heading
        pos = np.array(maze.get_tile_centre(start))
        ori = 0.0
        if m >= 3:
            pos = pos + rng.uniform(-0.03, 0.03, 2)
            ori = rng.uniform(-np.pi, np.pi)
        comps = [
            AnxietyComponent(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"]),
            BiomechanicalCostComponent(A8_UNITS, W_bcost=p["W_bcost"], k=p["k"], gamma=p["gamma"]),
            ExperienceComponent(maze.shape, W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"],
                                theta_home=p["theta_home"]),
            BiomechPersistenceComponent(A8_UNITS, W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"]),
            ValueTableComponent(v_table, W_values=p["W_values"]),
        ]
        model = BehaviouralModel(comps, start, beta=p["beta"], seed=seed)
        mouse = Mouse(pos, ori, actions)
        StandardExperiment(model, maze, mouse).run(T)
        hp, ho = mouse.history
        hp = np.array(hp)
        disc = discretize_positions(hp, maze)
        params.append([p[k] for k in PARAM_NAMES])
        seeds.append(seed)
        init_pose.append([pos[0], pos[1], ori])
        init_disc.append(start)
        hist_pos.append(hp)
        hist_ori.append(np.array(ho))
        draws.append(np.random.default_rng(seed).random(T))
        occ.append(occupancy_map(disc, maze))
        it80.append(it_perc_visited_maze(disc, maze, percentage=80))
        it50.append(it_perc_visited_maze(disc, maze, percentage=50))
        cmov.append(motion_analysis(disc)[0])
        counts = tile_analysis(disc, maze)[2]
        tcount.append([counts[t] for t in sorted(counts, key=int)])
        if m == 0:
            out["bcost_table_m0"] = np.array(comps[1]._cachedval)
            out["vtable_norm"] = np.array(comps[4]._v_table)
    out.update(params=np.array(params), seeds=np.array(seeds), init_pose=np.array(init_pose),
               init_disc=np.array(init_disc), hist_pos=np.array(hist_pos), hist_ori=np.array(hist_ori),
               draws=np.array(draws), occupancy=np.array(occ), it_visit_80=np.array(it80),
               it_visit_50=np.array(it50), count_moving=np.array(cmov), tile_counts=np.array(tcount))
    # per-step probabilities of one decision sequence (mouse 1), captured without touching the reference:
    # re-run with the model's generator replaced by a recording proxy
    out.update(_probabilities_of_mouse(maze, actions, v_table, out, 1, T))
    np.savez_compressed(os.path.join(HERE, "sim_golden.npz"), **out)
    print("sim_golden.npz:", {k: np.shape(v) for k, v in out.items()})


class _RecordingRng:
    """Wraps a numpy Generator and records the probability vectors passed to choice()."""

    def __init__(self, rng):
        self._rng = rng
        self.probs = []

    def choice(self, a, p=None):
        self.probs.append(np.array(p))
        return self._rng.choice(a, p=p)


def _probabilities_of_mouse(maze, actions, v_table, out, m, T):
    p = dict(zip(PARAM_NAMES, out["params"][m]))
    comps = [
        AnxietyComponent(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"]),
        BiomechanicalCostComponent(A8_UNITS, W_bcost=p["W_bcost"], k=p["k"], gamma=p["gamma"]),
        ExperienceComponent(maze.shape, W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                            kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"], theta_home=p["theta_home"]),
        BiomechPersistenceComponent(A8_UNITS, W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"]),
        ValueTableComponent(v_table, W_values=p["W_values"]),
    ]
    model = BehaviouralModel(comps, tuple(int(v) for v in out["init_disc"][m]), beta=p["beta"],
                             seed=int(out["seeds"][m]))
    rec = _RecordingRng(model._rng)
    model._rng = rec
    mouse = Mouse(out["init_pose"][m][:2], out["init_pose"][m][2], actions)
    StandardExperiment(model, maze, mouse).run(T)
    assert np.array_equal(np.array(mouse.history[0]), out["hist_pos"][m])
    probs = np.zeros((T, 8))
    ncand = np.zeros(T, dtype=np.int64)
    for t, pv in enumerate(rec.probs):
        probs[t, :len(pv)] = pv
        ncand[t] = len(pv)
    return {"m1_probs": probs, "m1_ncand": ncand}


if __name__ == "__main__":
    make_sim()
