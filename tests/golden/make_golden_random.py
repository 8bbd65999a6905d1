"""Golden fixtures on RANDOM irregular mazes, produced by the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_random.py

For three seeded 7 x 11 layouts with scattered void tiles and a random subset of the eight actions it stores
  * what ConditioningExperiment.run extracts from a seeded session with PositiveConditioning (transitions and the
    candidate next states of every step) and the V-tables of TD0 / PositiveConditioning / NegativeConditioning with one
    shuffled replay pass;
  * three StandardExperiment.run histories (all five components) with the draws of a twin generator.
tests/test_oracle_golden.py::test_random_mazes_match_reference replays them through the oracle -- the ray / AABB
walks, tile lookups and nearest-visitable searches of the reference are thereby pinned on irregular geometry too.
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

from navigation_model.simulation.behavioural_components import (AnxietyComponent, BiomechanicalCostComponent,  # noqa: E402
                                                                BiomechPersistenceComponent, ExperienceComponent,
                                                                ValueTableComponent)
from navigation_model.simulation.experiment import StandardExperiment  # noqa: E402
from navigation_model.simulation.exploration_model import BehaviouralModel  # noqa: E402
from navigation_model.simulation.maze import Maze  # noqa: E402
from navigation_model.simulation.mouse import Mouse  # noqa: E402
from navigation_model.simulation.rl import (ConditioningExperiment, NegativeConditioning,  # noqa: E402
                                            PositiveConditioning, ShuffledReplays, TD0)

from make_golden import synthetic_session  # noqa: E402  (the seeded random-walk recipe, on the reference's Maze)
from make_golden_sim import PARAM_NAMES, draw_params  # noqa: E402

X, Y, SIZE_TILE = 7, 11, 0.1
A8_UNITS = np.array([(0, 0), (1, 0), (2, 0), (1, 1), (1, -1), (0, 1), (0, -1), (-1, 0)], dtype=np.float64)
SEEDS = (21, 22, 23)


def random_layout(rng):
    rows = [[("V" if rng.random() < 0.22 else "CWM"[rng.integers(3)]) for _ in range(Y)] for _ in range(X)]
    rows[1][1], rows[1][2] = "M", "W"
    return "\n".join("".join(r) for r in rows)


def main():
    out = {"seeds": np.array(SEEDS), "size_tile": np.array(SIZE_TILE)}
    for seed in SEEDS:
        rng = np.random.default_rng(seed)
        layout = random_layout(rng)
        maze = Maze(layout, "", size_tile=SIZE_TILE)
        keep = np.concatenate([[0], 1 + np.sort(rng.permutation(7)[:int(rng.integers(3, 8))])])
        units = A8_UNITS[keep]
        actions = units * SIZE_TILE
        k = f"s{seed}_"
        out[k + "layout"], out[k + "units"] = np.array(layout), units
        # ---- value learning ---------------------------------------------------------------------------------
        traj, rew = synthetic_session(maze, 500 + seed, 260, goal=(1, 2), jitter=0.03)
        out[k + "traj"], out[k + "rew"] = traj, rew
        params = [(0.3, 0.9), (0.9, 0.55)]
        out[k + "params"] = np.array(params)
        for name in ("td0", "positive", "negative"):
            tabs = []
            for alpha, gamma in params:
                mouse = Mouse(traj[0], 0.0, actions)
                alg = {"td0": lambda: TD0(alpha, gamma), "positive": lambda: PositiveConditioning(alpha, gamma, mouse, maze),
                       "negative": lambda: NegativeConditioning(alpha, gamma, mouse, maze)}[name]()
                rep = ShuffledReplays(n_times=1, seed=seed)
                exp = ConditioningExperiment(maze, alg, rep)
                r = -rew if name == "negative" else rew
                exp.run(list(traj), [0.0] * len(traj), list(r), do_replays=True)
                tabs.append(np.array(exp.v_table))
            out[k + name + "_V"] = np.array(tabs)
        # the transitions as the reference extracted them: a run WITHOUT replays (offline_replays shuffles the buffer
        # in place, rl.py:420)
        mouse = Mouse(traj[0], 0.0, actions)
        rep = ShuffledReplays(n_times=0, seed=None)
        ConditioningExperiment(maze, PositiveConditioning(0.3, 0.9, mouse, maze), rep).run(
            list(traj), [0.0] * len(traj), list(rew))
        trs = [u.transition for u in rep._buffer]
        out[k + "tr_s"] = np.array([t.s for t in trs])
        out[k + "tr_s1"] = np.array([t.s1 for t in trs])
        out[k + "tr_ncand"] = np.array([len(t.extra["next_states"]) for t in trs])
        cand = np.full((len(trs), 8, 2), -1, dtype=np.int64)
        for i, t in enumerate(trs):
            for j, c in enumerate(t.extra["next_states"]):
                cand[i, j] = c
        out[k + "tr_cand"] = cand
        # ---- forward simulation -----------------------------------------------------------------------------
        T, n_mice = 150, 3
        v_table = rng.random(maze.shape) * 2.0 - 0.5
        out[k + "v_table"] = v_table
        cells = [tuple(int(v) for v in c) for c in np.argwhere(np.array([[maze.is_visitable(x, y) for y in range(Y)]
                                                                         for x in range(X)]))]
        P, pose, disc, hp_all, ho_all, draws = [], [], [], [], [], []
        for m in range(n_mice):
            p = draw_params(rng)
            start = cells[int(rng.integers(len(cells)))]
            pos = np.array(maze.get_tile_centre(start)) + rng.uniform(-0.03, 0.03, 2)
            ori = rng.uniform(-np.pi, np.pi)
            comps = [AnxietyComponent(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"]),
                     BiomechanicalCostComponent(units, W_bcost=p["W_bcost"], k=p["k"], gamma=p["gamma"]),
                     ExperienceComponent(maze.shape, W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                         kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"],
                                         theta_home=p["theta_home"]),
                     BiomechPersistenceComponent(units, W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"]),
                     ValueTableComponent(v_table, W_values=p["W_values"])]
            model = BehaviouralModel(comps, start, beta=p["beta"], seed=1000 * seed + m)
            mouse = Mouse(pos, ori, actions)
            StandardExperiment(model, maze, mouse).run(T)
            hp, ho = mouse.history
            P.append([p[n] for n in PARAM_NAMES])
            pose.append([pos[0], pos[1], ori])
            disc.append(start)
            hp_all.append(np.array(hp))
            ho_all.append(np.array(ho))
            draws.append(np.random.default_rng(1000 * seed + m).random(T))
        out.update({k + "sim_params": np.array(P), k + "sim_pose": np.array(pose), k + "sim_disc": np.array(disc),
                    k + "sim_hist_pos": np.array(hp_all), k + "sim_hist_ori": np.array(ho_all),
                    k + "sim_draws": np.array(draws)})
    out["param_names"] = np.array(PARAM_NAMES)
    np.savez_compressed(os.path.join(HERE, "random_mazes_golden.npz"), **out)
    print("random_mazes_golden.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
