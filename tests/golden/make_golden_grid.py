"""Golden fixtures for the GridTiles tile system and AnxietyComponent_Grid, produced by the UNMODIFIED reference:

    python tests/golden/make_golden_grid.py

A 9 x 10 grid maze (corridors of CENTRE tiles between VOID blocks, DECISION_POINT tiles at the junctions, WALL / CORNER
border; ``Maze(..., tile_system=GridTiles)``, maze.py:62-104) and four mice driven by
``BehaviouralModel([AnxietyComponent_Grid, BiomechanicalCostComponent, ExperienceComponent,
BiomechPersistenceComponent, ValueTableComponent])`` (behavioural_components.py:315-346: the DECISION_POINT tiles
take the value ``p4``) through ``StandardExperiment.run(300)``.  Stored: parameters (``p4`` included), start poses,
histories, the draws of a twin generator, per-step probabilities of mouse 0 (recording generator proxy) and the
reference's tile_analysis counts (which include the DECISION_POINT type).
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

from navigation_model.analysis.metrics import discretize_positions, occupancy_map, tile_analysis  # noqa: E402
from navigation_model.simulation.behavioural_components import (AnxietyComponent_Grid, BiomechanicalCostComponent,  # noqa: E402
                                                                BiomechPersistenceComponent, ExperienceComponent,
                                                                ValueTableComponent)
from navigation_model.simulation.experiment import StandardExperiment  # noqa: E402
from navigation_model.simulation.exploration_model import BehaviouralModel  # noqa: E402
from navigation_model.simulation.maze import GridTiles, Maze  # noqa: E402
from navigation_model.simulation.mouse import Mouse  # noqa: E402

from make_golden_sim import PARAM_NAMES, _RecordingRng, draw_params  # noqa: E402

LAYOUT = "\n".join([
    "CWWWWWWWWC",
    "WDMMDMMMDW",
    "WMVVMVVVMW",
    "WMVVMVVVMW",
    "WDMMDMMMDW",
    "WMVVMVVVMW",
    "WMVVMVVVMW",
    "WDMMDMMMDW",
    "CWWWWWWWWC",
])
SIZE_TILE = 0.1
A8_UNITS = [(0, 0), (1, 0), (2, 0), (1, 1), (1, -1), (0, 1), (0, -1), (-1, 0)]
NAMES = PARAM_NAMES + ["p4"]


def components(maze, p, v_table):
    return [AnxietyComponent_Grid(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"], p4=p["p4"]),
            BiomechanicalCostComponent(A8_UNITS, W_bcost=p["W_bcost"], k=p["k"], gamma=p["gamma"]),
            ExperienceComponent(maze.shape, W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"], theta_home=p["theta_home"]),
            BiomechPersistenceComponent(A8_UNITS, W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"]),
            ValueTableComponent(v_table, W_values=p["W_values"])]


def main():
    maze = Maze(LAYOUT, "", size_tile=SIZE_TILE, tile_system=GridTiles)
    assert maze.get_tile(1, 1) == GridTiles.TileType.DECISION_POINT
    actions = np.array(A8_UNITS, dtype=np.float64) * SIZE_TILE
    rng = np.random.default_rng(77)
    v_table = rng.random(maze.shape) * 2.0 - 0.5
    T, n_mice = 300, 4
    starts = [(1, 1), (4, 4), (7, 8), (1, 5)]
    out = {"layout": np.array(LAYOUT), "size_tile": np.array(SIZE_TILE), "units": np.array(A8_UNITS),
           "v_table": v_table, "param_names": np.array(NAMES)}
    P, pose, hp_all, ho_all, draws, occ, tcount = [], [], [], [], [], [], []
    for m in range(n_mice):
        p = draw_params(rng)
        p["p4"] = rng.uniform(-1.0, 2.0)
        p["beta"] = rng.uniform(0.05, 0.5)  # mild inverse temperature: the mice keep moving through the junctions
        pos = np.array(maze.get_tile_centre(starts[m])) + rng.uniform(-0.03, 0.03, 2)
        ori = rng.uniform(-np.pi, np.pi)
        model = BehaviouralModel(components(maze, p, v_table), starts[m], beta=p["beta"], seed=500 + m)
        rec = None
        if m == 0:
            rec = _RecordingRng(model._rng)
            model._rng = rec
        mouse = Mouse(pos, ori, actions)
        StandardExperiment(model, maze, mouse).run(T)
        hp, ho = mouse.history
        hp = np.array(hp)
        disc = discretize_positions(hp, maze)
        P.append([p[n] for n in NAMES])
        pose.append([pos[0], pos[1], ori])
        hp_all.append(hp)
        ho_all.append(np.array(ho))
        draws.append(np.random.default_rng(500 + m).random(T))
        occ.append(occupancy_map(disc, maze))
        counts = tile_analysis(disc, maze)[2]
        tcount.append([counts[t] for t in sorted(counts, key=int)])
        if rec is not None:
            probs = np.zeros((T, 8))
            for t, pv in enumerate(rec.probs):
                probs[t, :len(pv)] = pv
            out["m0_probs"] = probs
    out.update(params=np.array(P), pose=np.array(pose), init_disc=np.array(starts), hist_pos=np.array(hp_all),
               hist_ori=np.array(ho_all), draws=np.array(draws), occupancy=np.array(occ), tile_counts=np.array(tcount))
    np.savez_compressed(os.path.join(HERE, "grid_golden.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()})
    print("tiles visited per mouse:", [int((o > 0).sum()) for o in out["occupancy"]])
    print("tile counts (VOID, WALL, CORNER, CENTRE, DECISION_POINT):", out["tile_counts"].tolist())


if __name__ == "__main__":
    main()
