"""Golden fixture for StandardExperiment.run called several times on the same objects, produced by the UNMODIFIED
reference (`python tests/golden/make_golden_split_runs.py` in the build container, where /root/reference exists).

For three mice of sim_golden.npz the reference runs ``run(1); run(249); run(350)`` on ONE experiment.  The model, the
components and the mouse carry on between the calls (experiment.py:33-51 keeps no per-run state), so the moves equal
those of the single 600-step run stored in sim_golden.npz; only ``enable_history(True)`` (experiment.py:32 ->
mouse.cpp:38-41, 55-61) restarts the history at the mouse's INITIAL pose.  Stored: the history of the last call and the
state the objects are left in (visit map, familiarity, mental state, ``_moving``, ``_prev_pos``, next uniform draw).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle")]

import make_golden_sim as G  # noqa: E402  (imports the reference through oracle/refharness.py)
from make_golden_sim import (AnxietyComponent, BehaviouralModel, BiomechanicalCostComponent,  # noqa: E402
                             BiomechPersistenceComponent, ExperienceComponent, Maze, Mouse, StandardExperiment,
                             ValueTableComponent)

LEGS = (1, 249, 350)
MICE = (0, 3, 5)


def main():
    g = np.load(os.path.join(HERE, "sim_golden.npz"))
    maze = Maze.from_file(G.M9_FILE, size_tile=G.SIZE_TILE)
    out = {"legs": np.array(LEGS), "mice": np.array(MICE)}
    hist_pos, hist_ori, visits, state, next_draw = [], [], [], [], []
    for m in MICE:
        p = dict(zip(G.PARAM_NAMES, g["params"][m]))
        comps = [AnxietyComponent(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"]),
                 BiomechanicalCostComponent(G.A8_UNITS, W_bcost=p["W_bcost"], k=p["k"], gamma=p["gamma"]),
                 ExperienceComponent(maze.shape, W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                     kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"],
                                     theta_home=p["theta_home"]),
                 BiomechPersistenceComponent(G.A8_UNITS, W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"]),
                 ValueTableComponent(g["v_table"], W_values=p["W_values"])]
        model = BehaviouralModel(comps, tuple(int(v) for v in g["init_disc"][m]), beta=p["beta"], seed=int(g["seeds"][m]))
        mouse = Mouse(g["init_pose"][m][:2], g["init_pose"][m][2], g["actions"])
        exp = StandardExperiment(model, maze, mouse)
        for leg in LEGS:
            exp.run(leg)
        hp, ho = mouse.history
        hp, ho = np.array(hp), np.array(ho)
        assert len(hp) == LEGS[-1] + 1 and np.array_equal(hp[0], g["init_pose"][m][:2])
        assert np.array_equal(hp[1:], g["hist_pos"][m][sum(LEGS[:-1]) + 1:])  # same moves as the single run
        hist_pos.append(hp)
        hist_ori.append(ho)
        visits.append(np.array(comps[2]._maze_map_exp, dtype=np.float64))
        prev = np.asarray(model._prev_pos, dtype=np.float64).reshape(2)
        state.append([float(bool(model._moving)), float(comps[2]._familiarity), float(int(comps[2]._ment_state)),
                      prev[0], prev[1]])
        next_draw.append(model._rng.random())
    out.update(hist_pos=np.array(hist_pos), hist_ori=np.array(hist_ori), visits=np.array(visits),
               state=np.array(state), next_draw=np.array(next_draw))
    np.savez_compressed(os.path.join(HERE, "split_runs_golden.npz"), **out)
    print("split_runs_golden.npz:", {k: np.shape(v) for k, v in out.items()})
    print(out["state"])


if __name__ == "__main__":
    main()
