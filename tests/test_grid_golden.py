"""GridTiles + AnxietyComponent_Grid (maze.py:62-104, behavioural_components.py:315-346): outputs of the UNMODIFIED
reference (tests/golden/grid_golden.npz, written by make_golden_grid.py) against the oracle and the host plugin loop
on the CPU, and against the CUDA simulation kernel (the ``p4`` slot of its anxiety table, tile code 4) on the GPU."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O

from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "grid_golden.npz"), allow_pickle=False))


def _params(g):
    names = [str(s) for s in g["param_names"]]
    return {n: g["params"][:, i] for i, n in enumerate(names)}


def _product_maze(g):
    from navigation_model_b200.simulation.maze import GridTiles, Maze
    return Maze(str(g["layout"]), "", size_tile=float(g["size_tile"]), tile_system=GridTiles)


def _classes():
    from navigation_model_b200.simulation import behavioural_components as bc
    return [bc.AnxietyComponent_Grid, bc.BiomechanicalCostComponent, bc.ExperienceComponent,
            bc.BiomechPersistenceComponent, bc.ValueTableComponent]


def _instances(g, maze, m):
    from navigation_model_b200.simulation import behavioural_components as bc
    P = _params(g)
    p = {k: float(v[m]) for k, v in P.items()}
    units = [tuple(int(v) for v in u) for u in g["units"]]
    return p, [bc.AnxietyComponent_Grid(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"], p4=p["p4"]),
               bc.BiomechanicalCostComponent(units, W_bcost=p["W_bcost"], k=p["k"], gamma=p["gamma"]),
               bc.ExperienceComponent(maze.shape, W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                      kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"],
                                      theta_home=p["theta_home"]),
               bc.BiomechPersistenceComponent(units, W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"]),
               bc.ValueTableComponent(g["v_table"], W_values=p["W_values"])]


def test_oracle_matches_reference(g):
    st = float(g["size_tile"])
    mz = O.OracleMaze(str(g["layout"]), st, grid_tiles=True)
    assert mz.tiles[1, 1] == 4 and mz.tiles[4, 4] == 4
    P = _params(g)
    units = g["units"].astype(np.float64)
    for m in range(len(g["pose"])):
        comps = [(O.COMP_ANXIETY, {k: P[k][m] for k in ("W_anx", "p1", "p2", "p3", "p4")}),
                 (O.COMP_BCOST, {"W_bcost": P["W_bcost"][m], "table": O.bcost_table(units, P["k"][m], P["gamma"][m])}),
                 (O.COMP_EXPERIENCE, {k: P[k][m] for k in ("W_exp", "xi", "kappa_home", "kappa_explo", "theta_explo",
                                                           "theta_home")}),
                 (O.COMP_BPERSISTENCE, {k: P[k][m] for k in ("W_bper", "bp", "Wm", "Wnm")}),
                 (O.COMP_VTABLE, {"W_values": P["W_values"][m], "table": O.normalise_vtable(g["v_table"])})]
        res = O.simulate(mz, units * st, units, g["pose"][m, :2], g["pose"][m, 2], g["init_disc"][m], P["beta"][m], comps,
                         g["draws"][m], record_probs=(m == 0))
        assert np.array_equal(res["hist_pos"], g["hist_pos"][m])
        assert np.array_equal(res["hist_ori"], g["hist_ori"][m])
        if m == 0:
            for t, pv in enumerate(res["probs"]):
                np.testing.assert_allclose(pv, g["m0_probs"][t, :len(pv)], rtol=4e-16, atol=0)


def test_host_plugin_loop_matches_reference(g):
    """The scalar drop-in classes on a GridTiles maze (a user-defined exploration-model subclass takes the per-step
    host loop): bit-exact history."""
    from navigation_model_b200.simulation.experiment import StandardExperiment
    from navigation_model_b200.simulation.exploration_model import BehaviouralModel
    from navigation_model_b200.simulation.mouse import Mouse
    maze = _product_maze(g)
    user_model = type("UserModel", (BehaviouralModel,), {})
    for m in (0, 3):
        p, comps = _instances(g, maze, m)
        model = user_model(comps, tuple(int(v) for v in g["init_disc"][m]), beta=p["beta"], seed=500 + m)
        mouse = Mouse(g["pose"][m, :2], g["pose"][m, 2], g["units"] * float(g["size_tile"]))
        exp = StandardExperiment(model, maze, mouse)
        assert not exp._device_eligible(300)
        exp.run(300)
        hp, ho = mouse.history
        assert np.array_equal(np.array(hp), g["hist_pos"][m])
        assert np.array_equal(np.array(ho), g["hist_ori"][m])


@pytest.mark.gpu
def test_batched_kernel_matches_reference(cuda, g):
    from navigation_model_b200.batched_sim import BatchedExperiment
    maze = _product_maze(g)
    st = float(g["size_tile"])
    P = _params(g)
    T = g["draws"].shape[1]
    exp = BatchedExperiment(maze, g["units"] * st, g["units"], _classes(), v_table=g["v_table"])
    res = exp.run(P, T, g["pose"][:, :2], g["pose"][:, 2], init_disc=g["init_disc"], draws=g["draws"], history=True,
                  occupancy=True)
    assert not res.status.any()
    for m in range(len(g["pose"])):
        assert np.array_equal(np.floor_divide(res.history_position[m], st), np.floor_divide(g["hist_pos"][m], st))
        assert np.array_equal(res.history_position[m], g["hist_pos"][m])
        assert np.array_equal(res.history_orientation[m], g["hist_ori"][m])
    assert np.array_equal(res.occupancy, g["occupancy"].astype(np.int64))
    assert np.array_equal(res.tile_counts[:, :5], g["tile_counts"])  # VOID, WALL, CORNER, CENTRE, DECISION_POINT


@pytest.mark.gpu
def test_standard_experiment_device_path(cuda, g):
    """StandardExperiment.run with the built-in classes on the GridTiles maze is one launch (seeded from the model's
    generator) and leaves the mouse where the reference's loop does."""
    from navigation_model_b200.simulation.experiment import StandardExperiment
    from navigation_model_b200.simulation.exploration_model import BehaviouralModel
    from navigation_model_b200.simulation.mouse import Mouse
    maze = _product_maze(g)
    st = float(g["size_tile"])
    for m in (1, 2):
        p, comps = _instances(g, maze, m)
        model = BehaviouralModel(comps, tuple(int(v) for v in g["init_disc"][m]), beta=p["beta"], seed=500 + m)
        mouse = Mouse(g["pose"][m, :2], g["pose"][m, 2], g["units"] * st)
        exp = StandardExperiment(model, maze, mouse)
        assert exp._device_eligible(300)
        exp.run(300)
        hp = np.array(mouse.history[0])
        assert np.array_equal(np.floor_divide(hp, st), np.floor_divide(g["hist_pos"][m], st))
        assert np.array_equal(hp, g["hist_pos"][m])
        assert np.array_equal(np.array(mouse.history[1]), g["hist_ori"][m])
