"""Irregular random mazes: outputs of the UNMODIFIED reference (tests/golden/random_mazes_golden.npz, written by
tests/golden/make_golden_random.py) against the oracle on the CPU and against the CUDA path on the GPU.  Pins the
ray / AABB walks, tile lookups and nearest-visitable searches on geometry the 9 x 9 test maze does not have: void tiles
next to corridors, dead ends, action subsets."""
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "random_mazes_golden.npz")
ALGS = {"td0": O.ALG_TD0, "positive": O.ALG_POSITIVE, "negative": O.ALG_NEGATIVE}


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN, allow_pickle=False))


def _case(g, seed):
    k = f"s{seed}_"
    st = float(g["size_tile"])
    layout = str(g[k + "layout"])
    return k, st, layout, O.OracleMaze(layout, st), g[k + "units"]


def _components(g, k, m, units):
    p = dict(zip([str(s) for s in g["param_names"]], g[k + "sim_params"][m]))
    comps = [(O.COMP_ANXIETY, dict(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"])),
             (O.COMP_BCOST, dict(W_bcost=p["W_bcost"], table=O.bcost_table(units, p["k"], p["gamma"]))),
             (O.COMP_EXPERIENCE, dict(W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                      kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"],
                                      theta_home=p["theta_home"])),
             (O.COMP_BPERSISTENCE, dict(W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"])),
             (O.COMP_VTABLE, dict(W_values=p["W_values"], table=O.normalise_vtable(g[k + "v_table"])))]
    return p, comps


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_oracle_value_learning_matches_reference(g, seed):
    k, st, layout, mz, units = _case(g, seed)
    traj, rew = g[k + "traj"], g[k + "rew"]
    tr = O.extract_transitions(mz, traj, rew, O.OracleMouse(traj[0], 0.0, units * st))
    cells = np.array(mz.vis_cells)
    assert np.array_equal(cells[tr["s"]], g[k + "tr_s"]) and np.array_equal(cells[tr["s1"]], g[k + "tr_s1"])
    assert np.array_equal(tr["ncand"], g[k + "tr_ncand"])
    for i in range(len(tr["s"])):
        n = tr["ncand"][i]
        assert np.array_equal(cells[tr["cand"][i, :n]], g[k + "tr_cand"][i, :n])
    orders = O.replay_orders(len(tr["s"]), 1, seed)
    for name, code in ALGS.items():
        t = dict(tr, r=-tr["r"]) if name == "negative" else tr
        for p, (alpha, gamma) in enumerate(g[k + "params"]):
            V, _ = O.run_session(code, t, alpha, gamma, len(mz.vis_cells), orders=orders)
            assert np.array_equal(O.dense_table(mz, V), g[k + name + "_V"][p]), (name, p)
        Vc, _ = CO.fit(code, t, g[k + "params"][:, 0], g[k + "params"][:, 1], len(mz.vis_cells), orders=orders)
        assert np.array_equal(O.dense_table(mz, Vc[1]), g[k + name + "_V"][1])


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_oracle_simulation_matches_reference(g, seed):
    k, st, layout, mz, units = _case(g, seed)
    for m in range(len(g[k + "sim_params"])):
        p, comps = _components(g, k, m, units)
        pose = g[k + "sim_pose"][m]
        res = O.simulate(mz, units * st, units, pose[:2], pose[2], g[k + "sim_disc"][m], p["beta"], comps,
                         g[k + "sim_draws"][m])
        assert res["status"] == 0
        assert np.array_equal(res["hist_pos"], g[k + "sim_hist_pos"][m]), (seed, m)  # bit-exact continuous trajectory
        assert np.array_equal(res["hist_ori"], g[k + "sim_hist_ori"][m]), (seed, m)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_cuda_paths_match_reference(cuda, g, seed):
    """The same reference outputs through the product: V-tables bit-exact (sweep of the two parameter sets with the
    replay pass), simulated histories bit for bit (positions and headings)."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.sweep import ConditioningSweep
    k, st, layout, mz, units = _case(g, seed)
    maze = Maze(layout, "", size_tile=st)
    traj, rew = g[k + "traj"], g[k + "rew"]
    for name in ALGS:
        sw = ConditioningSweep(maze, name, possible_actions=units * st, replay_n_times=1, replay_seed=seed)
        sw.add_session(traj, -rew if name == "negative" else rew)
        V = sw.v_tables(g[k + "params"][:, 0], g[k + "params"][:, 1])
        assert np.array_equal(V[0], g[k + name + "_V"]), name
    names = [str(s) for s in g["param_names"]]
    P = {n: g[k + "sim_params"][:, i] for i, n in enumerate(names)}
    exp = BatchedExperiment(maze, units * st, units, [bc.AnxietyComponent, bc.BiomechanicalCostComponent,
                                                      bc.ExperienceComponent, bc.BiomechPersistenceComponent,
                                                      bc.ValueTableComponent], v_table=g[k + "v_table"])
    pose = g[k + "sim_pose"]
    res = exp.run(P, g[k + "sim_draws"].shape[1], pose[:, :2], pose[:, 2], init_disc=g[k + "sim_disc"],
                  draws=g[k + "sim_draws"], history=True)
    want = g[k + "sim_hist_pos"]
    for m in range(len(pose)):
        assert res.status[m] == 0
        got_tiles = np.floor_divide(res.history_position[m], st)
        assert np.array_equal(got_tiles, np.floor_divide(want[m], st)), (seed, m)
        assert np.array_equal(res.history_position[m], want[m])
        assert np.array_equal(res.history_orientation[m], g[k + "sim_hist_ori"][m])
