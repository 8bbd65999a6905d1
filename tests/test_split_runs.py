"""``StandardExperiment.run`` called several times on the same objects (``experiment.py:33-51``): the model, the
components and the mouse carry on between the calls, and ``enable_history(True)`` (``experiment.py:32`` ->
``mouse.cpp:38-41, 55-61``) restarts the history at the mouse's INITIAL pose.

Golden: ``tests/golden/split_runs_golden.npz`` -- the unmodified reference running ``run(1); run(249); run(350)`` on
three mice of ``sim_golden.npz`` (``tests/golden/make_golden_split_runs.py``).  The host plugin loop is held to it on
the CPU; on the GPU every leg is one ``nm_simulate`` launch seeded with the state the previous leg left in the Python
objects (visit map, familiarity, mental state, ``_moving``): discrete results AND poses exact (csrc/nm_libm.h).
"""
import os

import numpy as np
import pytest

import helpers as H
from conftest import GOLDEN

POSE_ATOL = 0.0


@pytest.fixture(scope="module")
def golden_split():
    return dict(np.load(os.path.join(GOLDEN, "split_runs_golden.npz"), allow_pickle=False))


def _experiment(g, m):
    from navigation_model_b200.simulation import behavioural_components as bc
    from navigation_model_b200.simulation.experiment import StandardExperiment
    from navigation_model_b200.simulation.exploration_model import BehaviouralModel
    from navigation_model_b200.simulation.mouse import Mouse
    maze = H.product_m9()
    p = dict(zip([str(s) for s in g["param_names"]], g["params"][m]))
    units = [tuple(int(v) for v in a) for a in g["actions_units"]]
    comps = [bc.AnxietyComponent(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"]),
             bc.BiomechanicalCostComponent(units, W_bcost=p["W_bcost"], k=p["k"], gamma=p["gamma"]),
             bc.ExperienceComponent(maze.shape, W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                    kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"],
                                    theta_home=p["theta_home"]),
             bc.BiomechPersistenceComponent(units, W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"]),
             bc.ValueTableComponent(g["v_table"], W_values=p["W_values"])]
    model = BehaviouralModel(comps, tuple(int(v) for v in g["init_disc"][m]), beta=p["beta"], seed=int(g["seeds"][m]))
    mouse = Mouse(g["init_pose"][m][:2], g["init_pose"][m][2], g["actions"])
    return StandardExperiment(model, maze, mouse), model, mouse, comps[2]


def _check(gs, i, model, mouse, experience, atol):
    hp, ho = mouse.history
    hp, ho = np.array(hp), np.array(ho)
    assert hp.shape == gs["hist_pos"][i].shape
    np.testing.assert_allclose(hp, gs["hist_pos"][i], rtol=0, atol=atol)
    d = np.abs(ho - gs["hist_ori"][i])
    assert np.all(np.minimum(d, np.abs(d - 2 * np.pi)) <= max(atol, 1e-15) * 8)
    assert np.array_equal(experience._maze_map_exp, gs["visits"][i])  # the visit map, exactly
    moving, fam, ment, px, py = gs["state"][i]
    assert bool(model._moving) == bool(moving) and int(experience._ment_state) == int(ment)
    np.testing.assert_allclose(experience._familiarity, fam, rtol=1e-13)
    np.testing.assert_allclose(np.asarray(model._prev_pos, dtype=np.float64).reshape(2), [px, py], rtol=0, atol=atol)
    assert model._rng.random() == gs["next_draw"][i]  # the generator advanced by exactly one draw per step


def test_host_loop_split_runs_match_reference(golden_sim, golden_split):
    """The per-step plugin loop (host) against the reference's fixture: bit-exact."""
    for i, m in enumerate(golden_split["mice"]):
        exp, model, mouse, experience = _experiment(golden_sim, int(m))
        for leg in golden_split["legs"]:
            exp._mouse.enable_history(True)
            exp._run_host(int(leg))
        _check(golden_split, i, model, mouse, experience, atol=0.0)


@pytest.mark.gpu
def test_device_split_runs_match_reference(cuda, golden_sim, golden_split):
    """Every leg is one kernel launch that starts from the state the previous leg wrote back."""
    for i, m in enumerate(golden_split["mice"]):
        exp, model, mouse, experience = _experiment(golden_sim, int(m))
        for leg in golden_split["legs"]:
            assert exp._device_eligible(int(leg))
            exp.run(int(leg))
        _check(golden_split, i, model, mouse, experience, atol=POSE_ATOL)


@pytest.mark.gpu
def test_device_leg_then_host_leg_and_back(cuda, golden_sim, golden_split):
    """Device and host legs interleave on the same objects: device -> host loop -> device."""
    legs = [int(v) for v in golden_split["legs"]]
    for i, m in enumerate(golden_split["mice"]):
        exp, model, mouse, experience = _experiment(golden_sim, int(m))
        exp.run(legs[0])
        mouse.enable_history(True)
        exp._run_host(legs[1])
        exp.run(legs[2])
        _check(golden_split, i, model, mouse, experience, atol=POSE_ATOL)


@pytest.mark.gpu
@pytest.mark.parametrize("counters", ["16", "32"])
def test_batched_seeded_visit_maps(cuda, golden_sim, counters, monkeypatch):
    """BatchedExperiment.run(init_visits=, init_model_state=): two seeded launches reproduce one long launch for a
    batch of mice (choices, occupancy and final state), with both widths of the visit counters; a seed that does not
    fit 16 bits takes the wide counters by itself."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    monkeypatch.setenv("NM_SIM_COUNTERS", counters)
    g = golden_sim
    maze = H.product_m9()
    names = [str(s) for s in g["param_names"]]
    N, T1, T2 = len(g["seeds"]), 220, 380
    params = {k: g["params"][:, j] for j, k in enumerate(names)}
    classes = [bc.AnxietyComponent, bc.BiomechanicalCostComponent, bc.ExperienceComponent,
               bc.BiomechPersistenceComponent, bc.ValueTableComponent]
    exp = BatchedExperiment(maze, g["actions"], g["actions_units"], classes, v_table=g["v_table"])
    draws = g["draws"]
    kw = dict(init_disc=g["init_disc"], history=True, occupancy=True)
    whole = exp.run(params, T1 + T2, g["init_pose"][:, :2], g["init_pose"][:, 2], draws=draws, **kw)
    a = exp.run(params, T1, g["init_pose"][:, :2], g["init_pose"][:, 2], draws=draws[:, :T1], **kw)
    seed = a.occupancy.copy()  # visits experienced so far: the final position is still to come
    last = maze.cont2disc_array(a.final_pose[:, :2])
    seed[np.arange(N), last[:, 0], last[:, 1]] -= 1
    b = exp.run(params, T2, a.final_pose[:, :2], a.final_pose[:, 2], draws=draws[:, T1:], history=True, occupancy=True,
                init_disc=a.model_state[:, :2], init_visits=seed, init_model_state=a.model_state[:, 2:5])
    assert np.array_equal(np.concatenate([a.choices, b.choices], axis=1), whole.choices)
    np.testing.assert_allclose(b.history_position, whole.history_position[:, T1:], rtol=0, atol=POSE_ATOL)
    assert np.array_equal(b.occupancy, whole.occupancy)
    np.testing.assert_allclose(b.model_state, whole.model_state, rtol=1e-13, atol=0)
    assert np.array_equal(whole.occupancy.sum(axis=(1, 2)), np.full(N, T1 + T2 + 1))
    # a seed beyond 16 bits: same trajectory as long as the ExperienceComponent saturates the same way -- here only the
    # bookkeeping is checked: the seeded visits come back on top of the run's own
    big = np.zeros(maze.shape, dtype=np.int64)
    big[1, 1] = 70000
    c = exp.run(params, 50, g["init_pose"][:, :2], g["init_pose"][:, 2], draws=draws[:, :50], occupancy=True,
                init_disc=g["init_disc"], init_visits=big)
    assert np.array_equal(c.occupancy.astype(np.int64).sum(axis=(1, 2)), np.full(N, 70000 + 51))
    with pytest.raises(ValueError):
        exp.run(params, 5, g["init_pose"][:, :2], g["init_pose"][:, 2], draws=draws[:, :5], init_visits=-np.ones(maze.shape))
