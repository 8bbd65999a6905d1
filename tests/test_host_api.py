"""Host-side mirror of the reference interface, checked with the known answers the reference's own
tests hold (test/test_rl.py, test_components.py, test_mouse.py, test_maze.py, test_metrics.py,
test_exploration.py, test_session.py, test_fitness.py) -- re-expressed here so that they also run where
/root/reference does not exist.  No GPU needed: these are the scalar plugin interfaces."""
import copy

import numpy as np
import numpy.testing as npt
import pytest
from scipy.stats import vonmises

from oracle import np_oracle as O

import helpers as H
from navigation_model_b200.analysis.data import Session, SessionList, adjust_positions_to_maze
from navigation_model_b200.analysis.metrics import (Analyzer, Metric, discretize_positions, it_perc_visited_maze,
                                                    metric, motion_analysis, occupancy_map, tile_analysis)
from navigation_model_b200.analysis.statistics import manhattan_distance, normalized_distance
from navigation_model_b200.optimization.fitness import MultiObjectiveFitness
from navigation_model_b200.simulation import behavioural_components as bc
from navigation_model_b200.simulation.exploration_model import BehaviouralModel, RandomModel
from navigation_model_b200.simulation.maze import GridTiles, Maze, StandardTiles
from navigation_model_b200.simulation.mouse import Mouse
from navigation_model_b200.simulation.rl import (NegativeConditioning, PositiveConditioning, ShuffledReplays, TD0,
                                                 Transition, Update)

M7 = "CWWWWWC\nWMMMMMW\nWMWWWMW\nWMWVWMW\nWMWWWMW\nWMMMMMW\nCWWWWWC"
SUB7 = "1111111\n1111111\n2222222\n2220222\n3333333\n3333333"


# ---------------------------------------------------------------------------------------------- rl
class _StillMouse:
    def move_to(self, _):
        pass


@pytest.mark.parametrize("cls,pick,r", [(PositiveConditioning, np.max, 1), (NegativeConditioning, np.min, -1)])
def test_lookahead_update_formula(cls, pick, r):
    rng = np.random.default_rng(0)
    alpha, gamma = rng.random(), rng.random()
    v = rng.random((3, 3))
    alg = cls(alpha=alpha, discount_rate=gamma, mouse=_StillMouse(), maze=None)
    alg.v_table = copy.copy(v)
    nb = [(1, 0), (0, 1), (2, 1), (1, 2)]
    upd = alg.update(Transition(s=(1, 1), r=r, extra={"next_states": nb}))
    dv = alpha * (r + gamma * pick([v[c] for c in nb]) - v[1, 1])
    assert upd.dv == dv and alg.v_table[1, 1] == v[1, 1] + dv and upd.rpe == abs(dv / alpha) or alpha == 0


def test_td0_update_formula_and_uninitialised_table():
    alg = TD0(alpha=0.3, discount_rate=0.7)
    with pytest.raises(RuntimeError):
        _ = alg.v_table
    v = np.random.default_rng(1).random((3, 3))
    alg.v_table = copy.copy(v)
    upd = alg.update(Transition(s=(1, 1), s1=(2, 0), r=1))
    dv = 0.3 * (1 + 0.7 * v[2, 0] - v[1, 1])
    assert upd.dv == dv and alg.v_table[1, 1] == v[1, 1] + dv


def test_shuffled_replays_present_every_transition_n_times():
    class Count:
        def __init__(self):
            self.counts = []

        def update(self, t):
            if "c" in t.extra:
                self.counts[t.extra["c"]] += 1
            else:
                t.extra["c"] = len(self.counts)
                self.counts.append(0)
            return Update(transition=t)

    alg, rep = Count(), ShuffledReplays(n_times=10)
    for _ in range(200):
        rep.append(alg.update(Transition()))
    rep.offline_replays(alg)
    assert sum(alg.counts) == 2000 and max(alg.counts) == 10 and min(alg.counts) == 10


# ------------------------------------------------------------------------------------------- mouse
def test_mouse_pose_endpoints_history():
    m = Mouse((1, 2), np.pi / 2, [(1, 0), (-1, 0)])
    npt.assert_allclose(m.position, (1, 2))
    assert m.orientation == np.pi / 2
    ends = m.get_endpoints()
    assert np.allclose(ends[0], (1, 3)) and np.allclose(ends[1], (1, 1))
    assert m.move_to((2, 2)) and m.orientation == 0
    assert m.move_to(0, 2) and abs(m.orientation - np.pi) < 1e-15
    assert not m.move_to(0, 2)  # isApprox: no move, orientation kept, history still pushed
    assert len(m.history_position) == 4
    m.reset()
    npt.assert_allclose(m.position, (1, 2))
    assert m.history_orientation == [np.pi / 2]
    m.enable_history(False)
    m.move_to(2, 2)
    assert m.history_position == [] and m.history_orientation == []
    m.reset(initial_pos=(5, 5), initial_ori=1.0)
    npt.assert_allclose(m.initial_position, (5, 5))
    assert m.initial_orientation == 1.0
    with pytest.raises(ValueError):
        Mouse((0, 0), 0.0, [(1, 0, 0)])


# -------------------------------------------------------------------------------------------- maze
def test_maze_known_answers():
    maze = Maze(M7, SUB7, size_tile=0.111)
    assert maze.shape == (7, 7) and maze.size_tile == 0.111
    d = maze.description()
    assert (d[maze.MazeTile.CORNER], d[maze.MazeTile.WALL], d[maze.MazeTile.CENTRE]) == (4, 28, 16)
    assert maze.cont2disc((0.443, 0.221)) == (3, 1)
    npt.assert_allclose(maze.get_tile_centre((3, 1)), (0.3885, 0.1665))
    npt.assert_allclose(maze.disc2cont((3, 5)), (0.3885, 0.6105))
    assert maze.get_tile(0, 0) == maze.MazeTile.CORNER
    assert maze.get_tiles([(6, 6), (3, 3)]) == [maze.MazeTile.CORNER, maze.MazeTile.VOID]
    assert maze.get_subareas([(0, 0), (3, 3)]) == [1, 0]
    assert maze.get_coordinates_of_type(maze.MazeTile.CORNER) == [(0, 0), (0, 6), (6, 0), (6, 6)]
    assert maze.get_number_visitable_tiles() == 48
    assert maze.get_closest_visitable(-1, -1) == (0, 0)
    assert maze.is_movement_possible(0, 0, 0, 6) and not maze.is_movement_possible(3, 0, 3, 6)
    assert maze.are_visitable([(1, 1), (3, 3)]) == [True, False]
    nosub = Maze(M7, "", size_tile=0.111)
    with pytest.raises(RuntimeError):
        nosub.get_subareas((0, 0))
    assert GridTiles.char_to_mazetile("D") == GridTiles.TileType.DECISION_POINT
    assert StandardTiles.char_to_mazetile("D") == StandardTiles.TileType.VOID


def test_maze_geometry_matches_oracle_and_native_handle():
    """Python Maze, the native host geometry (nm_maze_*) and the oracle agree on boundary cases."""
    import ctypes as C
    from navigation_model_b200 import _native
    from navigation_model_b200.sweep import host_maze
    maze, omz = H.product_m9(), H.oracle_m9()
    hm = host_maze(maze)
    lib = _native.lib()
    rng = np.random.default_rng(3)
    st = O.M9_SIZE_TILE
    pts = np.concatenate([rng.uniform(-0.2, 1.2, (500, 2)), rng.integers(-1, 11, (300, 2)) * st])
    for x, y in pts:
        dx, dy = C.c_int(), C.c_int()
        lib.nm_maze_cont2disc(hm.handle, x, y, C.byref(dx), C.byref(dy))
        assert (dx.value, dy.value) == maze.cont2disc(x, y) == omz.cont2disc(x, y)
        lib.nm_maze_closest_visitable_cont(hm.handle, x, y, C.byref(dx), C.byref(dy))
        assert (dx.value, dy.value) == maze.get_closest_visitable_cont(x, y) == omz.closest_visitable(x, y)
    segs = np.concatenate([rng.uniform(0, 1, (600, 4)), rng.integers(0, 10, (300, 4)) * st])
    for s in segs:
        want = omz.move_possible(*s)
        assert maze.is_movement_possible_cont(*s) == want
        assert bool(lib.nm_maze_movement_possible_cont(hm.handle, *s)) == want
    assert np.array_equal(maze.cont2disc_array(pts), [maze.cont2disc(x, y) for x, y in pts])
    tt = maze.transition_table([(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1)])
    assert tt.shape == (60, 5) and (tt[:, 0] == np.arange(60)).all()


# -------------------------------------------------------------------------------------- components
def test_component_known_answers():
    anx = bc.AnxietyComponent(W_anx=0.8, p1=1.0, p2=0.6, p3=0.5)
    assert anx.parameters == {"W_anx": 0.8, "p1": 1.0, "p2": 0.6, "p3": 0.5} and anx.name == "anx"
    T = StandardTiles.TileType
    anx.update([T.CORNER, T.CENTRE, T.WALL, T.VOID, T.CENTRE])
    npt.assert_allclose(anx.values, [1, 0.3, 0.6, -1, 0.3])
    npt.assert_allclose(anx.weighted_values, 0.8 * anx.values)
    anx.active = False
    assert anx.values == 0 and anx.weighted_values == 0
    with pytest.raises(RuntimeError):
        bc.AnxietyComponent(W_anx=0.8, p1=1.0, p2=1.6, p3=0.5)  # p2 out of [0, 1]
    with pytest.raises(RuntimeError):
        anx.W_anx = 11.0

    acts = [[0, 0], [1, 0], [2, 0], [0, -1]]
    bcost = bc.BiomechanicalCostComponent(acts, W_bcost=0.7, k=0.5, gamma=2.0)
    v = vonmises(0.5).pdf(0.0)
    bcost.update(True, [True, True, True, False])
    npt.assert_allclose(bcost.values, [0.0, v, 2.0 * v])
    bcost.update(False, [True, True, True, False])
    npt.assert_allclose(bcost.values, np.zeros(3))
    assert np.array_equal(bc.bcost_tables(acts, [0.5, 1.5], [2.0, 0.3])[0], bcost._cachedval)
    assert np.array_equal(bc.bcost_tables(O.A8_UNITS, 1.7, 0.4)[0], O.bcost_table(O.A8_UNITS, 1.7, 0.4))
    # the vectorised table is scipy's vonmises.pdf bit for bit (the reference calls vonmises(k).pdf, :395-399)
    ks = np.concatenate([np.random.default_rng(3).uniform(0, 60, 500), [0.0, 1e-12, 700.0]])
    ang = np.arctan2(O.A8_UNITS[:, 1], O.A8_UNITS[:, 0])
    want = vonmises.pdf(ang[None, :], ks[:, None])
    want[:, 0] = 0.0            # the null action
    want[:, 2] *= 0.25          # (2, 0) is two tiles long: times gamma
    assert np.array_equal(bc.bcost_tables(O.A8_UNITS, ks, np.full(len(ks), 0.25)), want)

    exp = bc.ExperienceComponent((7, 7), W_exp=0.7, theta_home=30, theta_explo=80, kappa_home=5.0, kappa_explo=2.0,
                                 xi=0.006)
    exp.update(4, 2, [(3, 2), (4, 3), (4, 2)])
    assert exp._maze_map_exp[4, 2] == 1 and exp._familiarity == 0.006
    assert exp._ment_state == bc.ExperienceComponent.State.HOME
    npt.assert_allclose(exp.values, [0.0, 0.0, 0.9932620530009145])

    bper = bc.BiomechPersistenceComponent(acts, W_bper=0.7, bp=2.0, Wm=5.0, Wnm=3.0)
    bper.update(True, [True, True, True, False])
    npt.assert_allclose(bper.values, [10.0, 2.0, 2.0])
    bper.update(False, [True, True, True, False])
    npt.assert_allclose(bper.values, [2.0, 6.0, 6.0])

    vt = np.random.default_rng(2).random((5, 5))
    comp = bc.ValueTableComponent(vt, W_values=0.7)
    comp.update([(3, 2), (4, 3)])
    want = (vt - vt.min()) / (vt.max() - vt.min()) - 1
    npt.assert_allclose(comp.values, [want[3, 2], want[4, 3]])
    assert not bc.ValueTableComponent(np.zeros((3, 3)), W_values=1.0)._v_table.any()
    assert [d.name for d in bc.ExperienceComponent.get_descriptors()] == \
        ["theta_home", "theta_explo", "kappa_explo", "kappa_home", "xi", "W_exp"]


# -------------------------------------------------------------------------------- exploration models
@bc.weight("W_choice")
class _Choice(bc.Component):
    _name = "choice"

    def __init__(self, n, **kw):
        super().__init__(**kw)
        self._n = n

    def update(self, chosen_next, **kw):
        self._val = np.zeros(self._n) - np.inf
        self._val[chosen_next] = 1


@bc.weight("W_fixed")
class _Fixed(bc.Component):
    _name = "fixed"

    def __init__(self, vals, **kw):
        super().__init__(**kw)
        self._val = vals

    def update(self, **kw):
        pass


def test_behavioural_model_softmax_choice_and_seed():
    rng = np.random.default_rng(4)
    positions = rng.integers(100, size=10)
    model = BehaviouralModel([_Choice(n=10)], (0, 0), 1.0)
    for _ in range(50):
        c = int(rng.integers(10))
        assert model.decision_making(positions, chosen_next=c) == positions[c]
    vals = rng.random(10)
    cont = rng.random(10)
    a = BehaviouralModel([_Fixed(vals=vals)], (0, 0), 1.0, seed=12345)
    b = BehaviouralModel([_Fixed(vals=vals)], (0, 0), 1.0, seed=12345)
    twin = np.random.default_rng(12345)
    p = np.exp(vals - vals.max()) / np.sum(np.exp(vals - vals.max()))
    for _ in range(50):
        pa = a.decision_making(next_positions_cont=cont)
        assert pa == b.decision_making(next_positions_cont=cont) == twin.choice(cont, p=p)
    assert a.parameters == {"beta": 1.0, "W_fixed": 1.0}
    r1, r2 = RandomModel(seed=7), RandomModel(seed=7)
    for _ in range(20):
        assert r1.decision_making(next_positions_cont=cont) == r2.decision_making(next_positions_cont=cont)


def test_host_plugin_experiment_loop_matches_oracle():
    """A model with a user-defined component cannot run on the device: the generic per-step loop is used
    and agrees with the oracle when the extra component adds zeros."""
    from navigation_model_b200.simulation.experiment import StandardExperiment
    maze, omz = H.product_m9(), H.oracle_m9()
    acts = H.a8()
    comps = [bc.AnxietyComponent(W_anx=1.3, p1=0.9, p2=0.5, p3=0.4),
             bc.BiomechPersistenceComponent(O.A8_UNITS, W_bper=0.8, bp=0.7, Wm=1.5, Wnm=0.6)]

    @bc.weight("W_zero")
    class Zero(bc.Component):
        _name = "zero"

        def update(self, next_positions, **kw):
            self._val = np.zeros(len(next_positions))

    model = BehaviouralModel(comps + [Zero()], (1, 1), beta=2.0, seed=3)
    mouse = Mouse(maze.get_tile_centre(1, 1), 0.0, acts)
    exp = StandardExperiment(model, maze, mouse)
    assert not exp._device_eligible()
    exp.run(60)
    draws = np.random.default_rng(3).random(60)
    oc = [(O.COMP_ANXIETY, dict(W_anx=1.3, p1=0.9, p2=0.5, p3=0.4)),
          (O.COMP_BPERSISTENCE, dict(W_bper=0.8, bp=0.7, Wm=1.5, Wnm=0.6))]
    ref = O.simulate(omz, acts, O.A8_UNITS, maze.get_tile_centre(1, 1), 0.0, (1, 1), 2.0, oc, draws)
    assert np.array_equal(np.array(mouse.history_position), ref["hist_pos"])


# ----------------------------------------------------------------------------------------- metrics
def test_metric_functions_known_answers():
    maze = H.product_m9()
    disc = discretize_positions([(0.1, 0.1), (0.35, 0.47)], maze)
    assert disc == [(0, 0), (3, 4)]
    occ = occupancy_map([(0, 0), (3, 4), (0, 0)], maze)
    assert occ.sum() == 3 and occ[0, 0] == 2 and occ[3, 4] == 1
    T = maze.MazeTile
    perc, ratio, count = tile_analysis([(0, 0), (1, 1), (0, 0)], maze, [T.CORNER, T.CENTRE, T.WALL])
    npt.assert_allclose([perc[T.CORNER], perc[T.CENTRE], perc[T.WALL]], [200 / 3, 100 / 3, 0])
    totals = maze.description()
    npt.assert_allclose(ratio[T.CORNER], 2 / totals[T.CORNER])
    assert count[T.CENTRE] == 1
    assert motion_analysis([(0, 0), (1, 1), (0, 0), (0, 0)]) == (2, [False, True, True, False])
    cells = maze.get_visitable_coordinates()
    assert it_perc_visited_maze(cells, maze, percentage=80) == 47      # 48 of 60 tiles at index 47
    assert it_perc_visited_maze(cells[:10], maze) == 10                 # never reached -> len
    omz = H.oracle_m9()
    hist = np.array([omz.centre(c) for c in [(1, 1), (1, 2), (1, 2), (1, 3), (2, 8)]])
    met = O.metrics_of_history(omz, hist, 5)
    d = discretize_positions(hist, maze)
    assert np.array_equal(occupancy_map(d, maze), met["occupancy"])
    assert motion_analysis(d)[0] == met["count_moving"] and it_perc_visited_maze(d, maze, 5) == met["it_visit"]


def test_metric_pipeline():
    @metric("c")
    def f1(a, b):
        return a + b

    @metric("d")
    def f2(b, c):
        return b * c

    @metric("f")
    def f3(d, e=3):
        return d ** e

    m = Metric(f1, f1.requirements, f1.products)
    assert m.requirements == ["a", "b"] and m.products is None and m(a=1, b=2) == 3 and m.products == {"c": 3}
    assert Analyzer(f1, f2, f3).analize(a=1, b=2) == {"a": 1, "b": 2, "c": 3, "d": 6, "f": 216}
    with pytest.raises(RuntimeError):
        Analyzer(f2).analize(b=2)
    res = Analyzer(discretize_positions, occupancy_map, motion_analysis)(
        continuous_positions=[(0.1, 0.1), (0.35, 0.47)], maze=H.product_m9())
    assert res["count_moving"] == 1 and res["occupancy"].sum() == 2


# ------------------------------------------------------------------------------- sessions / fitness
def test_session_and_session_list():
    rng = np.random.default_rng(6)
    times = np.linspace(0.0, 1.0, 10)
    traj, reward = rng.random((10, 2)), rng.random(10)
    s = Session(times, traj, reward=reward)
    npt.assert_allclose(s.sampling_time, 1 / 9)
    npt.assert_allclose(s.initial_orientation, np.arctan2(traj[1][1] - traj[0][1], traj[1][0] - traj[0][0]))
    assert len(s) == 10 and s.has_reward
    sub = Session(times, traj, reward=reward, new_sampling_time=0.2)
    npt.assert_allclose(sub.trajectory, traj[::2])
    npt.assert_allclose(sub.reward, reward[::2])
    # NaN timestamps / positions are skipped, the first non-zero reward of a window is kept
    t2 = times.copy()
    t2[3] = np.nan
    tr2 = traj.copy()
    tr2[4] = np.nan
    rw2 = np.zeros(10)
    rw2[5] = 2.0
    sub2 = Session(t2, tr2, reward=rw2, new_sampling_time=0.2)
    assert np.isfinite(np.array(sub2.trajectory)).all() and 2.0 in sub2.reward
    with pytest.raises(RuntimeError):
        Session(times[:-1], traj)
    with pytest.raises(RuntimeError):
        Session(times, traj, reward=reward[:-1])
    d = s.to_dict()
    s2 = Session.from_dict(d)
    npt.assert_allclose(s2.trajectory, s.trajectory)
    sl = SessionList(s, sub)
    assert len(sl) == 2 and len((sl + sl)) == 4 and len(3 * sl) == 6 and len(sl[0:1]) == 1
    assert len(sl.all_trajectories) == 15 and len(sl.flatten()) == 15
    assert len(SessionList.from_list(sl.to_list())) == 2
    with pytest.raises(RuntimeError):
        SessionList(s, "x")
    maze = H.product_m9()
    fixed = adjust_positions_to_maze([(0.5, 0.5), (0.16, 0.16), (np.nan, 0.1)], maze)  # (4,4) is VOID
    assert maze.is_visitable_cont(fixed[0]) and fixed[1] == (0.16, 0.16) and np.isnan(fixed[2][0])
    recs = sl.step_streams(maze, possible_actions=H.a8())
    assert [len(r) for r in recs] == [9, 4]


def test_multiobjective_fitness_known_answers():
    fit = MultiObjectiveFitness(compare_against={"a": 0, "b": 1, "c": np.zeros(5)})
    fit.add_objective("obj1", manhattan_distance, "a", "b")
    fit.add_objective("obj2", normalized_distance, "c")
    with pytest.raises(RuntimeError):
        fit.add_objective("obj1", manhattan_distance, "a")
    d = fit({"a": 1, "b": 0, "c": np.ones(5)})
    npt.assert_allclose(d["obj1"], 2)
    npt.assert_allclose(d["obj2"], np.linalg.norm(np.ones(5) / 5))
    fit({"a": 3, "b": 0, "c": np.zeros(5)})
    npt.assert_allclose(fit.means()["obj1"], 3)
    npt.assert_allclose(fit.stds()["obj1"], 1)
    npt.assert_allclose(fit.percentiles_list()[0], (2.5, 3.5))
    fit.reset()
    assert fit.data()["obj1"] == []


def test_multi_result_views_and_batched_leaves():
    """MultiResult (results.py:4-101): nested addressing, statistics over the result axis, views that see later
    appends, and the batched form whose leaves already carry the result axis."""
    from navigation_model_b200.analysis.results import MultiResult
    mr = MultiResult({"a": {"b": 0.0}, "h": np.zeros(4)}, {"a": {"b": 1.5}, "h": np.ones(4)})
    view = mr["a"]["b"]
    mr.append({"a": {"b": 2.0}, "h": np.full(4, 2.0)})
    assert len(mr) == 3 and np.array_equal(view.data(), [0.0, 1.5, 2.0])  # the view follows the root
    assert view.median() == 1.5 and np.isclose(view.mean(), np.mean([0.0, 1.5, 2.0]))
    assert np.isclose(view.std(), np.std([0.0, 1.5, 2.0])) and view.percentile() == (0.75, 1.75)
    assert np.array_equal(mr["h"].mean(), np.ones(4)) and mr["h"].data().shape == (3, 4)
    with pytest.raises(KeyError):
        mr["missing"]
    with pytest.raises(TypeError):
        view.append({"a": {"b": 1.0}})
    batch = MultiResult.from_batch({"moving": np.array([3, 5, 10]), "hist": np.arange(12.0).reshape(3, 4)})
    assert len(batch["moving"]) == 3 and batch["moving"].median() == 5
    assert np.array_equal(batch["hist"].mean(), np.arange(12.0).reshape(3, 4).mean(axis=0))
    mr.reset()
    assert len(mr) == 0


def test_where_the_visit_map_of_a_maze_lives():
    """A 60 x 60 maze does not fit the simulation kernel's shared-memory visit map: the map moves to device memory
    (nm_simulate_fits == 2; the GPU test is tests/test_sim_gpu.py::test_large_maze_visit_map_in_device_memory) -- no
    silent CPU fallback for the built-in models; a user-defined model still takes the per-step plugin loop.  (Value-table
    sweeps: beyond 863 states their tables move to device memory too, tests/test_fit_gpu.py.)"""
    from navigation_model_b200 import _native
    from navigation_model_b200.simulation.behavioural_components import AnxietyComponent
    from navigation_model_b200.simulation.experiment import StandardExperiment
    from navigation_model_b200.simulation.exploration_model import BehaviouralModel
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.simulation.mouse import Mouse
    assert _native.lib().nm_vtable_max_states() == 863
    fits = _native.lib().nm_simulate_fits
    assert fits(9, 9, 10000) == 1 and fits(60, 60, 100) == 2 and fits(250, 250, 100) == 2 and fits(400, 400, 10) == 0
    assert _native.lib().nm_simulate_scratch_columns(1) == 128 and _native.lib().nm_simulate_scratch_columns(129) == 256
    layout = "\n".join(["M" * 60] * 60)
    maze = Maze(layout, "", size_tile=0.02)
    acts = np.array([(0, 0), (1, 0), (0, 1), (-1, 0)]) * 0.02
    mouse = Mouse(maze.get_tile_centre(30, 30), 0.0, acts)
    user_model = type("UserModel", (BehaviouralModel,), {})  # a plugin: per-step host loop, any maze size
    model = user_model([AnxietyComponent(W_anx=1.0, p1=0.5, p2=0.5, p3=0.5)], (30, 30), beta=1.0, seed=3)
    sim = StandardExperiment(model, maze, mouse)
    assert not sim._device_eligible(50)
    sim.run(50)
    assert len(mouse.history_position) == 51
    builtin = BehaviouralModel([AnxietyComponent(W_anx=1.0, p1=0.5, p2=0.5, p3=0.5)], (30, 30), beta=1.0, seed=3)
    assert StandardExperiment(builtin, maze, Mouse(maze.get_tile_centre(30, 30), 0.0, acts))._device_eligible(50)
    huge = Maze("\n".join(["M" * 400] * 400), "", size_tile=0.01)
    with pytest.raises(RuntimeError, match="too large for the simulation kernel"):
        StandardExperiment(builtin, huge, Mouse(huge.get_tile_centre(30, 30), 0.0, acts)).run(5)
