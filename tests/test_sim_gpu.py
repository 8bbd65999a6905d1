"""Parity of the CUDA forward simulation (through the C-ABI) with the reference's golden histories and
with the oracle, under replayed uniform draws.

Bar (BASELINE.json north_star / SURVEY section 7 "hard parts"): the DISCRETE trajectory -- chosen action
at every step, tile sequence, occupancy / exploration metrics -- is bit-exact, AND SO ARE THE CONTINUOUS POSES (positions
and headings): since round 2 the kernel evaluates cos / sin / atan2 with the host C library's algorithm
(csrc/nm_libm.h), so every comparison below is an equality (POSE_ATOL = 0)."""
import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu

POSE_ATOL = 0.0


def _classes():
    from navigation_model_b200.simulation import behavioural_components as bc
    return [bc.AnxietyComponent, bc.BiomechanicalCostComponent, bc.ExperienceComponent,
            bc.BiomechPersistenceComponent, bc.ValueTableComponent]


def _oracle_components(P, m, units, V):
    return [(O.COMP_ANXIETY, {k: P[k][m] for k in ("W_anx", "p1", "p2", "p3")}),
            (O.COMP_BCOST, {"W_bcost": P["W_bcost"][m], "table": O.bcost_table(units, P["k"][m], P["gamma"][m])}),
            (O.COMP_EXPERIENCE, {k: P[k][m] for k in ("W_exp", "xi", "kappa_home", "kappa_explo", "theta_explo",
                                                      "theta_home")}),
            (O.COMP_BPERSISTENCE, {k: P[k][m] for k in ("W_bper", "bp", "Wm", "Wnm")}),
            (O.COMP_VTABLE, {"W_values": P["W_values"][m], "table": O.normalise_vtable(V)})]


def _random_params(rng, N):
    u = rng.uniform
    return {"beta": u(0.3, 6, N), "W_anx": u(0, 10, N), "p1": u(0, 1, N), "p2": u(0, 1, N), "p3": u(0, 1, N),
            "W_bcost": u(0, 10, N), "k": u(0.05, 4, N), "gamma": u(0.2, 2, N), "W_exp": u(0, 10, N),
            "xi": u(0.006, 0.6, N), "kappa_home": u(0.01, 5, N), "kappa_explo": u(0.01, 5, N),
            "theta_explo": u(80, 120, N), "theta_home": u(30, 70, N), "W_bper": u(0, 10, N), "bp": u(0, 1, N),
            "Wm": u(0, 2, N), "Wnm": u(0, 2, N), "W_values": u(0, 10, N)}


@pytest.mark.parametrize("counters", ["16", "32"])
def test_golden_histories(cuda, golden_sim, counters, monkeypatch):
    """The reference's own StandardExperiment.run histories (6 mice x 600 steps, 5 components), with both widths of
    the kernel's visit counters (16 bits is what a run of fewer than 65 535 steps uses)."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    monkeypatch.setenv("NM_SIM_COUNTERS", counters)
    g = golden_sim
    maze = H.product_m9()
    names = [str(s) for s in g["param_names"]]
    P = {n: g["params"][:, i] for i, n in enumerate(names)}
    T = int(g["T"])
    exp = BatchedExperiment(maze, g["actions"], g["actions_units"], _classes(), v_table=g["v_table"])
    for pct, key in ((80, "it_visit_80"), (50, "it_visit_50")):
        res = exp.run(P, T, g["init_pose"][:, :2], g["init_pose"][:, 2], init_disc=g["init_disc"], draws=g["draws"],
                      visit_percentage=pct, history=True, occupancy=True, occupancy_total=True)
        assert not res.status.any()
        omz = H.oracle_m9()
        for m in range(len(g["seeds"])):
            want_disc = [omz.cont2disc(x, y) for x, y in g["hist_pos"][m]]
            got_disc = [omz.cont2disc(x, y) for x, y in res.history_position[m]]
            assert got_disc == want_disc, f"mouse {m}: tile sequence differs"
            np.testing.assert_allclose(res.history_position[m], g["hist_pos"][m], rtol=0, atol=POSE_ATOL)
            assert np.array_equal(res.history_orientation[m], g["hist_ori"][m])  # headings: the reference's bits
        assert np.array_equal(res.occupancy, g["occupancy"].astype(np.int64))
        assert np.array_equal(res.it_visit, g[key])
        assert np.array_equal(res.count_moving, g["count_moving"])
        assert np.array_equal(res.tile_counts[:, :4], g["tile_counts"])
        assert np.array_equal(res.occupancy_total, g["occupancy"].sum(0).astype(np.int64))
        np.testing.assert_allclose(res.final_pose[:, :2], g["hist_pos"][:, -1], rtol=0, atol=POSE_ATOL)


def test_device_pcg64_equals_replayed_draws(cuda, golden_sim):
    """Seeds (numpy-identical PCG64 on the device) and replayed draws give the same run."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    g = golden_sim
    maze = H.product_m9()
    names = [str(s) for s in g["param_names"]]
    P = {n: g["params"][:, i] for i, n in enumerate(names)}
    T = int(g["T"])
    exp = BatchedExperiment(maze, g["actions"], g["actions_units"], _classes(), v_table=g["v_table"])
    kw = dict(init_pos=g["init_pose"][:, :2], init_ori=g["init_pose"][:, 2], init_disc=g["init_disc"], history=True)
    a = exp.run(P, T, draws=g["draws"], **kw)
    b = exp.run(P, T, seeds=g["seeds"], **kw)
    assert np.array_equal(a.choices, b.choices)
    assert np.array_equal(a.history_position, b.history_position)
    # the generator ends where numpy's ends after T random() calls
    for m, seed in enumerate(g["seeds"]):
        gen = np.random.default_rng(int(seed))
        gen.random(T)
        assert np.array_equal(b.pcg_state[m], _words(gen))


def _words(gen):
    st = gen.bit_generator.state["state"]
    m = (1 << 64) - 1
    return np.array([st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)


def test_standard_experiment_api_matches_reference(cuda, golden_sim):
    """StandardExperiment / BehaviouralModel / components / Mouse of this repo, driven like the reference
    (seeded model, device path), leave the same history and the same object state."""
    from navigation_model_b200.simulation import behavioural_components as bc
    from navigation_model_b200.simulation.experiment import StandardExperiment
    from navigation_model_b200.simulation.exploration_model import BehaviouralModel
    from navigation_model_b200.simulation.mouse import Mouse
    g = golden_sim
    maze = H.product_m9()
    names = [str(s) for s in g["param_names"]]
    units = [tuple(int(v) for v in a) for a in g["actions_units"]]
    T = int(g["T"])
    for m in (0, 3):
        p = dict(zip(names, g["params"][m]))
        comps = [bc.AnxietyComponent(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"]),
                 bc.BiomechanicalCostComponent(units, W_bcost=p["W_bcost"], k=p["k"], gamma=p["gamma"]),
                 bc.ExperienceComponent(maze.shape, W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                        kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"],
                                        theta_home=p["theta_home"]),
                 bc.BiomechPersistenceComponent(units, W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"]),
                 bc.ValueTableComponent(g["v_table"], W_values=p["W_values"])]
        model = BehaviouralModel(comps, tuple(int(v) for v in g["init_disc"][m]), beta=p["beta"], seed=int(g["seeds"][m]))
        mouse = Mouse(g["init_pose"][m][:2], g["init_pose"][m][2], g["actions"])
        exp = StandardExperiment(model, maze, mouse)
        assert exp._device_eligible()
        exp.run(T)
        hp, ho = mouse.history
        assert len(hp) == T + 1
        np.testing.assert_allclose(np.array(hp), g["hist_pos"][m], rtol=0, atol=POSE_ATOL)
        np.testing.assert_allclose(mouse.position, g["hist_pos"][m][-1], rtol=0, atol=POSE_ATOL)
        # experience map = occupancy of history[0..T-1]
        omz = H.oracle_m9()
        occ = g["occupancy"][m].copy()
        lx, ly = omz.cont2disc(*g["hist_pos"][m][-1])
        occ[lx, ly] -= 1
        assert np.array_equal(comps[2]._maze_map_exp, occ)
        # the model's generator advanced by exactly T draws
        twin = np.random.default_rng(int(g["seeds"][m]))
        twin.random(T)
        assert model._rng.random() == twin.random()


def test_random_batch_vs_oracle(cuda):
    """64 mice, random parameters inside the descriptor bounds, jittered starts, 300 steps."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    rng = np.random.default_rng(99)
    maze, omz = H.product_m9(), H.oracle_m9()
    N, T = 64, 300
    acts = H.a8()
    V = rng.random((9, 9)) * 2 - 0.5
    P = _random_params(rng, N)
    starts = [(1, 1), (7, 4), (1, 6), (2, 8), (6, 8), (8, 0)]
    cells = np.array([starts[i % len(starts)] for i in range(N)])
    pos0 = np.array([maze.get_tile_centre(c) for c in cells]) + rng.uniform(-0.04, 0.04, (N, 2))
    ori0 = rng.uniform(-np.pi, np.pi, N)
    draws = rng.random((N, T))
    exp = BatchedExperiment(maze, acts, O.A8_UNITS, _classes(), v_table=V)
    res = exp.run(P, T, pos0, ori0, init_disc=cells, draws=draws, history=True, occupancy=True)
    mismatched = 0
    for m in range(N):
        ref = O.simulate(omz, acts, O.A8_UNITS, pos0[m], ori0[m], cells[m], P["beta"][m],
                         _oracle_components(P, m, O.A8_UNITS, V), draws[m])
        assert ref["status"] == 0 and res.status[m] == 0
        if not np.array_equal(ref["choice"], res.choices[m]):
            mismatched += 1
            continue
        np.testing.assert_allclose(res.history_position[m], ref["hist_pos"], rtol=0, atol=POSE_ATOL)
        met = O.metrics_of_history(omz, ref["hist_pos"])
        assert np.array_equal(met["occupancy"], res.occupancy[m])
        assert met["it_visit"] == res.it_visit[m] and met["count_moving"] == res.count_moving[m]
        assert np.array_equal(met["tile_counts"], res.tile_counts[m])
    assert mismatched == 0, f"{mismatched} of {N} discrete trajectories diverged"


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_mazes_vs_oracle(cuda, seed):
    """Irregular 7 x 11 mazes with scattered void tiles (ray walks past void corners, dead ends), a random subset of the
    eight actions that keeps the null action, 48 mice x 150 steps with replayed draws: choices, poses, occupancy."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation.maze import Maze
    rng = np.random.default_rng(seed)
    X, Y, st = 7, 11, 0.1
    rows = [[("V" if rng.random() < 0.22 else "CWM"[rng.integers(3)]) for _ in range(Y)] for _ in range(X)]
    rows[1][1], rows[1][2] = "M", "W"
    layout = "\n".join("".join(r) for r in rows)
    maze, omz = Maze(layout, "", size_tile=st), O.OracleMaze(layout, st)
    keep = np.concatenate([[0], 1 + np.sort(rng.permutation(7)[:int(rng.integers(3, 8))])])
    units = O.A8_UNITS[keep]
    acts = units * st
    N, T = 48, 150
    V = rng.random((X, Y))
    P = _random_params(rng, N)
    cells = np.array([omz.vis_cells[i] for i in rng.integers(0, len(omz.vis_cells), N)])
    pos0 = np.array([maze.get_tile_centre(c) for c in cells]) + rng.uniform(-0.03, 0.03, (N, 2))
    ori0 = rng.uniform(-np.pi, np.pi, N)
    draws = rng.random((N, T))
    exp = BatchedExperiment(maze, acts, units, _classes(), v_table=V)
    res = exp.run(P, T, pos0, ori0, init_disc=cells, draws=draws, history=True, occupancy=True)
    for m in range(N):
        ref = O.simulate(omz, acts, units, pos0[m], ori0[m], cells[m], P["beta"][m],
                         _oracle_components(P, m, units, V), draws[m])
        assert ref["status"] == res.status[m] == 0
        assert np.array_equal(ref["choice"], res.choices[m]), (seed, m)
        np.testing.assert_allclose(res.history_position[m], ref["hist_pos"], rtol=0, atol=POSE_ATOL)
        met = O.metrics_of_history(omz, ref["hist_pos"])
        assert np.array_equal(met["occupancy"], res.occupancy[m])
        assert met["it_visit"] == res.it_visit[m] and met["count_moving"] == res.count_moving[m]


def test_component_subsets_order_and_inactive(cuda):
    """Any ordered subset of components; an inactive component contributes nothing."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    rng = np.random.default_rng(5)
    maze, omz = H.product_m9(), H.oracle_m9()
    N, T = 8, 120
    acts = H.a8()
    V = rng.random((9, 9))
    P = _random_params(rng, N)
    pos0 = np.array(maze.get_tile_centre(7, 4)) + rng.uniform(-0.03, 0.03, (N, 2))
    draws = rng.random((N, T))
    cls = _classes()
    for order, active in (([4, 2, 0], None), ([1, 3], None), ([0, 1, 2, 3, 4], [True, False, True, False, True])):
        comps = [cls[i] for i in order]
        exp = BatchedExperiment(maze, acts, O.A8_UNITS, comps, v_table=V, active=active)
        res = exp.run(P, T, pos0, 0.3, init_disc=(7, 4), draws=draws, history=True)
        for m in range(N):
            oc = _oracle_components(P, m, O.A8_UNITS, V)
            sel = [oc[i] for k, i in enumerate(order) if active is None or active[k]]
            ref = O.simulate(omz, acts, O.A8_UNITS, pos0[m], 0.3, (7, 4), P["beta"][m], sel, draws[m])
            assert np.array_equal(ref["choice"], res.choices[m]), (order, m)


def test_device_ray_check_matches_oracle(cuda):
    """Maze.is_movement_possible_cont on the device (FMA-residual floor division) vs the oracle's Python
    float floor division, including segments whose crossings land exactly on tile corners."""
    from navigation_model_b200.batched_sim import movements_possible
    maze, omz = H.product_m9(), H.oracle_m9()
    rng = np.random.default_rng(17)
    st = O.M9_SIZE_TILE
    segs = [rng.uniform(-0.1, 1.1, (4000, 4))]
    # centre-to-centre moves (diagonals cross grid corners exactly)
    cells = np.array(omz.vis_cells)
    a = cells[rng.integers(len(cells), size=3000)]
    d = rng.integers(-2, 3, size=(3000, 2))
    ca = np.array([omz.centre(c) for c in a])
    cb = np.array([omz.centre(c) for c in a + d])
    segs.append(np.concatenate([ca, cb], axis=1))
    # end points exactly on tile boundaries
    k = rng.integers(0, 10, size=(2000, 4)).astype(np.float64)
    segs.append(k * st)
    segs = np.concatenate(segs)
    got = movements_possible(maze, segs)
    want = np.array([omz.move_possible(*s) for s in segs])
    assert np.array_equal(got, want)
    assert want.any() and not want.all()


def test_stranded_mouse_reports_status(cuda):
    """No reachable endpoint: the reference raises from np.max([]); the kernel reports the failing step."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    maze = H.product_m9()
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, [bc.AnxietyComponent])
    P = {"beta": [1.0, 1.0], "W_anx": [1.0, 1.0], "p1": [1.0, 1.0], "p2": [0.5, 0.5], "p3": [0.5, 0.5]}
    pos0 = np.array([maze.get_tile_centre(1, 1), maze.get_tile_centre(4, 3)])  # (4,3) is VOID
    res = exp.run(P, 10, pos0, 0.0, draws=np.full((2, 10), 0.5))
    assert res.status[0] == 0 and res.status[1] == 1


def test_parameter_validation(cuda):
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    maze = H.product_m9()
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, [bc.ExperienceComponent])
    good = {"beta": 1.0, "W_exp": 1.0, "xi": 0.01, "kappa_home": 1.0, "kappa_explo": 1.0, "theta_explo": 100.0,
            "theta_home": 50.0}
    with pytest.raises(RuntimeError, match="Wrong value set for property 'xi'"):
        exp.run(dict(good, xi=0.9), 5, maze.get_tile_centre(1, 1), draws=np.full((1, 5), 0.5))
    missing = dict(good)
    del missing["kappa_home"]
    with pytest.raises(RuntimeError, match="kappa_home has not been specified"):
        exp.run(missing, 5, maze.get_tile_centre(1, 1), draws=np.full((1, 5), 0.5))


def test_random_model_matches_numpy_choice(cuda):
    """RandomModel on the device: Generator.choice(a) == a[integers(0, K)] with numpy's buffered 32-bit Lemire
    draws, reproduced from the PCG64 state; checked against the host per-step loop driven by numpy itself."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation.experiment import StandardExperiment
    from navigation_model_b200.simulation.exploration_model import RandomModel
    from navigation_model_b200.simulation.mouse import Mouse
    maze = H.product_m9()
    acts = H.a8()
    T = 400
    for seed, start, ori in ((11, (1, 1), 0.0), (12, (7, 4), 2.1), (13, (2, 8), -1.3)):
        pos = np.array(maze.get_tile_centre(start)) + np.array([0.013, -0.021])
        # host plugin loop (numpy's own rng.choice)
        m_host = Mouse(pos, ori, acts)
        model_h = RandomModel(seed=seed)
        exp_h = StandardExperiment(model_h, maze, m_host)
        m_host.enable_history(True)
        exp_h._run_host(T)
        # device path through the same class
        m_dev = Mouse(pos, ori, acts)
        model_d = RandomModel(seed=seed)
        exp_d = StandardExperiment(model_d, maze, m_dev)
        assert exp_d._device_eligible()
        exp_d.run(T)
        hp_h, hp_d = np.array(m_host.history_position), np.array(m_dev.history_position)
        omz = H.oracle_m9()
        assert [omz.cont2disc(*p) for p in hp_h] == [omz.cont2disc(*p) for p in hp_d]
        np.testing.assert_allclose(hp_d, hp_h, rtol=0, atol=POSE_ATOL)
        # both generators end in the same state (including the buffered 32-bit half)
        assert model_h._rng.bit_generator.state == model_d._rng.bit_generator.state
    # batched: many seeds at once, metrics fused
    exp = BatchedExperiment(maze, acts, O.A8_UNITS, [], model="random")
    res = exp.run({}, 200, maze.get_tile_centre(1, 1), 0.0, seeds=list(range(64)), occupancy=True)
    assert not res.status.any() and (res.occupancy.sum(axis=(1, 2)) == 201).all()


def test_device_seeded_streams(cuda):
    """device_seed: per-mouse PCG64 streams seeded on the device -- deterministic, distinct across mice and seeds."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    maze = H.product_m9()
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, [bc.AnxietyComponent])
    P = {"beta": 1.5, "W_anx": 2.0, "p1": 0.9, "p2": 0.5, "p3": 0.4}
    kw = dict(init_pos=maze.get_tile_centre(1, 1), init_ori=0.0, init_disc=(1, 1), history=True, n_mice=512)
    a = exp.run(P, 150, device_seed=7, **kw)
    b = exp.run(P, 150, device_seed=7, **kw)
    c = exp.run(P, 150, device_seed=8, **kw)
    assert not a.status.any() and np.array_equal(a.choices, b.choices)
    assert not np.array_equal(a.choices, c.choices)
    assert len({tuple(r) for r in a.choices}) > 500          # the mice do not share a stream
    freq = np.bincount(a.choices.ravel(), minlength=8) / a.choices.size
    assert freq.max() < 0.6 and (freq > 0).sum() >= 6          # a softmax policy, not a stuck generator
    with pytest.raises(ValueError):
        exp.run(P, 10, device_seed=1, seeds=[1], **kw)


def test_results_survive_the_next_run(cuda):
    """The per-mouse results come back through a pinned staging buffer that the next run reuses: what a run returned
    must not change when the experiment runs again (large enough for the threaded copy-out, > 16 MB of results)."""
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    maze = H.product_m9()
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, [bc.AnxietyComponent])
    P = {"beta": 1.5, "W_anx": 2.0, "p1": 0.9, "p2": 0.5, "p3": 0.4}
    N = 120_000
    kw = dict(init_pos=maze.get_tile_centre(1, 1), init_ori=0.0, init_disc=(1, 1), n_mice=N, shock_zone={"tile_lines": 5},
              corridors=(1, 7), static_intervals=True, subareas=True, orientation_histogram={})
    a = exp.run(P, 40, device_seed=3, **kw)
    fields = ("final_pose", "status", "it_visit", "count_moving", "tile_counts", "pcg_state", "model_state",
              "zone_latency", "zone_occupancy", "zone_entries", "corridor_switches", "static_intervals",
              "subarea_counts", "orientations_histogram")
    kept = {f: np.array(getattr(a, f), copy=True) for f in fields}
    b = exp.run(P, 40, device_seed=4, **kw)
    assert not np.array_equal(b.final_pose, kept["final_pose"])
    for f in fields:
        assert np.array_equal(getattr(a, f), kept[f]), f
    c = exp.run(P, 40, device_seed=3, **kw)  # and the threaded copy returns the same bytes as the first time
    for f in fields:
        assert np.array_equal(getattr(c, f), kept[f]), f
    assert a.tile_counts.sum(axis=1).min() == 41 and a.orientations_histogram.shape[0] == N


@pytest.mark.parametrize("forced", [False, True])
def test_large_maze_visit_map_in_device_memory(cuda, forced, monkeypatch):
    """beyond ~50 x 50 tiles the simulation used to fall back to a Python per-step loop.  The visit maps now
    move to device memory (nm_sim_kernel<..., GMAP>): a 70 x 64 maze with void blocks, 96 mice x 200 steps against the
    compiled oracle -- choices, poses, visit maps bit for bit; and the 9 x 9 maze forced onto the same variant
    (NM_SIM_MAP=global) reproduces the shared-memory run exactly, metrics included."""
    from navigation_model_b200 import _native
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation.maze import Maze
    from oracle import c_oracle as CO
    rng = np.random.default_rng(8)
    if forced:
        maze, omz = H.product_m9(), H.oracle_m9()
        X, Y, st = 9, 9, O.M9_SIZE_TILE
    else:
        X, Y, st = 70, 64, 0.02
        rows = [["M"] * Y for _ in range(X)]
        for bx in range(5, X - 5, 9):
            for by in range(4, Y - 4, 7):
                for dx in range(3):
                    for dy in range(2):
                        rows[bx + dx][by + dy] = "V"
        layout = "\n".join("".join(r) for r in rows)
        maze, omz = Maze(layout, "", size_tile=st), O.OracleMaze(layout, st)
        assert _native.lib().nm_simulate_fits(X, Y, 200) == 2
    N, T = 96, 200
    acts = O.A8_UNITS * st
    P = _random_params(rng, N)
    P["beta"] = rng.uniform(0.02, 0.6, N)
    V = rng.random((X, Y))
    cells = np.array([omz.vis_cells[i] for i in rng.integers(0, len(omz.vis_cells), N)])
    pos0 = np.array([maze.get_tile_centre(c) for c in cells]) + rng.uniform(-0.3, 0.3, (N, 2)) * st
    ori0 = rng.uniform(-np.pi, np.pi, N)
    draws = rng.random((N, T))
    exp = BatchedExperiment(maze, acts, O.A8_UNITS, _classes(), v_table=V)
    kw = dict(init_disc=cells, draws=draws, history=True, occupancy=True, occupancy_total=True)
    if forced:
        base = exp.run(P, T, pos0, ori0, **kw)
        monkeypatch.setenv("NM_SIM_MAP", "global")
    res = exp.run(P, T, pos0, ori0, **kw)
    if forced:
        for f in ("choices", "history_position", "history_orientation", "occupancy", "occupancy_total", "it_visit",
                  "count_moving", "tile_counts", "final_pose", "model_state"):
            assert np.array_equal(getattr(res, f), getattr(base, f)), f
    ref = CO.simulate(omz, acts, O.A8_UNITS, [O.COMP_ANXIETY, O.COMP_BCOST, O.COMP_EXPERIENCE, O.COMP_BPERSISTENCE,
                                              O.COMP_VTABLE], P, np.concatenate([pos0, ori0[:, None]], axis=1), cells,
                      draws, vtable_norm=O.normalise_vtable(V))
    assert not ref["status"].any() and not res.status.any()
    assert np.array_equal(ref["choice"], res.choices)
    assert np.array_equal(ref["hist_pos"], res.history_position) and np.array_equal(ref["hist_ori"], res.history_orientation)
    assert np.array_equal(ref["occupancy"], res.occupancy.astype(np.int32))
    assert np.array_equal(res.occupancy_total, ref["occupancy"].sum(0))


def test_standard_experiment_on_a_large_maze_runs_on_the_device(cuda):
    from navigation_model_b200.simulation.behavioural_components import AnxietyComponent, ExperienceComponent
    from navigation_model_b200.simulation.experiment import StandardExperiment
    from navigation_model_b200.simulation.exploration_model import BehaviouralModel
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.simulation.mouse import Mouse
    maze = Maze("\n".join(["M" * 60] * 60), "", size_tile=0.02)
    acts = np.array([(0, 0), (1, 0), (0, 1), (-1, 0), (1, 1)]) * 0.02

    def run(model_cls):
        comps = [AnxietyComponent(W_anx=1.0, p1=0.5, p2=0.5, p3=0.5),
                 ExperienceComponent(maze.shape, W_exp=2.0, xi=0.1, kappa_home=1.0, kappa_explo=1.0, theta_explo=100.0,
                                     theta_home=50.0)]
        mouse = Mouse(maze.get_tile_centre(30, 30), 0.3, acts)
        sim = StandardExperiment(model_cls(comps, (30, 30), beta=0.7, seed=3), maze, mouse)
        eligible = sim._device_eligible(150)
        sim.run(150)
        return eligible, np.array(mouse.history[0]), np.array(mouse.history[1]), comps[1]._maze_map_exp.copy()

    on_device = run(BehaviouralModel)
    plugin = run(type("UserModel", (BehaviouralModel,), {}))  # the reference's per-step loop on the host
    assert on_device[0] and not plugin[0]
    for a, b in zip(on_device[1:], plugin[1:]):
        assert np.array_equal(a, b)
