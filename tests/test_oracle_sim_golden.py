"""The simulation oracle against histories / probabilities / metrics produced by the unmodified reference
(tests/golden/sim_golden.npz, written by tests/golden/make_golden_sim.py)."""
import numpy as np

from oracle import np_oracle as O

import helpers as H


def components_of(g, m):
    p = dict(zip([str(s) for s in g["param_names"]], g["params"][m]))
    comps = [
        (O.COMP_ANXIETY, dict(W_anx=p["W_anx"], p1=p["p1"], p2=p["p2"], p3=p["p3"])),
        (O.COMP_BCOST, dict(W_bcost=p["W_bcost"], table=O.bcost_table(g["actions_units"], p["k"], p["gamma"]))),
        (O.COMP_EXPERIENCE, dict(W_exp=p["W_exp"], xi=p["xi"], kappa_home=p["kappa_home"],
                                 kappa_explo=p["kappa_explo"], theta_explo=p["theta_explo"],
                                 theta_home=p["theta_home"])),
        (O.COMP_BPERSISTENCE, dict(W_bper=p["W_bper"], bp=p["bp"], Wm=p["Wm"], Wnm=p["Wnm"])),
        (O.COMP_VTABLE, dict(W_values=p["W_values"], table=O.normalise_vtable(g["v_table"]))),
    ]
    return p, comps


def test_simulation_histories_match_reference(golden_sim):
    g = golden_sim
    mz = H.oracle_m9()
    assert np.array_equal(O.normalise_vtable(g["v_table"]), g["vtable_norm"])
    for m in range(len(g["seeds"])):
        p, comps = components_of(g, m)
        if m == 0:
            assert np.array_equal(comps[1][1]["table"], g["bcost_table_m0"])
        res = O.simulate(mz, g["actions"], g["actions_units"], g["init_pose"][m][:2], g["init_pose"][m][2],
                         g["init_disc"][m], p["beta"], comps, g["draws"][m], record_probs=(m == 1))
        assert res["status"] == 0
        assert np.array_equal(res["hist_pos"], g["hist_pos"][m]), m      # bit-exact continuous trajectory
        assert np.array_equal(res["hist_ori"], g["hist_ori"][m]), m
        if m == 1:
            for t, pv in enumerate(res["probs"]):
                assert len(pv) == g["m1_ncand"][t]
                assert np.array_equal(pv, g["m1_probs"][t, :len(pv)])
        met = O.metrics_of_history(mz, res["hist_pos"], 80)
        assert np.array_equal(met["occupancy"], g["occupancy"][m])
        assert met["it_visit"] == g["it_visit_80"][m]
        assert O.metrics_of_history(mz, res["hist_pos"], 50)["it_visit"] == g["it_visit_50"][m]
        assert met["count_moving"] == g["count_moving"][m]
        assert np.array_equal(met["tile_counts"][:4], g["tile_counts"][m])


def test_draws_are_the_generator_stream(golden_sim):
    """Generator.choice(p=...) consumes one random() per call; random() = (PCG64 u64 >> 11) * 2**-53."""
    g = golden_sim
    seed = int(g["seeds"][0])
    words = O.pcg64_state_words(seed)
    state = (int(words[0]) << 64) | int(words[1])
    inc = (int(words[2]) << 64) | int(words[3])
    mult = 0x2360ED051FC65DA44385DF649FCCF645
    mask = (1 << 128) - 1
    out = []
    for _ in range(5):
        state = (state * mult + inc) & mask
        hi, lo = state >> 64, state & ((1 << 64) - 1)
        x, rot = hi ^ lo, hi >> 58
        u64 = ((x >> rot) | (x << ((64 - rot) & 63))) & ((1 << 64) - 1)
        out.append((u64 >> 11) * (1.0 / 9007199254740992.0))
    assert np.array_equal(out, g["draws"][0][:5])


# ---- the compiled twin (oracle/nm_oracle.c:orc_simulate) against the same reference outputs --------------------
KINDS5 = [O.COMP_ANXIETY, O.COMP_BCOST, O.COMP_EXPERIENCE, O.COMP_BPERSISTENCE, O.COMP_VTABLE]


def _cols(names, rows):
    return {str(n): np.ascontiguousarray(rows[:, i]) for i, n in enumerate(names)}


def test_c_simulation_oracle_matches_reference_m9(golden_sim):
    from oracle import c_oracle as CO
    g = golden_sim
    mz = H.oracle_m9()
    res = CO.simulate(mz, g["actions"], g["actions_units"], KINDS5, _cols(g["param_names"], g["params"]), g["init_pose"],
                      g["init_disc"], g["draws"], vtable_norm=O.normalise_vtable(g["v_table"]))
    assert not res["status"].any()
    assert np.array_equal(res["hist_pos"], g["hist_pos"])  # bit-exact continuous trajectories, 6 mice x 600 steps
    assert np.array_equal(res["hist_ori"], g["hist_ori"])
    assert np.array_equal(res["occupancy"], g["occupancy"].astype(np.int32))


def test_c_simulation_oracle_matches_reference_random_mazes_and_grid_tiles():
    import os
    from conftest import GOLDEN
    from oracle import c_oracle as CO
    g = dict(np.load(os.path.join(GOLDEN, "random_mazes_golden.npz"), allow_pickle=False))
    st = float(g["size_tile"])
    for seed in g["seeds"]:
        k = f"s{seed}_"
        mz = O.OracleMaze(str(g[k + "layout"]), st)
        units = g[k + "units"]
        res = CO.simulate(mz, units * st, units, KINDS5, _cols(g["param_names"], g[k + "sim_params"]), g[k + "sim_pose"],
                          g[k + "sim_disc"], g[k + "sim_draws"], vtable_norm=O.normalise_vtable(g[k + "v_table"]))
        assert np.array_equal(res["hist_pos"], g[k + "sim_hist_pos"]), seed
        assert np.array_equal(res["hist_ori"], g[k + "sim_hist_ori"]), seed
    g = dict(np.load(os.path.join(GOLDEN, "grid_golden.npz"), allow_pickle=False))
    st = float(g["size_tile"])
    mz = O.OracleMaze(str(g["layout"]), st, grid_tiles=True)
    res = CO.simulate(mz, g["units"] * st, g["units"], KINDS5, _cols(g["param_names"], g["params"]), g["pose"],
                      g["init_disc"], g["draws"], vtable_norm=O.normalise_vtable(g["v_table"]))
    assert np.array_equal(res["hist_pos"], g["hist_pos"]) and np.array_equal(res["hist_ori"], g["hist_ori"])
    assert np.array_equal(res["occupancy"], g["occupancy"].astype(np.int32))


def test_c_simulation_oracle_equals_numpy_oracle_on_fresh_mice():
    """200 fresh mice x 150 steps (random parameters, jittered starts): the two restatements agree bit for bit."""
    from oracle import c_oracle as CO
    mz = H.oracle_m9()
    rng = np.random.default_rng(31)
    N, T = 40, 150
    u = rng.uniform
    P = {"beta": u(0.3, 6, N), "W_anx": u(0, 10, N), "p1": u(0, 1, N), "p2": u(0, 1, N), "p3": u(0, 1, N),
         "W_bcost": u(0, 10, N), "k": u(0.05, 4, N), "gamma": u(0.2, 2, N), "W_exp": u(0, 10, N),
         "xi": u(0.006, 0.6, N), "kappa_home": u(0.01, 5, N), "kappa_explo": u(0.01, 5, N),
         "theta_explo": u(80, 120, N), "theta_home": u(30, 70, N), "W_bper": u(0, 10, N), "bp": u(0, 1, N),
         "Wm": u(0, 2, N), "Wnm": u(0, 2, N), "W_values": u(0, 10, N)}
    V = rng.random((9, 9))
    acts = O.A8_UNITS * O.M9_SIZE_TILE
    cells = mz.vis_cells
    start = [cells[int(rng.integers(len(cells)))] for _ in range(N)]
    pose = np.array([[mz.centre(c)[0] + u(-0.03, 0.03), mz.centre(c)[1] + u(-0.03, 0.03), u(-3.1, 3.1)] for c in start])
    draws = rng.random((N, T))
    res = CO.simulate(mz, acts, O.A8_UNITS, KINDS5, P, pose, np.array(start), draws, vtable_norm=O.normalise_vtable(V))
    for m in range(N):
        comps = [(O.COMP_ANXIETY, {k: P[k][m] for k in ("W_anx", "p1", "p2", "p3")}),
                 (O.COMP_BCOST, {"W_bcost": P["W_bcost"][m], "table": O.bcost_table(O.A8_UNITS, P["k"][m], P["gamma"][m])}),
                 (O.COMP_EXPERIENCE, {k: P[k][m] for k in ("W_exp", "xi", "kappa_home", "kappa_explo", "theta_explo",
                                                           "theta_home")}),
                 (O.COMP_BPERSISTENCE, {k: P[k][m] for k in ("W_bper", "bp", "Wm", "Wnm")}),
                 (O.COMP_VTABLE, {"W_values": P["W_values"][m], "table": O.normalise_vtable(V)})]
        ref = O.simulate(mz, acts, O.A8_UNITS, pose[m, :2], pose[m, 2], start[m], P["beta"][m], comps, draws[m])
        assert ref["status"] == res["status"][m] == 0
        assert np.array_equal(ref["choice"], res["choice"][m])
        assert np.array_equal(ref["hist_pos"], res["hist_pos"][m]) and np.array_equal(ref["hist_ori"], res["hist_ori"][m])
