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
