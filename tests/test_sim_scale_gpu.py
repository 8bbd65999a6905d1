"""Simulation parity AT SCALE (SURVEY 8(d) config 5 parity subset): 4 096 mice x 1 000 steps with
replayed draws through the CUDA kernel against the compiled oracle (oracle/nm_oracle.c:orc_simulate, itself bit-exact
on every history the unmodified reference produced for tests/golden/).

Bar: the DISCRETE trajectory (action chosen at every one of the 4.1e6 steps, visit maps, first-visit iterations, motion
counts) AND the continuous poses (positions, headings) are bit-identical for every mouse (north_star: "simulated
trajectories are bit-exact when fed the reference's own replayed uniform draws").  Statistics in units in the last place
go to gpurun_out/sim_scale_parity.json (copied to profiles/; round 1's CUDA-libm kernel: 23 % of the positions identical,
max 1.8e-14 absolute; round 2: 100 %)."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu

N, T = 4096, 1000
KINDS5 = [O.COMP_ANXIETY, O.COMP_BCOST, O.COMP_EXPERIENCE, O.COMP_BPERSISTENCE, O.COMP_VTABLE]


def _ulps(a, b):
    """|a - b| in units of the spacing of doubles at max(|a|, |b|)."""
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b)))


def test_four_million_agent_steps_against_the_oracle(cuda):
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    rng = np.random.default_rng(2025)
    maze, omz = H.product_m9(), H.oracle_m9()
    u = rng.uniform
    P = {"beta": u(0.3, 6, N), "W_anx": u(0, 10, N), "p1": u(0, 1, N), "p2": u(0, 1, N), "p3": u(0, 1, N),
         "W_bcost": u(0, 10, N), "k": u(0.05, 4, N), "gamma": u(0.2, 2, N), "W_exp": u(0, 10, N),
         "xi": u(0.006, 0.6, N), "kappa_home": u(0.01, 5, N), "kappa_explo": u(0.01, 5, N),
         "theta_explo": u(80, 120, N), "theta_home": u(30, 70, N), "W_bper": u(0, 10, N), "bp": u(0, 1, N),
         "Wm": u(0, 2, N), "Wnm": u(0, 2, N), "W_values": u(0, 10, N)}
    P["beta"][: N // 2] = u(0.02, 0.4, N // 2)  # half the population explores widely, half exploits
    V = rng.random((9, 9)) * 2 - 0.5
    acts = H.a8()
    cells = np.array([omz.vis_cells[i] for i in rng.integers(0, len(omz.vis_cells), N)])
    pos0 = np.array([maze.get_tile_centre(c) for c in cells]) + u(-0.04, 0.04, (N, 2))
    ori0 = u(-np.pi, np.pi, N)
    draws = rng.random((N, T))
    exp = BatchedExperiment(maze, acts, O.A8_UNITS, [bc.AnxietyComponent, bc.BiomechanicalCostComponent,
                                                     bc.ExperienceComponent, bc.BiomechPersistenceComponent,
                                                     bc.ValueTableComponent], v_table=V)
    res = exp.run(P, T, pos0, ori0, init_disc=cells, draws=draws, history=True, occupancy=True)
    # the von-Mises tables of the oracle side: scipy, vectorised over the mice (checked against the per-mouse form)
    from scipy.stats import vonmises
    angles = np.arctan2(O.A8_UNITS[:, 1], O.A8_UNITS[:, 0])
    tabs = vonmises.pdf(angles[None, :], P["k"][:, None])
    tabs[:, np.all(O.A8_UNITS == 0, axis=1)] = 0.0
    far = np.abs(O.A8_UNITS).max(axis=1) >= 2
    tabs[:, far] = tabs[:, far] * P["gamma"][:, None]
    for m in (0, 17, N - 1):
        assert np.array_equal(tabs[m], O.bcost_table(O.A8_UNITS, P["k"][m], P["gamma"][m]))
    pose = np.concatenate([pos0, ori0[:, None]], axis=1)
    ref = CO.simulate(omz, acts, O.A8_UNITS, KINDS5, P, pose, cells, draws, vtable_norm=O.normalise_vtable(V),
                      bcost_tables=tabs)
    assert not ref["status"].any() and not res.status.any()
    # ---- discrete level: identical, everywhere
    diverged = int((ref["choice"] != res.choices).any(axis=1).sum())
    assert diverged == 0, f"{diverged} of {N} discrete trajectories diverged"
    assert np.array_equal(ref["occupancy"], res.occupancy.astype(np.int32))
    st = maze.size_tile
    assert np.array_equal(np.floor_divide(res.history_position, st), np.floor_divide(ref["hist_pos"], st))
    moved = (np.diff(np.floor_divide(ref["hist_pos"], st), axis=1) != 0).any(axis=2).sum(axis=1)
    assert np.array_equal(moved, res.count_moving)
    # ---- continuous level: units in the last place
    up = _ulps(res.history_position, ref["hist_pos"])
    dori = res.history_orientation - ref["hist_ori"]
    flips = int((np.abs(np.abs(dori) - 2 * np.pi) < 1e-9).sum())  # atan2(+-0, -dx): +pi on one side, -pi on the other
    dori = np.angle(np.exp(1j * dori))
    uo = np.abs(dori) / np.spacing(np.maximum(np.abs(ref["hist_ori"]), 1e-3))
    stats = {
        "mice": N, "steps": T, "agent_steps": N * T, "discrete_divergences": diverged,
        "pose_changes": int((np.diff(ref["hist_pos"], axis=1) != 0).any(axis=2).sum()),
        "position_bit_identical_fraction": float((up == 0).mean()),
        "position_ulp_max": float(up.max()), "position_ulp_p50": float(np.percentile(up, 50)),
        "position_ulp_p99": float(np.percentile(up, 99)), "position_abs_max": float(
            np.abs(res.history_position - ref["hist_pos"]).max()),
        "position_ulp_histogram": {str(k): int(((up >= lo) & (up < hi)).sum()) for k, lo, hi in
                                   (("0", 0, 0.5), ("1", 0.5, 1.5), ("2-3", 1.5, 3.5), ("4-15", 3.5, 15.5),
                                    ("16-255", 15.5, 255.5), (">=256", 255.5, np.inf))},
        "heading_bit_identical_fraction": float((res.history_orientation == ref["hist_ori"]).mean()),
        "heading_ulp_max_mod_2pi": float(uo.max()), "heading_pi_flips": flips,
        "final_pose_bit_identical_mice": int((res.final_pose[:, :2] == ref["final_pose"][:, :2]).all(axis=1).sum()),
    }
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "sim_scale_parity.json"), "w") as fh:
        json.dump(stats, fh, indent=1)
    print(json.dumps(stats))
    # round 2: the kernel's cos / sin / atan2 are the host library's (csrc/nm_libm.h) -> every pose bit-identical
    assert np.array_equal(res.history_position, ref["hist_pos"])
    assert np.array_equal(res.history_orientation, ref["hist_ori"])
    assert np.array_equal(res.final_pose, ref["final_pose"])
