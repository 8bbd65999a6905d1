"""BASELINE configs[4] at full WIDTH (2^20 device-seeded mice, the population of bench.py) through size-independent
properties plus a sample checked against the oracle.

The device seeds every mouse's PCG64 stream itself (csrc/nm_sim.cu: splitmix64 of (device_seed, global mouse index) ->
pcg64_srandom); the host restates that seeding here, turns a sample's streams into the uniforms numpy's ``random()`` would
return, and replays them through the compiled oracle: final poses bit-identical, per-mouse counters equal.  Properties
over the whole population: every mouse accounts for n_steps + 1 positions, the population occupancy map is the sum of the
per-mouse maps, the same seed gives the same bytes, and a mouse does not depend on where in the launch it sits."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu

N, T = 1 << 20, 200
M64 = (1 << 64) - 1
M128 = (1 << 128) - 1
PCG_MULT = (0x2360ED051FC65DA4 << 64) | 0x4385DF649FCCF645
KINDS5 = [O.COMP_ANXIETY, O.COMP_BCOST, O.COMP_EXPERIENCE, O.COMP_BPERSISTENCE, O.COMP_VTABLE]


def device_stream(seed, index, n):
    """The first n uniforms of mouse `index` under device_seed `seed` (restates the seeding in nm_sim_kernel and numpy's
    PCG64 XSL-RR output / random())."""
    z = (seed + 0x9E3779B97F4A7C15 * (4 * index + 1)) & M64
    w = []
    for _ in range(4):
        z = (z + 0x9E3779B97F4A7C15) & M64
        y = z
        y = ((y ^ (y >> 30)) * 0xBF58476D1CE4E5B9) & M64
        y = ((y ^ (y >> 27)) * 0x94D049BB133111EB) & M64
        w.append(y ^ (y >> 31))
    initstate, initseq = (w[0] << 64) | w[1], (w[2] << 64) | w[3]
    inc = ((initseq << 1) | 1) & M128
    state = (0 * PCG_MULT + inc) & M128
    state = (state + initstate) & M128
    state = (state * PCG_MULT + inc) & M128
    out = np.empty(n)
    for i in range(n):
        state = (state * PCG_MULT + inc) & M128
        hi, lo = state >> 64, state & M64
        x, r = hi ^ lo, hi >> 58
        v = ((x >> r) | (x << ((64 - r) & 63))) & M64
        out[i] = (v >> 11) * (1.0 / 9007199254740992.0)
    return out


def test_a_million_device_seeded_mice(cuda):
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    import bench
    rng = np.random.default_rng(404)
    maze, omz = H.product_m9(), H.oracle_m9()
    P = bench.sim_population(rng, N)
    V = rng.random((9, 9))
    acts = H.a8()
    pos0 = np.array(maze.get_tile_centre(1, 1)) + rng.uniform(-0.03, 0.03, (N, 2))
    ori0 = rng.uniform(-3, 3, N)
    exp = BatchedExperiment(maze, acts, O.A8_UNITS, [bc.AnxietyComponent, bc.BiomechanicalCostComponent,
                                                     bc.ExperienceComponent, bc.BiomechPersistenceComponent,
                                                     bc.ValueTableComponent], v_table=V)
    seed = 20261020
    res = exp.run(P, T, pos0, ori0, init_disc=(1, 1), device_seed=seed, n_mice=N, occupancy=True, occupancy_total=True)
    assert not res.status.any()
    # ---- the whole population
    occ = res.occupancy.reshape(N, -1)
    assert np.array_equal(occ.sum(axis=1), np.full(N, T + 1))                    # every mouse: n_steps + 1 positions
    assert np.array_equal(res.tile_counts.sum(axis=1), np.full(N, T + 1))
    assert np.array_equal(res.occupancy_total.reshape(-1), occ.sum(axis=0, dtype=np.int64))
    assert not occ[:, [i for i in range(81) if (i // 9, i % 9) not in set(map(tuple, omz.vis_cells))]].any()
    assert np.isfinite(res.final_pose).all() and (res.count_moving <= T).all() and (res.it_visit <= T + 1).all()
    again = exp.run(P, T, pos0, ori0, init_disc=(1, 1), device_seed=seed, n_mice=N)
    assert np.array_equal(again.final_pose, res.final_pose) and np.array_equal(again.pcg_state, res.pcg_state)
    # a mouse does not depend on its place in the launch: the last 4 096 mice alone, as mice N - 4096 ... N - 1
    lo = N - 4096
    tail = exp.run({k: v[lo:] for k, v in P.items()}, T, pos0[lo:], ori0[lo:], init_disc=(1, 1), device_seed=seed,
                   n_mice=4096, mouse_index_base=lo)
    assert np.array_equal(tail.final_pose, res.final_pose[lo:]) and np.array_equal(tail.tile_counts, res.tile_counts[lo:])
    # ---- a sample against the oracle, through the host restatement of the device's seeding
    sample = np.concatenate([[0, 1, 31, 32, 127, 128, N // 2, N - 1], rng.integers(0, N, 40)])
    draws = np.stack([device_stream(seed, int(m), T) for m in sample])
    Ps = {k: v[sample] for k, v in P.items()}
    pose = np.concatenate([pos0[sample], ori0[sample, None]], axis=1)
    cells = np.tile([1, 1], (len(sample), 1))
    tabs = np.stack([O.bcost_table(O.A8_UNITS, Ps["k"][i], Ps["gamma"][i]) for i in range(len(sample))])
    ref = CO.simulate(omz, acts, O.A8_UNITS, KINDS5, Ps, pose, cells, draws, vtable_norm=O.normalise_vtable(V),
                      bcost_tables=tabs)
    assert not ref["status"].any()
    assert np.array_equal(ref["final_pose"], res.final_pose[sample])             # bit for bit
    assert np.array_equal(ref["occupancy"].reshape(len(sample), -1), occ[sample].astype(np.int32))
    st = maze.size_tile
    moved = (np.diff(np.floor_divide(ref["hist_pos"], st), axis=1) != 0).any(axis=2).sum(axis=1)
    assert np.array_equal(moved, res.count_moving[sample])
