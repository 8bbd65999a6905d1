"""A population of simulated mice sharded over the ranks of a process group (contiguous blocks of
the flat mouse index, no data-path collective, ONE small all-reduce of the population totals).

CPU: world_size-2 ``gloo`` run of ``BatchedExperiment.run_sharded`` with the launch replaced by a stub that records
what every rank was asked to simulate (the slicing, the global stream index, the all-reduce are host logic).
GPU: the shards of a device-seeded population, simulated one after the other on one device, are the same mice as the
unsharded run, and their packed totals add up to its totals."""
import os
import sys

import numpy as np
import pytest

import helpers as H
from oracle import np_oracle as O


def _stub_result(N, T, lo):
    """A SimResult whose numbers depend on the GLOBAL mouse index only."""
    from navigation_model_b200.batched_sim import SimResult
    g = np.arange(lo, lo + N)
    res = SimResult(final_pose=np.zeros((N, 3)), status=(g % 7 == 3).astype(np.int32), it_visit=np.full(N, T, np.int32),
                    count_moving=(g % 11).astype(np.int32),
                    tile_counts=np.stack([(g + c) % 5 for c in range(8)], axis=1).astype(np.int32))
    res.n_positions = T + 1
    res.zone_occupancy, res.zone_entries = (g % 3).astype(np.int32), (g % 2).astype(np.int32)
    res.corridor_switches, res.static_intervals = (g % 4).astype(np.int32), (g % 6).astype(np.int32)
    res.subarea_counts = np.stack([(g + c) % 3 for c in range(16)], axis=1).astype(np.int32)
    res.orientations_histogram = np.stack([(g + c) % 4 for c in range(6)], axis=1).astype(np.int32)
    res.occupancy_total = np.bincount(g % 81, minlength=81).reshape(9, 9).astype(np.int64)
    return res


def _sharded_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch.distributed as dist
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    maze = H.product_m9()
    N, T = 1000, 20
    calls = []

    def fake_run(self, params, n_steps, init_pos, init_ori, init_disc, **kw):
        calls.append({"beta": np.asarray(params["beta"]).copy(), "W_anx": params["W_anx"], "pos": np.asarray(init_pos).copy(),
                      "ori": init_ori, "base": kw["mouse_index_base"], "n": kw["n_mice"], "seed": kw["device_seed"]})
        return _stub_result(kw["n_mice"], n_steps, kw["mouse_index_base"])

    BatchedExperiment.run = fake_run
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, [bc.AnxietyComponent])
    P = {"beta": np.linspace(1, 2, N), "W_anx": 2.0, "p1": 0.9, "p2": 0.5, "p3": 0.4}
    pos = np.tile(maze.get_tile_centre(1, 1), (N, 1)) + np.arange(N)[:, None] * 1e-6
    res, (lo, hi), tot = exp.run_sharded(P, T, pos, 0.25, (1, 1), device_seed=5, n_mice=N, orientation_histogram={})
    c = calls[0]
    ok = (np.array_equal(c["beta"], P["beta"][lo:hi]) and c["W_anx"] == 2.0 and np.array_equal(c["pos"], pos[lo:hi])
          and c["ori"] == 0.25 and c["base"] == lo and c["n"] == hi - lo and c["seed"] == 5)
    # more ranks than mice: the rank without a block still takes part in the all-reduce
    one, (lo1, hi1), tot1 = exp.run_sharded({**P, "beta": 1.5}, T, pos[0], 0.25, (1, 1), device_seed=5, n_mice=1)
    ok = ok and (one is None) == (rank == 1) and (lo1, hi1) == ((0, 1) if rank == 0 else (1, 1)) and tot1["n_mice"] == 1 \
        and tot1["n_positions"] == T + 1
    q.put((rank, lo, hi, ok, {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in tot.items()}))
    dist.destroy_process_group()


def test_run_sharded_world2_gloo():
    import torch.multiprocessing as mp
    from navigation_model_b200.batched_sim import pack_population_totals, shard_bounds_mice, unpack_population_totals
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(o[1], o[2]) for o in out] == [shard_bounds_mice(1000, r, 2) for r in range(2)] == [(0, 512), (512, 1000)]
    assert all(o[3] for o in out)
    whole = unpack_population_totals(pack_population_totals(_stub_result(1000, 20, 0), 81), (9, 9), 6)
    for o in out:  # every rank holds the totals of the whole population
        for k, v in whole.items():
            assert (v.tolist() if isinstance(v, np.ndarray) else v) == o[4][k], k
    assert whole["n_mice"] == 1000 and whole["n_positions"] == 21000 and whole["occupancy_total"].sum() == 1000


def test_pipeline_bounds_cover_the_population_in_whole_ctas():
    from navigation_model_b200.batched_sim import pipeline_bounds
    for N, chunks in ((1 << 20, 6), (5000, 4), (5000, 64), (700, 3), (100, 4), (128, 2), (1, 6), (1000000, 6)):
        b = pipeline_bounds(N, chunks)
        assert b[0][0] == 0 and b[-1][1] == N and all(hi > lo for lo, hi in b)
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert all(lo % 128 == 0 for lo, _ in b)          # every block starts on a CTA boundary
        if len(b) > 2 and b[1][1] - b[0][0] >= 4 * 128:
            assert b[0][1] - b[0][0] <= (b[1][1] - b[0][0]) // 4 + 128   # the short first block
    assert pipeline_bounds(100, 4) == [(0, 100)]


def test_shard_bounds_mice_cover_the_population():
    from navigation_model_b200.batched_sim import shard_bounds_mice
    for N in (0, 1, 127, 128, 129, 1000, 909312, 10 ** 6):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds_mice(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            if N >= 128 * world:  # whole CTAs of the simulation kernel, evenly
                assert all(lo % 128 == 0 for lo, _ in spans) and max(h - l for l, h in spans) - min(h - l for l, h in spans) < 256


@pytest.mark.gpu
def test_shards_are_the_same_mice(cuda):
    """Device-seeded streams follow the global mouse index: simulating [0, 300) and [300, 1000) separately gives the
    mice of the unsharded run bit for bit, and the packed totals of the shards add up to its totals; without a process
    group run_sharded is run + totals."""
    from navigation_model_b200.batched_sim import BatchedExperiment, pack_population_totals
    from navigation_model_b200.simulation import behavioural_components as bc
    maze = H.product_m9()
    rng = np.random.default_rng(5)
    N, T = 1000, 120
    exp = BatchedExperiment(maze, H.a8(), O.A8_UNITS, [bc.AnxietyComponent, bc.ExperienceComponent])
    P = {"beta": rng.uniform(0.5, 4, N), "W_anx": rng.uniform(0, 5, N), "p1": 0.9, "p2": 0.5, "p3": 0.4,
         "W_exp": rng.uniform(0, 5, N), "xi": 0.1, "kappa_home": 1.0, "kappa_explo": 1.0, "theta_explo": 100.0,
         "theta_home": 50.0}
    pos = np.array(maze.get_tile_centre(1, 1)) + rng.uniform(-0.02, 0.02, (N, 2))
    kw = dict(init_ori=0.0, init_disc=(1, 1), history=True, occupancy_total=True, subareas=True, static_intervals=True,
              shock_zone={"tile_lines": 5}, corridors=(1, 7), orientation_histogram={})
    whole, (lo, hi), tot = exp.run_sharded(P, T, pos, device_seed=11, n_mice=N, **kw)
    assert (lo, hi) == (0, N) and tot["n_mice"] == N and tot["n_positions"] == N * (T + 1)
    assert np.array_equal(tot["occupancy_total"], whole.occupancy_total) and tot["occupancy_total"].sum() == N * (T + 1)
    assert np.array_equal(tot["tile_counts"], whole.tile_counts.sum(axis=0))
    vec = np.zeros_like(pack_population_totals(whole, 81))
    for a, b in ((0, 300), (300, N)):
        part = exp.run({k: (v[a:b] if isinstance(v, np.ndarray) else v) for k, v in P.items()}, T, pos[a:b],
                       device_seed=11, n_mice=b - a, mouse_index_base=a, **kw)
        assert np.array_equal(part.choices, whole.choices[a:b])
        assert np.array_equal(part.history_position, whole.history_position[a:b])
        assert np.array_equal(part.pcg_state, whole.pcg_state[a:b])
        vec += pack_population_totals(part, 81)
    assert np.array_equal(vec, pack_population_totals(whole, 81))
    other = exp.run({k: (v[:300] if isinstance(v, np.ndarray) else v) for k, v in P.items()}, T, pos[:300],
                    device_seed=11, n_mice=300, mouse_index_base=300, **kw)
    assert not np.array_equal(other.choices, whole.choices[:300])  # another block of the population: other streams
