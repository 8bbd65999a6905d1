"""A device-seeded population of mice sharded over the GPUs of one box (`BatchedExperiment.run_sharded`): one process per
GPU, contiguous blocks of the flat mouse index, ONE NCCL all-reduce of the population totals.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/run_sim_sharded.py --mice 400000 --steps 500 [--check]

`--check` (rank 0): simulates the whole population on its own GPU as well and compares the all-reduced totals with the
totals of that run (exact) -- the population must not depend on the number of ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mice", type=int, default=400000)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from navigation_model_b200 import workloads as W
    from navigation_model_b200.batched_sim import BatchedExperiment, pack_population_totals
    from navigation_model_b200.simulation import behavioural_components as bc
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    maze = W.maze_m9()
    rng = np.random.default_rng(4)  # the same population on every rank; run_sharded takes this rank's block
    N, T = a.mice, a.steps
    u = rng.uniform
    P = {"beta": u(0.5, 6, N), "W_anx": u(0, 10, N), "p1": u(0, 1, N), "p2": u(0, 1, N), "p3": u(0, 1, N),
         "W_bcost": u(0, 10, N), "k": u(0.05, 4, N), "gamma": u(0.2, 2, N), "W_exp": u(0, 10, N),
         "xi": u(0.006, 0.6, N), "kappa_home": u(0.01, 5, N), "kappa_explo": u(0.01, 5, N),
         "theta_explo": u(80, 120, N), "theta_home": u(30, 70, N), "W_bper": u(0, 10, N), "bp": u(0, 1, N),
         "Wm": u(0, 2, N), "Wnm": u(0, 2, N), "W_values": u(0, 10, N)}
    exp = BatchedExperiment(maze, W.a8(), W.A8_UNITS,
                            [bc.AnxietyComponent, bc.BiomechanicalCostComponent, bc.ExperienceComponent,
                             bc.BiomechPersistenceComponent, bc.ValueTableComponent], v_table=rng.random((9, 9)),
                            device=local)
    pos0 = np.array(maze.get_tile_centre(1, 1)) + rng.uniform(-0.03, 0.03, (N, 2))
    ori0 = rng.uniform(-3, 3, N)
    kw = dict(init_disc=(1, 1), device_seed=21, n_mice=N, occupancy_total=True, shock_zone={"tile_lines": 5},
              corridors=(1, 7), static_intervals=True, subareas=True, orientation_histogram={})
    times = []
    for rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, (lo, hi), tot = exp.run_sharded(P, T, pos0, ori0, **kw)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt.item()))
    out = {"n_gpus": world, "mice": N, "steps": T, "block_rank0": [lo, hi], "seconds_e2e_max_over_ranks": times,
           "agent_steps_per_s_e2e": N * T / min(times[1:]), "kernel_ms_rank0": exp.last_kernel_ms,
           "totals": {k: (int(v) if not isinstance(v, np.ndarray) else int(v.sum())) for k, v in tot.items()}}
    if a.check and rank == 0:
        whole = exp.run(P, T, pos0, ori0, **kw)
        ref = pack_population_totals(whole, 81)
        got = np.concatenate([[tot[k] for k in ("n_mice", "mice_ok", "n_positions", "count_moving", "zone_occupancy",
                                                "zone_entries", "corridor_switches", "static_intervals")],
                              tot["tile_counts"], tot["subarea_counts"],
                              np.pad(tot["orientations_histogram"], (0, 17 - len(tot["orientations_histogram"]))),
                              tot["occupancy_total"].reshape(-1)])
        out["totals_equal_unsharded_run"] = bool(np.array_equal(ref, got))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
