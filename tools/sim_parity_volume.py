#!/usr/bin/env python
"""Simulation parity at volume: N mice x T steps with replayed draws through the CUDA kernel and through the compiled
oracle (oracle/nm_oracle.c:orc_simulate, OpenMP over the mice), compared mouse by mouse WITHOUT histories -- final pose
(bit for bit: a pose is a chain, one differing step would show), visit map, motion count, status.

    python tools/sim_parity_volume.py [--mice 131072] [--steps 1000] [--seed 1]     # 1.3e8 agent-steps, ~1 GB of draws

Prints one JSON line (kept as profiles/r02_sim_parity_volume.json)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mice", type=int, default=131072)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    import bench
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    from navigation_model_b200 import workloads as W
    from oracle import c_oracle as CO
    from oracle import np_oracle as O
    N, T = args.mice, args.steps
    rng = np.random.default_rng(args.seed)
    maze, omz = W.maze_m9(), O.OracleMaze(O.M9_LAYOUT, O.M9_SIZE_TILE)
    P = bench.sim_population(rng, N)
    P["beta"][: N // 2] = rng.uniform(0.02, 0.5, N // 2)  # half the population wanders widely
    V = rng.random((9, 9)) * 2 - 0.5
    acts = W.a8()
    cells = np.array([omz.vis_cells[i] for i in rng.integers(0, len(omz.vis_cells), N)])
    pos0 = np.array([maze.get_tile_centre(c) for c in cells]) + rng.uniform(-0.04, 0.04, (N, 2))
    ori0 = rng.uniform(-np.pi, np.pi, N)
    draws = rng.random((N, T))
    exp = BatchedExperiment(maze, acts, O.A8_UNITS, [bc.AnxietyComponent, bc.BiomechanicalCostComponent,
                                                     bc.ExperienceComponent, bc.BiomechPersistenceComponent,
                                                     bc.ValueTableComponent], v_table=V, device=0)
    res = exp.run(P, T, pos0, ori0, init_disc=cells, draws=draws, occupancy=True)
    tabs = bc.bcost_tables(O.A8_UNITS, P["k"], P["gamma"])  # the product's vectorised von-Mises tables ...
    for m in (0, N // 3, N - 1):                            # ... equal the oracle's per-mouse form
        assert np.array_equal(tabs[m], O.bcost_table(O.A8_UNITS, P["k"][m], P["gamma"][m]))
    pose = np.concatenate([pos0, ori0[:, None]], axis=1)
    kinds = [O.COMP_ANXIETY, O.COMP_BCOST, O.COMP_EXPERIENCE, O.COMP_BPERSISTENCE, O.COMP_VTABLE]
    t0 = time.perf_counter()
    ref = CO.simulate(omz, acts, O.A8_UNITS, kinds, P, pose, cells, draws, vtable_norm=O.normalise_vtable(V),
                      bcost_tables=tabs, history=False)
    cpu_s = time.perf_counter() - t0
    same_pose = (res.final_pose == ref["final_pose"]).all(axis=1)
    same_occ = (res.occupancy.reshape(N, -1) == ref["occupancy"].reshape(N, -1)).all(axis=1)
    moved_tiles = int((res.occupancy.reshape(N, -1) > 0).sum())
    out = {"mice": N, "steps": T, "agent_steps": N * T, "kernel_ms": exp.last_kernel_ms, "oracle_seconds": cpu_s,
           "mice_with_identical_final_pose": int(same_pose.sum()), "mice_with_identical_visit_map": int(same_occ.sum()),
           "mice_status_ok_gpu": int((res.status == 0).sum()), "mice_status_ok_oracle": int((ref["status"] == 0).sum()),
           "count_moving_total_gpu": int(res.count_moving.sum()), "distinct_tiles_visited_total": moved_tiles,
           "all_identical": bool(same_pose.all() and same_occ.all() and np.array_equal(res.status, ref["status"]))}
    print(json.dumps(out))
    return 0 if out["all_identical"] else 1


if __name__ == "__main__":
    sys.exit(main())
