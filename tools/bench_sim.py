#!/usr/bin/env python
"""Throughput of the forward-simulation kernel (BASELINE.json configs[4], scaled): N mice x T steps on the
9x9 maze, 8 actions, all five components, parameters uniform inside the descriptor bounds, device PCG64.

    python tools/bench_sim.py [--mice 262144] [--steps 1000] [--reps 3] [--cpu-mice 8]

Prints one JSON line: agent-steps/s on the GPU (CUDA events around the launch, inputs resident) and of
the numpy oracle on one host core for a small sample."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def params(rng, N):
    u = rng.uniform
    return {"beta": u(0.5, 6, N), "W_anx": u(0, 10, N), "p1": u(0, 1, N), "p2": u(0, 1, N), "p3": u(0, 1, N),
            "W_bcost": u(0, 10, N), "k": u(0.05, 4, N), "gamma": u(0.2, 2, N), "W_exp": u(0, 10, N),
            "xi": u(0.006, 0.6, N), "kappa_home": u(0.01, 5, N), "kappa_explo": u(0.01, 5, N),
            "theta_explo": u(80, 120, N), "theta_home": u(30, 70, N), "W_bper": u(0, 10, N), "bp": u(0, 1, N),
            "Wm": u(0, 2, N), "Wnm": u(0, 2, N), "W_values": u(0, 10, N)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mice", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-mice", type=int, default=4)
    ap.add_argument("--pipelined", type=int, default=0, metavar="CHUNKS", help="run_pipelined with this many blocks")
    ap.add_argument("--workers", type=int, default=2)
    args = ap.parse_args()
    import torch
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    from navigation_model_b200 import workloads as W
    from oracle import np_oracle as O  # the CPU sample at the end only

    maze = W.maze_m9()
    acts = W.a8()
    rng = np.random.default_rng(0)
    N, T = args.mice, args.steps
    V = rng.random((9, 9))
    P = params(rng, N)
    exp = BatchedExperiment(maze, acts, W.A8_UNITS, [bc.AnxietyComponent, bc.BiomechanicalCostComponent,
                                                     bc.ExperienceComponent, bc.BiomechPersistenceComponent,
                                                     bc.ValueTableComponent], v_table=V, device=0)
    pos0 = np.array(maze.get_tile_centre(1, 1)) + rng.uniform(-0.03, 0.03, (N, 2))
    ori0 = rng.uniform(-3, 3, N)
    times, kernel_ms = [], []
    res = None
    for rep in range(args.reps + 1):
        t0 = time.perf_counter()
        run = (lambda *a, **k: exp.run_pipelined(*a, chunks=args.pipelined, workers=args.workers, **k)) if args.pipelined else exp.run
        res = run(P, T, pos0, ori0, init_disc=(1, 1), device_seed=1234 + rep, n_mice=N, occupancy_total=True)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        kernel_ms.append(exp.last_kernel_ms)
    best = min(times[1:])
    ok = int((res.status == 0).sum())
    # CPU oracle sample
    omz = O.OracleMaze(O.M9_LAYOUT, O.M9_SIZE_TILE)
    t0 = time.perf_counter()
    Tc = min(T, 300)
    for m in range(args.cpu_mice):
        comps = [(O.COMP_ANXIETY, {k: P[k][m] for k in ("W_anx", "p1", "p2", "p3")}),
                 (O.COMP_BCOST, {"W_bcost": P["W_bcost"][m], "table": O.bcost_table(O.A8_UNITS, P["k"][m], P["gamma"][m])}),
                 (O.COMP_EXPERIENCE, {k: P[k][m] for k in ("W_exp", "xi", "kappa_home", "kappa_explo", "theta_explo",
                                                           "theta_home")}),
                 (O.COMP_BPERSISTENCE, {k: P[k][m] for k in ("W_bper", "bp", "Wm", "Wnm")}),
                 (O.COMP_VTABLE, {"W_values": P["W_values"][m], "table": O.normalise_vtable(V)})]
        O.simulate(omz, acts, O.A8_UNITS, pos0[m], ori0[m], (1, 1), P["beta"][m], comps,
                   np.random.default_rng(m).random(Tc))
    cpu_rate = args.cpu_mice * Tc / (time.perf_counter() - t0)
    print(json.dumps({"metric": "agent-steps/s (forward simulation, 5 components, 9x9 maze, 8 actions)",
                      "value": N * T / (min(kernel_ms[1:]) * 1e-3), "kernel_ms_best": min(kernel_ms[1:]),
                      "e2e_value_incl_host_prep_h2d_d2h": N * T / best, "mice": N, "steps": T,
                      "seconds_best_incl_h2d_d2h": best,
                      "all_reps_s": times, "mice_ok": ok, "moving_mean": float(res.count_moving.mean()),
                      "it_visit_80_median": float(np.median(res.it_visit)),
                      "cpu_numpy_oracle_1core_steps_per_s": cpu_rate}))


if __name__ == "__main__":
    main()
