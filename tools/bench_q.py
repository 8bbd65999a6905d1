#!/usr/bin/env python
"""Tabular Q-learning agent (explicit Q[s, a] table; EXTENSION) on the grid of BASELINE.json configs[1]: 64 x 64 x 32
(alpha, beta, gamma) parameter sets over one synthetic 5000-step session on the 9 x 9 maze, 8 tile actions.  Prints one
JSON line (agent-trial updates/s; one update = one observed step of one parameter set: softmax log-probability + one
Q update)."""
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
    ap.add_argument("--grid", type=int, nargs=3, default=[64, 64, 32])
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--maze", choices=["m9", "m20"], default="m9")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-params", type=int, default=2)
    args = ap.parse_args()
    import torch
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.sweep import QLearningSweep, parameter_grid
    from navigation_model_b200 import workloads as W
    from oracle import np_oracle as O  # spot check of the result only
    acts = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    if args.maze == "m9":
        maze = Maze(O.M9_LAYOUT, O.M9_SUBAREAS, size_tile=O.M9_SIZE_TILE)
        omz = O.OracleMaze(O.M9_LAYOUT, O.M9_SIZE_TILE)
        traj, rew = W.synthetic_session(maze, 0, args.steps)
    else:
        layout = "\n".join(["M" * 20] * 20)
        maze = Maze(layout, "", size_tile=0.05)
        omz = O.OracleMaze(layout, 0.05)
        traj, rew = W.synthetic_session(maze, 0, args.steps, start=(3, 3), goal=(15, 12), jitter=0.02)
    na, nb, ng = args.grid
    # beta fastest: runs of 32 consecutive parameter sets share (alpha, gamma) and with it their table
    g, a, b = parameter_grid(np.linspace(0.5, 0.99, ng), np.linspace(0.01, 1, na), np.logspace(-1, 1.5, nb))
    P = a.size
    ql = QLearningSweep(maze, acts, device=0)
    ql.add_session(traj, rew)
    wall, kern = [], []
    for _ in range(args.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ll = ql.log_likelihood(a, b, g)
        wall.append(time.perf_counter() - t0)
        kern.append(ql.last_kernel_ms)
    tr = O.extract_transitions(omz, traj, rew, None)
    nxt = O.transition_table(omz, acts)
    t0 = time.perf_counter()
    chk = 0.0
    picks = np.linspace(0, P - 1, args.cpu_params).astype(int)
    for p in picks:
        _, llw = O.run_q_learning(tr, nxt, a[p], b[p], g[p])
        chk = max(chk, abs(llw - ll[0, p]) / abs(llw))
    cpu = len(picks) * (args.steps - 1) / (time.perf_counter() - t0)
    upd = P * (args.steps - 1)
    k_ms = min(kern[1:])
    print(json.dumps({"metric": "agent-trial updates/s (Q-table agent, %s maze, 8 actions)" % args.maze,
                      "value": upd / (k_ms * 1e-3), "kernel_ms": k_ms, "sweep_mode": ql.last_sweep_mode(), "params": P, "steps": args.steps,
                      "seconds_best_incl_h2d_d2h": min(wall[1:]), "oracle_max_rel_err": chk,
                      "cpu_numpy_oracle_1core_updates_per_s": cpu}))


if __name__ == "__main__":
    main()
