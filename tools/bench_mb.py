#!/usr/bin/env python
"""BASELINE.json configs[3]: model-based (world-model + value-iteration) agent fitness on a 400-state maze (20 x 20,
all visitable), 8 tile actions, P parameter sets, one synthetic session of T steps, n_sweeps value-iteration sweeps
per observed step.  Prints one JSON line (agent-trial updates/s; one update = one observed step of one parameter set:
likelihood + reward-model update + n_sweeps sweeps over 400 states x 8 actions)."""
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
    ap.add_argument("--params", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--sweeps", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--cpu-params", type=int, default=2)
    ap.add_argument("--grid", type=int, nargs=3, default=None, metavar=("N_ALPHA", "N_GAMMA", "N_BETA"),
                    help="a Cartesian grid with beta fastest instead of --params random parameter sets")
    args = ap.parse_args()
    import torch
    from navigation_model_b200.simulation.maze import Maze
    from navigation_model_b200.sweep import ModelBasedSweep
    from navigation_model_b200 import workloads as W
    from oracle import np_oracle as O  # spot check of the result only
    acts = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    layout = "\n".join(["M" * 20] * 20)
    maze = Maze(layout, "", size_tile=0.05)
    omz = O.OracleMaze(layout, 0.05)
    traj, rew = W.synthetic_session(maze, 0, args.steps, start=(3, 3), goal=(15, 12), jitter=0.02)
    rng = np.random.default_rng(0)
    P = args.params
    a, b, g = rng.uniform(0.01, 1, P), np.exp(rng.uniform(np.log(0.1), np.log(30), P)), rng.uniform(0.5, 0.99, P)
    if args.grid:
        from navigation_model_b200.sweep import parameter_grid
        na, ng, nb = args.grid
        a, g, b = parameter_grid(np.linspace(0.01, 1, na), np.linspace(0.5, 0.99, ng), np.logspace(-1, 1.5, nb))
        P = a.size
    mb = ModelBasedSweep(maze, acts, n_sweeps=args.sweeps, device=0)
    mb.add_session(traj, rew)
    times = []
    for _ in range(args.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ll = mb.log_likelihood(a, b, g)
        times.append(time.perf_counter() - t0)
    best = min(times[1:])
    tr = O.extract_transitions(omz, traj, rew, None)
    nxt = O.transition_table(omz, acts)
    t0 = time.perf_counter()
    chk = 0.0
    for p in range(args.cpu_params):
        _, _, llw = O.run_model_based(tr, nxt, a[p], b[p], g[p], n_sweeps=args.sweeps)
        chk = max(chk, abs(llw - ll[0, p]) / abs(llw))
    cpu = args.cpu_params * (args.steps - 1) / (time.perf_counter() - t0)
    upd = P * (args.steps - 1)
    print(json.dumps({"metric": "agent-trial updates/s (model-based agent, 400-state maze, 8 actions)",
                      "value": upd / best, "params": P, "steps": args.steps, "n_sweeps": args.sweeps, "sweep_mode": mb.last_sweep_mode(),
                      "seconds_best_incl_h2d_d2h": best, "all_reps_s": times,
                      "state_action_backups_per_s": upd * args.sweeps * 400 * 8 / best,
                      "oracle_max_rel_err": chk, "cpu_numpy_oracle_1core_updates_per_s": cpu}))


if __name__ == "__main__":
    main()
