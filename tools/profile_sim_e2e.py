"""Where the end-to-end time of a batched simulation goes on the host (cProfile of BatchedExperiment.run on the
bench.py workload: 909 312 mice x 1000 steps, all fused metrics).  `python tools/profile_sim_e2e.py [--mice N]`."""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mice", type=int, default=148 * 6 * 128 * 8)
    ap.add_argument("--steps", type=int, default=1000)
    a = ap.parse_args()
    import torch
    from navigation_model_b200 import workloads as W
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc
    maze = W.maze_m9()
    rng = np.random.default_rng(1000)
    N, T = a.mice, a.steps
    u = rng.uniform
    P = {"beta": u(0.5, 6, N), "W_anx": u(0, 10, N), "p1": u(0, 1, N), "p2": u(0, 1, N), "p3": u(0, 1, N),
         "W_bcost": u(0, 10, N), "k": u(0.05, 4, N), "gamma": u(0.2, 2, N), "W_exp": u(0, 10, N),
         "xi": u(0.006, 0.6, N), "kappa_home": u(0.01, 5, N), "kappa_explo": u(0.01, 5, N),
         "theta_explo": u(80, 120, N), "theta_home": u(30, 70, N), "W_bper": u(0, 10, N), "bp": u(0, 1, N),
         "Wm": u(0, 2, N), "Wnm": u(0, 2, N), "W_values": u(0, 10, N)}
    exp = BatchedExperiment(maze, W.a8(), W.A8_UNITS,
                            [bc.AnxietyComponent, bc.BiomechanicalCostComponent, bc.ExperienceComponent,
                             bc.BiomechPersistenceComponent, bc.ValueTableComponent], v_table=rng.random((9, 9)))
    pos0 = np.array(maze.get_tile_centre(1, 1)) + rng.uniform(-0.03, 0.03, (N, 2))
    ori0 = rng.uniform(-3, 3, N)

    def run(rep):
        return exp.run(P, T, pos0, ori0, init_disc=(1, 1), device_seed=77 + rep, n_mice=N, occupancy_total=True,
                       shock_zone={"tile_lines": 5}, corridors=(1, 7), static_intervals=True, subareas=True,
                       orientation_histogram={})
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(rep)
        print(f"rep {rep}: e2e {time.perf_counter() - t0:.4f} s, kernel {exp.last_kernel_ms:.2f} ms")
    pr = cProfile.Profile()
    pr.enable()
    run(3)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print(s.getvalue())


if __name__ == "__main__":
    main()
