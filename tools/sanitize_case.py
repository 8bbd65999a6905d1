#!/usr/bin/env python
"""Smallest end-to-end invocation of every kernel, for compute-sanitizer (one tool per call):

    compute-sanitizer --tool memcheck python tools/sanitize_case.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from navigation_model_b200.batched_sim import BatchedExperiment, movements_possible  # noqa: E402
from navigation_model_b200.simulation import behavioural_components as bc  # noqa: E402
from navigation_model_b200.sweep import ConditioningSweep, ModelBasedSweep  # noqa: E402
from navigation_model_b200 import workloads as W  # noqa: E402

maze = W.maze_m9()
acts = W.a8()
rng = np.random.default_rng(0)
P = 70
a, b, g = rng.uniform(0.01, 1, P), rng.uniform(0.1, 30, P), rng.uniform(0.5, 0.99, P)
for alg in ("td0", "positive", "negative"):
    sw = ConditioningSweep(maze, alg, possible_actions=acts, replay_n_times=1, replay_seed=3, device=0)
    for sid, T in ((0, 130), (1, 33)):
        sw.add_session(*W.synthetic_session(maze, sid, T))
    ll, V = sw.log_likelihood(a, b, g, return_v=True)
    os.environ["NM_FIT_MEMO"] = "1"
    ll2 = sw.log_likelihood(a, b, g)
    os.environ["NM_FIT_MEMO"] = "0"
    assert np.allclose(ll, ll2, rtol=1e-9)
    sw.v_tables(a, g)
    print(alg, "best", sw.best_fit(a, b, g))
mb = ModelBasedSweep(maze, [(1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1), (1, -1), (-1, 1)], n_sweeps=2, device=0)
mb.add_session(*W.synthetic_session(maze, 2, 90))
print("mb", mb.log_likelihood(a[:9], b[:9], g[:9]).sum())
N, T = 40, 60
u = rng.uniform
params = {"beta": u(0.5, 4, N), "W_anx": u(0, 5, N), "p1": u(0, 1, N), "p2": u(0, 1, N), "p3": u(0, 1, N),
          "W_bcost": u(0, 5, N), "k": u(0.1, 3, N), "gamma": u(0.3, 2, N), "W_exp": u(0, 5, N), "xi": u(0.006, 0.6, N),
          "kappa_home": u(0.01, 5, N), "kappa_explo": u(0.01, 5, N), "theta_explo": u(80, 120, N),
          "theta_home": u(30, 70, N), "W_bper": u(0, 5, N), "bp": u(0, 1, N), "Wm": u(0, 2, N), "Wnm": u(0, 2, N),
          "W_values": u(0, 5, N)}
exp = BatchedExperiment(maze, acts, W.A8_UNITS, [bc.AnxietyComponent, bc.BiomechanicalCostComponent,
                                                 bc.ExperienceComponent, bc.BiomechPersistenceComponent,
                                                 bc.ValueTableComponent], v_table=rng.random((9, 9)), device=0)
res = exp.run(params, T, maze.get_tile_centre(1, 1), 0.2, init_disc=(1, 1), seeds=list(range(N)), history=True,
              occupancy=True, occupancy_total=True)
print("sim", res.count_moving.sum(), res.occupancy_total.sum())
rnd = BatchedExperiment(maze, acts, W.A8_UNITS, [], model="random", device=0)
print("rnd", rnd.run({}, T, maze.get_tile_centre(7, 4), 0.0, seeds=list(range(N))).count_moving.sum())
print("seg", movements_possible(maze, rng.uniform(0, 1, (500, 4))).sum())
print("sanitize case ok")
