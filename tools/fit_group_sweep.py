#!/usr/bin/env python
"""Time of one wave of the warp-shared-table fit kernel as a function of its consumer warps per CTA
(NM_FIT_GROUP_NW): one CTA per SM, 148 * nw * 32 parameter sets.

    python tools/fit_group_sweep.py            # prints nw, kernel ms, parameter sets / (SM ms)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from navigation_model_b200.sweep import ConditioningSweep
    from navigation_model_b200 import workloads as W
    maze = W.maze_m9()
    acts = W.a8()
    traj, rew = W.synthetic_session(maze, 0, 5000)
    sw = ConditioningSweep(maze, "positive", possible_actions=acts, device=0)
    sw.add_session(traj, rew)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    mult = int(os.environ.get("SWEEP_CTAS_PER_SM", "1"))  # co-resident CTAs per SM (small CTAs only)
    for nw in [int(x) for x in (sys.argv[1:] or range(28, 0, -4))]:
        os.environ["NM_FIT_GROUP_NW"] = str(nw)
        runs = 148 * nw * mult  # uniform runs of 32 parameter sets, one per consumer warp
        P = runs * 32
        a = torch.from_numpy(np.repeat(rng.uniform(0.01, 1, runs), 32)).to(dev)
        g = torch.from_numpy(np.repeat(rng.uniform(0.5, 0.99, runs), 32)).to(dev)
        b = torch.from_numpy(rng.uniform(0.1, 30, P)).to(dev)
        ll = torch.empty((1, P), dtype=torch.float64, device=dev)
        best = 1e9
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sw.log_likelihood_device(a, b, g, ll_out=ll)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"nw": nw, "kernel_ms": best, "ctas_per_sm": mult, "sets_per_sm_ms": P / 148 / best,
                          "updates_per_s": P * 4999 / (best * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
