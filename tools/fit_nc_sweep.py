#!/usr/bin/env python
"""Time of one CTA of the fit kernel as a function of its consumer-thread count (NM_FIT_NC): the table behind
`plan_launch`'s cost model in csrc/nm_fit.cu.

    python tools/fit_nc_sweep.py            # prints nc, waves, kernel ms, ms per wave, parameter sets / (SM ms)"""
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
    out = []
    for nc in range(448, 31, -32):
        os.environ["NM_FIT_NC"] = str(nc)
        P = 148 * nc * 3  # exactly three full waves at one CTA per SM
        a, b, g = (torch.from_numpy(rng.uniform(lo, hi, P)).to(dev) for lo, hi in ((0.01, 1), (0.1, 30), (0.5, 0.99)))
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
        out.append({"nc": nc, "kernel_ms": best, "updates_per_s": P * 4999 / (best * 1e-3)})
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
