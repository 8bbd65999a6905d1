"""Log-likelihoods of the BASELINE configs[1] grid (64 x 64 x 32 parameter sets, one 5 000-step session) saved to a
.npy file: `python tools/ll_dump.py out.npy` -- run once per build variant (NM_NATIVE_LIB=...) and compare the files
(`python tools/ll_dump.py --compare a.npy b.npy`) to see what a compile-time change does to the results."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    if sys.argv[1] == "--compare":
        a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
        rel = np.abs(a - b) / np.abs(b)
        print(f"max relative difference {rel.max():.3e}, mean {rel.mean():.3e}, identical {np.array_equal(a, b)}, "
              f"arg-min {int(np.argmin(-a))} / {int(np.argmin(-b))}")
        return
    from navigation_model_b200 import workloads as W
    from navigation_model_b200.sweep import ConditioningSweep, parameter_grid
    maze = W.maze_m9()
    traj, rew = W.synthetic_session(maze, 0, 5000)
    sw = ConditioningSweep(maze, "positive", possible_actions=W.a8())
    sw.add_session(traj, rew)
    g, a, b = parameter_grid(np.linspace(0.5, 0.99, 32), np.linspace(0.01, 1, 64), np.logspace(-1, 1.5, 64))
    np.save(sys.argv[1], sw.log_likelihood(a, b, g)[0])


if __name__ == "__main__":
    main()
