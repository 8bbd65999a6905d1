"""Golden replay-buffer contents produced by the UNMODIFIED reference:

    python tests/golden/make_golden_updates.py

ConditioningExperiment.run appends what the learner's update returns -- Update(transition, dv, |rpe|) -- for every
online step (rl.py:97-98, 190-194, 233-237, 260-262).  Stored for TD0 / PositiveConditioning / NegativeConditioning on a
300-step session, with two shuffled replay passes after the run (the buffer keeps the ONLINE values; the final table
includes the replays) and starting from a given table -> updates_golden.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import refharness  # noqa: E402

if not refharness.install():
    raise SystemExit("reference not available at /root/reference")

from navigation_model.simulation.maze import Maze  # noqa: E402
from navigation_model.simulation.mouse import Mouse  # noqa: E402
from navigation_model.simulation.rl import (ConditioningExperiment, NegativeConditioning,  # noqa: E402
                                            PositiveConditioning, ShuffledReplays, TD0)

from make_golden import A8, M9_FILE, SIZE_TILE, synthetic_session  # noqa: E402


def main():
    maze = Maze.from_file(M9_FILE, size_tile=SIZE_TILE)
    traj, rew = synthetic_session(maze, 2, 300)  # reaches the goal 12 times
    v0 = np.random.default_rng(3).random(maze.shape) * 0.3
    out = {"traj": traj, "rew": rew, "A8": A8, "v0": v0, "alpha": np.array(0.35), "gamma": np.array(0.9)}
    for name in ("td0", "positive", "negative"):
        r = -rew if name == "negative" else rew
        mouse = Mouse(traj[0], 0.0, A8)
        alg = {"td0": lambda: TD0(0.35, 0.9), "positive": lambda: PositiveConditioning(0.35, 0.9, mouse, maze),
               "negative": lambda: NegativeConditioning(0.35, 0.9, mouse, maze)}[name]()
        rep = ShuffledReplays(n_times=2, seed=9)
        exp = ConditioningExperiment(maze, alg, rep, v_table=v0.copy())
        # capture the buffer BEFORE the replays shuffle it: run without, copy, then replay
        exp.run(list(traj), [0.0] * len(traj), list(r), do_replays=False)
        out[f"{name}_dv"] = np.array([float(u.dv) for u in rep._buffer])
        out[f"{name}_rpe"] = np.array([float(u.rpe) for u in rep._buffer])
        out[f"{name}_V_online"] = np.array(exp.v_table)
        exp.do_replays()
        out[f"{name}_V_replayed"] = np.array(exp.v_table)
    np.savez_compressed(os.path.join(HERE, "updates_golden.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()}, float(np.abs(out["positive_dv"]).max()))


if __name__ == "__main__":
    main()
