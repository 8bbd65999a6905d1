"""Golden for continued learning on ONE experiment object after shuffled replays, produced by the UNMODIFIED reference.

    python tests/golden/make_golden_replay_continue.py

``run(session A, do_replays=True)`` -> ``do_replays()`` -> ``run(session B)`` on the same ConditioningExperiment /
algorithm / mouse / ShuffledReplays objects.  Every replayed update of PositiveConditioning / NegativeConditioning ends
with ``mouse.move_to(t.cs1)`` (rl.py:189, 232), so after the replays the learner's mouse stands at the arrival point of
the last replayed transition with the heading of its last displacement; the first transition of session B computes its
candidate tiles from THAT pose (rl.py:177-182).  Stored: tables after each stage, the mouse pose after each stage, the
length of the mouse history, and the candidate lists of session B's first transitions.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refharness  # noqa: E402

if not refharness.install():
    raise SystemExit("reference not available at /root/reference")

from navigation_model.simulation.maze import Maze  # noqa: E402  (the reference)
from navigation_model.simulation.mouse import Mouse  # noqa: E402
from navigation_model.simulation.rl import (ConditioningExperiment, NegativeConditioning,  # noqa: E402
                                            PositiveConditioning, ShuffledReplays)

from make_golden import A8, M9_FILE, SIZE_TILE, synthetic_session  # noqa: E402


def main():
    maze = Maze.from_file(M9_FILE, size_tile=SIZE_TILE)
    trajA, rewA = synthetic_session(maze, 21, 300)
    trajB, rewB = synthetic_session(maze, 22, 200, start=(7, 7), goal=(1, 2))
    out = {"trajA": trajA, "rewA": rewA, "trajB": trajB, "rewB": rewB, "A8": A8, "alpha": np.array(0.3),
           "gamma": np.array(0.85), "n_times": np.array(2), "seed": np.array(5)}
    for name, cls, sign in (("positive", PositiveConditioning, 1.0), ("negative", NegativeConditioning, -1.0)):
        mouse = Mouse(trajA[0], 0.0, A8)
        alg = cls(0.3, 0.85, mouse, maze)
        rep = ShuffledReplays(n_times=2, seed=5)
        exp = ConditioningExperiment(maze, alg, rep)
        poses, tabs = [], []

        def snap():
            poses.append([mouse.position[0], mouse.position[1], mouse.orientation])
            tabs.append(np.array(exp.v_table))

        exp.run(list(trajA), [0.0] * len(trajA), list(sign * rewA), do_replays=True)
        snap()
        exp.do_replays()
        snap()
        exp.run(list(trajB), [0.0] * len(trajB), list(sign * rewB), do_replays=False)
        snap()
        out[f"{name}_pose"] = np.array(poses)
        out[f"{name}_V"] = np.array(tabs)
        out[f"{name}_hist_len"] = np.array(len(mouse.history[0]))
        out[f"{name}_hist_tail"] = np.array(mouse.history[0][-5:])
        firstB = [u.transition for u in rep._buffer][len(trajA) - 1:][:4]
        cand = np.full((4, 8, 2), -1, dtype=np.int64)
        for i, t in enumerate(firstB):
            for k, c in enumerate(t.extra["next_states"]):
                cand[i, k] = c
        out[f"{name}_firstB_cand"] = cand
    np.savez_compressed(os.path.join(HERE, "replay_continue_golden.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()})
    print(out["positive_pose"])


if __name__ == "__main__":
    main()
