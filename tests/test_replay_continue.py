"""Continued learning on one experiment object after shuffled replays: the replay passes of the look-ahead
learners move the learner's mouse (rl.py:189, 232), and the next run() starts from that pose.  Golden values from the
unmodified reference (tests/golden/make_golden_replay_continue.py)."""
import os

import numpy as np
import pytest

import helpers as H
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "replay_continue_golden.npz"), allow_pickle=False))


def _stages(g, name, cls_of):
    from navigation_model_b200.simulation.mouse import Mouse
    from navigation_model_b200.simulation.rl import ConditioningExperiment, ShuffledReplays
    maze = H.product_m9()
    sign = -1.0 if name == "negative" else 1.0
    mouse = Mouse(g["trajA"][0], 0.0, g["A8"])
    alg = cls_of(name)(float(g["alpha"]), float(g["gamma"]), mouse, maze)
    rep = ShuffledReplays(n_times=int(g["n_times"]), seed=int(g["seed"]))
    exp = ConditioningExperiment(maze, alg, rep)
    poses, tabs = [], []

    def snap():
        poses.append([mouse.position[0], mouse.position[1], mouse.orientation])
        tabs.append(np.array(exp.v_table))

    exp.run(list(g["trajA"]), [0.0] * len(g["trajA"]), list(sign * g["rewA"]), do_replays=True)
    snap()
    exp.do_replays()
    snap()
    exp.run(list(g["trajB"]), [0.0] * len(g["trajB"]), list(sign * g["rewB"]), do_replays=False)
    snap()
    hist = np.array(mouse.history[0])
    firstB = [u.transition for u in rep._buffer][len(g["trajA"]) - 1:][:4]
    cand = np.full((4, 8, 2), -1, dtype=np.int64)
    for i, t in enumerate(firstB):
        for k, c in enumerate(t.extra["next_states"]):
            cand[i, k] = c
    return exp, np.array(poses), np.array(tabs), hist, cand


def _check(g, name, got):
    exp, poses, tabs, hist, cand = got
    assert np.array_equal(poses, g[f"{name}_pose"])
    assert np.array_equal(tabs, g[f"{name}_V"])
    assert len(hist) == int(g[f"{name}_hist_len"])
    assert np.array_equal(hist[-5:], g[f"{name}_hist_tail"])
    assert np.array_equal(cand, g[f"{name}_firstB_cand"])


@pytest.mark.parametrize("name", ["positive", "negative"])
def test_host_plugin_loop(g, name):
    """User-defined subclasses take the per-transition host loop (plugin interface): same tables and mouse."""
    from navigation_model_b200.simulation import rl

    def cls_of(n):
        base = rl.PositiveConditioning if n == "positive" else rl.NegativeConditioning
        return type("User" + base.__name__, (base,), {})

    got = _stages(g, name, cls_of)
    assert not got[0]._device_eligible()
    _check(g, name, got)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["positive", "negative"])
def test_device_launches(cuda, g, name):
    """The built-in learners run every stage as one launch; the learner's mouse is walked through the replayed arrival
    points natively (nm_mouse_move_through)."""
    from navigation_model_b200.simulation import rl
    got = _stages(g, name, lambda n: rl.PositiveConditioning if n == "positive" else rl.NegativeConditioning)
    assert got[0]._device_eligible()
    _check(g, name, got)
