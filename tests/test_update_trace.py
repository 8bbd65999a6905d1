"""The replay buffer's Update.dv / Update.rpe: ConditioningExperiment.run on the device fills them with what
the reference's update returns per online step (tests/golden/updates_golden.npz, from the unmodified reference:
rl.py:190-194, 233-237, 260-262) -- bit for bit, through nm_vtable_trace; the host plugin loop does too."""
import os

import numpy as np
import pytest

import helpers as H
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "updates_golden.npz"), allow_pickle=False))


def _run(g, name, subclass):
    from navigation_model_b200.simulation import rl
    from navigation_model_b200.simulation.mouse import Mouse
    maze = H.product_m9()
    sign = -1.0 if name == "negative" else 1.0
    mouse = Mouse(g["traj"][0], 0.0, g["A8"])
    base = {"td0": rl.TD0, "positive": rl.PositiveConditioning, "negative": rl.NegativeConditioning}[name]
    cls = type("User" + base.__name__, (base,), {}) if subclass else base
    alg = cls(float(g["alpha"]), float(g["gamma"])) if name == "td0" else cls(float(g["alpha"]), float(g["gamma"]), mouse, maze)
    rep = rl.ShuffledReplays(n_times=2, seed=9)
    exp = rl.ConditioningExperiment(maze, alg, rep, v_table=g["v0"].copy())
    exp.run(list(g["traj"]), [0.0] * len(g["traj"]), list(sign * g["rew"]), do_replays=False)
    dv = np.array([float(u.dv) for u in rep._buffer])
    rpe = np.array([float(u.rpe) for u in rep._buffer])
    v_online = np.array(exp.v_table)
    exp.do_replays()
    return exp, dv, rpe, v_online, np.array(exp.v_table)


def _check(g, name, got):
    _, dv, rpe, v_online, v_replayed = got
    assert np.array_equal(dv, g[f"{name}_dv"]) and np.array_equal(rpe, g[f"{name}_rpe"])
    assert np.array_equal(v_online, g[f"{name}_V_online"]) and np.array_equal(v_replayed, g[f"{name}_V_replayed"])


@pytest.mark.parametrize("name", ["td0", "positive", "negative"])
def test_host_plugin_loop(g, name):
    got = _run(g, name, subclass=True)
    assert not got[0]._device_eligible()
    _check(g, name, got)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["td0", "positive", "negative"])
def test_device_launch_fills_the_buffer(cuda, g, name):
    got = _run(g, name, subclass=False)
    assert got[0]._device_eligible()
    _check(g, name, got)
