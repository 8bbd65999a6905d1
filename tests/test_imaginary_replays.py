"""ShuffledReplays(imaginary=True) (reference simulation/rl.py:388-418): transitions sampled
from the maze with the stdlib ``random`` module.  Seeding ``random`` makes the reference replayable; the golden file
(tests/golden/imaginary_golden.npz, written by ``make_golden.py imaginary``) holds what the unmodified reference
produced under ``random.seed(4242)``.  Host plugin loop: CPU test.  Device path (the sampled transitions are flattened
to step records and the updates run in the sweep kernel): GPU test.  V-tables bit-exact."""
import os
import random

import numpy as np
import pytest

import helpers as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gi():
    return dict(np.load(os.path.join(GOLDEN, "imaginary_golden.npz"), allow_pickle=False))


def _experiment(gi, alg_name, a, g):
    from navigation_model_b200.simulation.mouse import Mouse
    from navigation_model_b200.simulation.rl import (ConditioningExperiment, NegativeConditioning,
                                                      PositiveConditioning, ShuffledReplays, TD0)
    maze = H.product_m9()
    acts = H.a8()
    mouse = Mouse(gi["traj"][0], 0.0, acts)
    alg = {"td0": lambda: TD0(a, g), "positive": lambda: PositiveConditioning(a, g, mouse, maze),
           "negative": lambda: NegativeConditioning(a, g, mouse, maze)}[alg_name]()
    rep = ShuffledReplays(n_times=2, seed=1, imaginary=True, maze=maze, possible_actions=acts)
    return ConditioningExperiment(maze, alg, rep), alg, rep


def test_sampled_transitions_match_reference(gi):
    """Same stdlib draws in the same order: start tile, heading, arrival tile, reward of every sampled transition."""
    exp, alg, rep = _experiment(gi, "positive", 0.1, 0.9)
    random.seed(int(gi["random_seed"]))
    exp._run_host(list(gi["traj"]), [0.0] * len(gi["traj"]), list(gi["rew"]), False)
    trans = rep.imaginary_transitions(alg)
    assert len(trans) == len(gi["im_s"]) == 2 * (len(gi["traj"]) - 1)
    assert np.array_equal(np.array([t.s for t in trans]), gi["im_s"])
    assert np.array_equal(np.array([t.s1 for t in trans]), gi["im_s1"])
    assert np.array_equal(np.array([t.r for t in trans], dtype=np.float64), gi["im_r"])
    assert np.array_equal(np.array([t.ori for t in trans]), gi["im_ori"])
    # flattened to records by the native builder: the algorithm's mouse finds the reference's candidate counts
    from navigation_model_b200 import sweep
    recs = sweep.records_from_transitions(H.product_m9(), trans, mouse=alg._mouse)
    assert np.array_equal(recs["ncand"], gi["im_ncand"])
    assert np.array_equal(recs["r"], gi["im_r"])


@pytest.mark.parametrize("alg_name", ["td0", "positive", "negative"])
def test_host_loop_matches_reference(gi, alg_name):
    for k, (a, g) in enumerate(gi["params"]):
        exp, alg, rep = _experiment(gi, alg_name, a, g)
        rew = -gi["rew"] if alg_name == "negative" else gi["rew"]
        random.seed(int(gi["random_seed"]))
        exp._run_host(list(gi["traj"]), [0.0] * len(gi["traj"]), list(rew), False)
        rep.offline_replays(alg)
        rep.offline_replays(alg)  # second round, the stdlib stream continues
        assert np.array_equal(exp.v_table, gi[f"{alg_name}_V"][k]), (alg_name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("alg_name", ["td0", "positive", "negative"])
def test_device_path_matches_reference(cuda, gi, alg_name):
    for k, (a, g) in enumerate(gi["params"]):
        exp, alg, rep = _experiment(gi, alg_name, a, g)
        assert exp._device_eligible()
        rew = -gi["rew"] if alg_name == "negative" else gi["rew"]
        random.seed(int(gi["random_seed"]))
        exp.run(list(gi["traj"]), [0.0] * len(gi["traj"]), list(rew), do_replays=True)
        exp.do_replays()
        assert np.array_equal(exp.v_table, gi[f"{alg_name}_V"][k]), (alg_name, k)
