"""Session ingestion on the device against the host mirror of the reference's `Session` /
`compute_orientations` / `adjust_positions_to_maze` (itself pinned by the reference's own test_session /
test_metrics suites, tests/test_reference_dropin.py) and against the host stream builder `nm_stream_build`.

Subsampling, adjusted positions and tile / candidate indices: bit-exact.  Headings: 4 ulp (device atan2 vs libm)."""
import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu


def _raw_recording(sid, n=4000, rate=30.0, holes=True):
    """A raw 30 Hz recording: jittered timestamps, a slow random walk over the maze (with stand-stills), sparse
    rewards held for a few frames; with `holes`, NaN timestamps, NaN / inf positions and NaN rewards."""
    rng = np.random.default_rng(1000 + sid)
    t = np.cumsum(rng.uniform(0.5, 1.5, n) / rate)
    omz = H.oracle_m9()
    traj, _ = O.synthetic_session(omz, sid, n // 10 + 1)
    pos = np.repeat(traj, 10, axis=0)[:n] + rng.normal(0, 1e-3, (n, 2))
    still = rng.random(n) < 0.2
    for i in range(1, n):
        if still[i]:
            pos[i] = pos[i - 1]
    rew = np.zeros(n)
    for s in rng.integers(0, n - 5, n // 100):
        rew[s:s + 4] = 1.0
    if holes:
        t[rng.random(n) < 0.02] = np.nan
        pos[rng.random(n) < 0.03] = np.nan
        pos[rng.random(n) < 0.002, 0] = np.inf
        rew[rng.random(n) < 0.02] = np.nan
        pos[200:260] = np.nan  # whole windows without a finite point are dropped
    return t, pos, rew


def _batch(sids, **kw):
    from navigation_model_b200.batched_sessions import SessionBatch
    raws = [_raw_recording(s, n=3000 + 137 * s, **kw) for s in sids]
    return raws, SessionBatch([r[0] for r in raws], [r[1] for r in raws], [r[2] for r in raws])


@pytest.mark.parametrize("dt", [0.1, 0.25, 1.0])
def test_subsample_matches_session(cuda, dt):
    from navigation_model_b200.analysis.data import Session
    raws, batch = _batch(range(7))
    sub = batch.subsample(dt)
    got = sub.to_host()
    for (t, pos, rew), (gt, gr) in zip(raws, got):
        want_t, want_r = Session._subsample(t, pos, rew, dt)
        assert len(gt) == len(want_t) > 10
        assert np.array_equal(gt, np.array(want_t))
        assert np.array_equal(gr, np.array(want_r))


def test_subsample_without_reward_and_degenerate(cuda):
    from navigation_model_b200.analysis.data import Session
    from navigation_model_b200.batched_sessions import SessionBatch
    t0, p0, _ = _raw_recording(3, n=500)
    cases = [(t0, p0), (np.array([0.0]), np.zeros((1, 2))), (np.full(5, np.nan), np.ones((5, 2))),
             (np.arange(4.0), np.full((4, 2), np.nan))]
    batch = SessionBatch([c[0] for c in cases], [c[1] for c in cases])
    sub = batch.subsample(0.2)
    for (t, p), (gt, gr) in zip(cases, sub.to_host()):
        want, _ = Session._subsample(t, p, None, 0.2)
        assert gr is None
        assert len(gt) == len(want)
        if len(want):
            assert np.array_equal(gt, np.array(want))
    with pytest.raises(RuntimeError):
        SessionBatch([np.arange(3.0)], [np.zeros((4, 2))])
    with pytest.raises(RuntimeError):
        SessionBatch([np.arange(3.0)], [np.zeros((3, 2))], [np.zeros(2)])


def test_adjust_positions(cuda):
    from navigation_model_b200.analysis.data import adjust_positions_to_maze
    from navigation_model_b200.batched_sessions import SessionBatch
    maze = H.product_m9()
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.2, 1.2, (5000, 2))  # inside, in the void block and outside of the 9 x 9 maze
    pts[rng.random(5000) < 0.02] = np.nan
    batch = SessionBatch(None, [pts[:2000], pts[2000:]])
    got = np.concatenate([p for p, _ in batch.adjust_to_maze(maze).to_host()])
    want = np.array(adjust_positions_to_maze(list(pts), maze))
    assert np.array_equal(got, want, equal_nan=True)
    assert (got != pts)[~np.isnan(pts)].any()


@pytest.mark.parametrize("with_maze", [False, True])
def test_orientations(cuda, with_maze):
    from navigation_model_b200.analysis.metrics import compute_orientations
    raws, batch = _batch(range(5), holes=False)
    sub = batch.subsample(0.2)
    maze = H.product_m9() if with_maze else None
    acts = H.a8() if with_maze else None
    got = sub.orientations(maze, acts)
    for (traj, _), g in zip(sub.to_host(), got):
        want = np.array(compute_orientations(list(traj), maze=maze, possible_actions=None if acts is None else
                                             [tuple(a) for a in acts]))
        assert g.shape == want.shape
        # the device heading is the C library's atan2 (csrc/nm_libm.h); the host mirror, like the reference
        # (metrics.py:679), calls np.arctan2, which on AVX-512 hosts is numpy's SVML kernel and differs from the C
        # library in the last place for ~7 % of the arguments -- hence a few ulp, not equality, on THIS comparison
        np.testing.assert_allclose(g, want, rtol=0, atol=4 * np.finfo(float).eps * np.pi)
        assert len(np.unique(want)) > 20


def test_orientations_never_moving(cuda):
    from navigation_model_b200.batched_sessions import SessionBatch
    batch = SessionBatch(None, [np.ones((6, 2)), np.cumsum(np.ones((6, 2)), axis=0)])
    with pytest.raises(IndexError):
        batch.orientations()


@pytest.mark.parametrize("with_actions", [False, True])
def test_streams_match_host_builder(cuda, with_actions):
    from navigation_model_b200.simulation.mouse import Mouse
    from navigation_model_b200.sweep import build_session_records
    maze = H.product_m9()
    raws, batch = _batch(range(6), holes=True)
    sub = batch.subsample(0.3).adjust_to_maze(maze)
    acts = H.a8() if with_actions else None
    ds = sub.streams(maze, acts)
    got = ds.download()
    for (traj, rew), g in zip(sub.to_host(), got):
        mouse = Mouse(traj[0], 0.0, acts, save_history=False) if with_actions else None
        want = build_session_records(maze, traj, rew, mouse=mouse)
        assert len(g) == len(want) == len(traj) - 1
        for f in ("r", "s", "s1", "ncand", "obs", "cand"):
            assert np.array_equal(g[f], want[f]), f


def test_streams_with_mouse_init(cuda):
    from navigation_model_b200.simulation.mouse import Mouse
    from navigation_model_b200.sweep import build_session_records
    maze = H.product_m9()
    omz = H.oracle_m9()
    from navigation_model_b200.batched_sessions import SessionBatch
    sessions = [O.synthetic_session(omz, sid, 400 + 31 * sid) for sid in range(4)]
    batch = SessionBatch(None, [s[0] for s in sessions], [s[1] for s in sessions])
    init = np.array([[t[0, 0] + 0.01, t[0, 1] - 0.01, 0.3 * j] for j, (t, _) in enumerate(sessions)])
    got = batch.streams(maze, H.a8(), mouse_init=init).download()
    for (traj, rew), g, ini in zip(sessions, got, init):
        want = build_session_records(maze, traj, rew, mouse=Mouse(ini[:2], ini[2], H.a8(), save_history=False))
        assert g.tobytes() == want.tobytes()


@pytest.mark.parametrize("alg", ["td0", "positive"])
def test_sweep_on_device_built_streams(cuda, alg):
    """raw recordings -> subsample -> streams on the device -> fitness sweep == the same through host sessions."""
    from navigation_model_b200.batched_sessions import SessionBatch
    from navigation_model_b200.sweep import ConditioningSweep
    maze = H.product_m9()
    omz = H.oracle_m9()
    sessions = [O.synthetic_session(omz, sid, 700 + 50 * sid) for sid in range(5)]
    rng = np.random.default_rng(3)
    a, b, g = H.grid_params(rng, 200)
    host = ConditioningSweep(maze, alg, possible_actions=H.a8())
    for traj, rew in sessions:
        host.add_session(traj, rew)
    dev = ConditioningSweep(maze, alg, possible_actions=H.a8())
    dev.use_streams(SessionBatch(None, [s[0] for s in sessions], [s[1] for s in sessions]).streams(maze, H.a8()))
    assert dev.n_sessions == 5 and dev.updates_per_parameter_set == host.updates_per_parameter_set
    ll_h, v_h = host.log_likelihood(a, b, g, return_v=True)
    ll_d, v_d = dev.log_likelihood(a, b, g, return_v=True)
    assert np.array_equal(ll_h, ll_d) and np.array_equal(v_h, v_d)
    assert dev.best_fit(a, b, g) == host.best_fit(a, b, g)
    with pytest.raises(RuntimeError):
        dev.add_session(*sessions[0])


@pytest.mark.parametrize("agent", ["qtable", "modelbased"])
def test_tile_action_agents_on_device_built_streams(cuda, agent):
    """The two tile-action agents only need (s, s1, r): streams built on the device without actions serve them."""
    from navigation_model_b200.batched_sessions import SessionBatch
    from navigation_model_b200.sweep import ModelBasedSweep, QLearningSweep
    maze = H.product_m9()
    omz = H.oracle_m9()
    moves = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    sessions = [O.synthetic_session(omz, sid, 300 + 40 * sid) for sid in range(3)]
    rng = np.random.default_rng(5)
    a, b, g = H.grid_params(rng, 50)
    make = (lambda: QLearningSweep(maze, moves)) if agent == "qtable" else (lambda: ModelBasedSweep(maze, moves, n_sweeps=2))
    host, dev = make(), make()
    for traj, rew in sessions:
        host.add_session(traj, rew)
    dev.use_streams(SessionBatch(None, [s[0] for s in sessions], [s[1] for s in sessions]).streams(maze, None))
    assert dev.n_sessions == 3 and dev.updates_per_parameter_set == host.updates_per_parameter_set
    assert np.array_equal(host.log_likelihood(a, b, g), dev.log_likelihood(a, b, g))
    assert dev.best_fit(a, b, g) == host.best_fit(a, b, g)


@pytest.mark.parametrize("alg", ["td0", "positive"])
def test_replays_on_device_built_streams(cuda, alg):
    """Shuffled replay passes (rl.py:419-424) appended to device-built streams by a device gather: same records, same
    tables and likelihoods as host sessions with the same replay strategy."""
    from navigation_model_b200.batched_sessions import SessionBatch
    from navigation_model_b200.sweep import ConditioningSweep
    maze = H.product_m9()
    omz = H.oracle_m9()
    sessions = [O.synthetic_session(omz, sid, 200 + 30 * sid) for sid in range(4)]
    rng = np.random.default_rng(7)
    a, b, g = H.grid_params(rng, 70)
    host = ConditioningSweep(maze, alg, possible_actions=H.a8(), replay_n_times=2, replay_seed=11)
    for traj, rew in sessions:
        host.add_session(traj, rew)
    plain = SessionBatch(None, [s[0] for s in sessions], [s[1] for s in sessions]).streams(maze, H.a8())
    with_rep = plain.with_replays(2, seed=11)
    for got, want, n_on in zip(with_rep.download(), host._records, host._online):
        assert got.tobytes() == want.tobytes() and len(got) == 3 * n_on
    assert plain.total * 3 == with_rep.total  # the source streams are untouched
    dev = ConditioningSweep(maze, alg, possible_actions=H.a8(), replay_n_times=2, replay_seed=11)
    dev.use_streams(plain)  # applies the sweep's replay strategy itself
    assert dev.updates_per_parameter_set == host.updates_per_parameter_set
    ll_h, v_h = host.log_likelihood(a, b, g, return_v=True)
    ll_d, v_d = dev.log_likelihood(a, b, g, return_v=True)
    assert np.array_equal(ll_h, ll_d) and np.array_equal(v_h, v_d)
    with pytest.raises(RuntimeError):
        with_rep.with_replays(1)


def test_no_candidate_is_an_error(cuda):
    """A step whose rotated endpoints all fall outside the maze: np.max([]) raises in the reference (rl.py:184)."""
    from navigation_model_b200.batched_sessions import SessionBatch
    maze = H.product_m9()
    traj = np.array([[0.05, 0.05], [0.06, 0.05], [0.07, 0.05]])
    batch = SessionBatch(None, [traj], [np.zeros(3)])
    with pytest.raises((ValueError, RuntimeError)):
        batch.streams(maze, np.array([[-5.0, -5.0]]))
