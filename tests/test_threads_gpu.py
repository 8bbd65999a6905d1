"""Re-entrancy of the C-ABI (include/navmodel_b200.h: "re-entrant per (device, stream)"): three host threads sweep
concurrently on their own CUDA streams -- one with a grid that shares tables per run, two with arbitrary parameter
sets (tables in device memory) -- and must get exactly what the same calls return one after the other (the per-stream decision word, the scratch
blocks and the launch sequences must not leak between them)."""
import threading

import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H

pytestmark = pytest.mark.gpu


def test_threads_on_their_own_streams(cuda):
    import torch
    from navigation_model_b200.sweep import ConditioningSweep, QLearningSweep
    omz, maze = H.oracle_m9(), H.product_m9()
    moves = [(dx, dy) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (dx, dy) != (0, 0)]
    traj, rew = O.synthetic_session(omz, 1, 400)
    rng = np.random.default_rng(0)
    grid = (np.repeat(rng.uniform(0.01, 1, 8), 32), np.tile(np.logspace(-1, 1.5, 32), 8),
            np.repeat(rng.uniform(0.5, 0.99, 8), 32))
    arbitrary = H.grid_params(rng, 300)

    def make():
        sw = ConditioningSweep(maze, "positive", possible_actions=H.a8())
        sw.add_session(traj, rew)
        ql = QLearningSweep(maze, moves)
        ql.add_session(traj, rew)
        return sw, ql

    def work(params, n_rep, out, want_modes):
        sw, ql = make()
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for _ in range(n_rep):
                ll = sw.log_likelihood(*params)
                m1 = sw.last_sweep_mode()
                lq = ql.log_likelihood(*params, regroup=False)
                m2 = ql.last_sweep_mode()
                out.append((ll, lq, m1, m2))
        assert all(o[2] == want_modes[0] and o[3] == want_modes[1] for o in out)

    serial_grid, serial_arb = [], []
    work(grid, 1, serial_grid, (1, 1))
    work(arbitrary, 1, serial_arb, (2, 2))
    arbitrary2 = H.grid_params(rng, 517)
    serial_arb2 = []
    work(arbitrary2, 1, serial_arb2, (2, 2))
    got_grid, got_arb, got_arb2, errors = [], [], [], []

    def guarded(*args):
        try:
            work(*args)
        except Exception as exc:  # surfaced below: an assertion in a thread would otherwise be lost
            errors.append(exc)

    threads = [threading.Thread(target=guarded, args=(grid, 6, got_grid, (1, 1))),
               threading.Thread(target=guarded, args=(arbitrary, 6, got_arb, (2, 2))),
               threading.Thread(target=guarded, args=(arbitrary2, 6, got_arb2, (2, 2)))]  # device-memory tables twice
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for ll, lq, _, _ in got_grid:
        assert np.array_equal(ll, serial_grid[0][0]) and np.array_equal(lq, serial_grid[0][1])
    for ll, lq, _, _ in got_arb:
        assert np.array_equal(ll, serial_arb[0][0]) and np.array_equal(lq, serial_arb[0][1])
    for ll, lq, _, _ in got_arb2:
        assert np.array_equal(ll, serial_arb2[0][0]) and np.array_equal(lq, serial_arb2[0][1])
