"""Host-side logic that needs no GPU: native stream builder vs the oracle's extraction, replay orders,
grid sharding + the world_size-2 gather of per-shard best fits over gloo, loud failure without CUDA."""
import os
import sys

import numpy as np
import pytest

from oracle import np_oracle as O

import helpers as H
from conftest import has_cuda


def _records(maze, traj, rew, actions):
    from navigation_model_b200.simulation.mouse import Mouse
    from navigation_model_b200.sweep import build_session_records
    mouse = None if actions is None else Mouse(traj[0], 0.0, actions, save_history=False)
    return build_session_records(maze, traj, rew, mouse=mouse)


@pytest.mark.parametrize("sid,T", [(0, 5000), (7, 300), (9, 2)])
def test_native_stream_builder_matches_oracle(sid, T):
    omz, maze = H.oracle_m9(), H.product_m9()
    traj, rew = O.synthetic_session(omz, sid, T)
    tr = H.oracle_transitions(omz, traj, rew, H.a8())
    recs = _records(maze, traj, rew, H.a8())
    assert len(recs) == T - 1
    assert np.array_equal(recs["s"], tr["s"]) and np.array_equal(recs["s1"], tr["s1"])
    assert np.array_equal(recs["r"], tr["r"]) and np.array_equal(recs["ncand"], tr["ncand"])
    assert np.array_equal(recs["obs"], tr["obs"])
    for i in range(len(recs)):
        k = recs["ncand"][i]
        assert np.array_equal(recs["cand"][i, :k], tr["cand"][i, :k])
    # without a mouse: no candidates
    plain = _records(maze, traj, rew, None)
    assert not plain["ncand"].any() and (plain["obs"] == 0xFF).all()
    assert np.array_equal(plain["s"], tr["s"])


def test_stream_builder_golden_transitions(golden_fit):
    """...and against the transitions the reference itself extracted."""
    g = golden_fit
    maze = H.product_m9()
    recs = _records(maze, g["m9_traj"], g["m9_rew"], g["A8"])
    tos = maze.flat_tables()["tile_of_state"]
    cells = np.stack([tos // 9, tos % 9], axis=1)
    assert np.array_equal(cells[recs["s"]], g["m9_tr_s"]) and np.array_equal(cells[recs["s1"]], g["m9_tr_s1"])
    assert np.array_equal(recs["r"], g["m9_tr_r"]) and np.array_equal(recs["ncand"], g["m9_tr_ncand"])
    for i in range(len(recs)):
        k = recs["ncand"][i]
        assert np.array_equal(cells[recs["cand"][i, :k]], g["m9_tr_cand"][i, :k])


def test_replay_records_follow_generator_shuffle():
    from navigation_model_b200.sweep import replay_records
    omz, maze = H.oracle_m9(), H.product_m9()
    traj, rew = O.synthetic_session(omz, 1, 200)
    recs = _records(maze, traj, rew, H.a8())
    rep = replay_records(recs, 3, seed=5)
    orders = O.replay_orders(len(recs), 3, 5)
    assert np.array_equal(rep, recs[np.concatenate(orders)])
    # same permutation as shuffling a list of objects in place (what the reference does, rl.py:420)
    objs = list(range(len(recs)))
    rng = np.random.default_rng(5)
    rng.shuffle(objs)
    assert objs == orders[0].tolist()


def test_shard_bounds_cover_grid():
    from navigation_model_b200.sweep import shard_bounds
    for P in (1, 7, 8, 100, 255, 256, 131072, 100352, 2097152 + 3):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(P, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            if -(-P // 32) >= world:  # shards start on multiples of 32 parameter sets, sizes within two blocks
                assert all(lo % 32 == 0 for lo, _ in spans) and max(sizes) - min(sizes) < 64
            else:
                assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    from navigation_model_b200.sweep import reduce_best, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nll = np.random.default_rng(123).random(1001)
    nll[[17, 600]] = nll.min() - 1.0  # a tie across the two shards: the lowest index must win
    lo, hi = shard_bounds(len(nll), rank, world)
    local = nll[lo:hi]
    best = torch.empty(2, dtype=torch.float64)
    best[0] = float(local.min())
    best[1:].view(torch.int64)[0] = int(np.argmin(local))
    v, i = reduce_best(best, lo)
    q.put((rank, v, i))
    dist.destroy_process_group()


def test_best_fit_gather_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nll = np.random.default_rng(123).random(1001)
    nll[[17, 600]] = nll.min() - 1.0
    for _, v, i in res:
        assert i == 17 and v == nll[17]


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    if has_cuda():
        pytest.skip("CUDA device present")
    from navigation_model_b200.sweep import ConditioningSweep
    omz, maze = H.oracle_m9(), H.product_m9()
    traj, rew = O.synthetic_session(omz, 0, 50)
    sw = ConditioningSweep(maze, "td0")
    sw.add_session(traj, rew)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        sw.v_tables([0.1], [0.9])
    from navigation_model_b200.simulation.rl import ConditioningExperiment, ShuffledReplays, TD0
    exp = ConditioningExperiment(maze, TD0(0.1, 0.9), ShuffledReplays(0))
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        exp.run(list(traj), [0.0] * len(traj), list(rew))


def test_product_never_imports_oracle():
    """No product file imports, loads, links or executes anything under oracle/ (comments may name it)."""
    import re
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "navigation_model_b200")
    bad = re.compile(r"(^|\s)(from|import)\s+oracle\b|liboracle|oracle/|np_oracle|c_oracle|nm_oracle|refharness")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f"{f} depends on the oracle"


def test_workload_recipes_match_the_oracle_restatement():
    """bench.py and tools/ build their inputs from navigation_model_b200.workloads (product side); the oracle keeps its
    own restatement of the same seeded recipe for the parity tests.  The two must give identical bits."""
    from navigation_model_b200 import workloads as W
    from oracle import np_oracle as O
    assert W.M9_LAYOUT == O.M9_LAYOUT and W.M9_SUBAREAS == O.M9_SUBAREAS and W.M9_SIZE_TILE == O.M9_SIZE_TILE
    assert np.array_equal(W.A8_UNITS, O.A8_UNITS)
    m9, om9 = W.maze_m9(), O.OracleMaze(O.M9_LAYOUT, O.M9_SIZE_TILE)
    for sid, T in ((0, 5000), (3, 257), (9, 1)):
        got, want = W.synthetic_session(m9, sid, T), O.synthetic_session(om9, sid, T)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    m20, om20 = W.maze_m20(), O.OracleMaze("\n".join(["M" * 20] * 20), W.M20_SIZE_TILE)
    got = W.synthetic_session(m20, 0, 1000, start=(3, 3), goal=(15, 12), jitter=0.02)
    want = O.synthetic_session(om20, 0, 1000, start=(3, 3), goal=(15, 12), jitter=0.02)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.array_equal(m20.transition_table(W.MOVES8), O.transition_table(om20, W.MOVES8))


def test_bench_product_arm_touches_the_oracle_only_in_its_cpu_legs():
    """bench.py may execute oracle/ only in its CPU legs: the cpu_baseline measurements (headline, configs[0],
    simulation) and --impl reference."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    allowed = {"oracle_transitions", "cpu_port_rate", "run_reference", "cpu_baseline_config0", "cpu_port_sim_rate"}
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        uses = any(isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle" for n in ast.walk(fn)) or \
            any(isinstance(n, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in n.names) for n in ast.walk(fn))
        assert not uses or fn.name in allowed, f"bench.py:{fn.name} imports the oracle"
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any("oracle" in ast.dump(n) for n in top)


def test_regroup_order_of_tile_action_sweeps():
    """The host-side re-ordering that lets grids given in any axis order share tables per run of 32 sets."""
    from navigation_model_b200.sweep import _TileActionSweep, parameter_grid
    a, b, g = parameter_grid(np.linspace(0.1, 1, 7), np.logspace(-1, 1, 64), np.linspace(0.5, 0.9, 3))  # gamma fastest
    order = _TileActionSweep.regroup_order(a, g)
    assert order is not None and sorted(order.tolist()) == list(range(a.size))
    A, G, B = a[order].reshape(-1, 32), g[order].reshape(-1, 32), b[order].reshape(-1, 64)
    assert (A == A[:, :1]).all() and (G == G[:, :1]).all()
    assert (np.diff(B, axis=1) > 0).all()  # stable: the beta values of a pair keep their order
    g2, a2, _ = parameter_grid(np.linspace(0.5, 0.9, 3), np.linspace(0.1, 1, 7), np.logspace(-1, 1, 64))
    assert _TileActionSweep.regroup_order(a2, g2) is None  # already laid out in runs
    a3, _, g3 = parameter_grid(np.linspace(0.1, 1, 7), np.logspace(-1, 1, 20), np.linspace(0.5, 0.9, 3))
    assert _TileActionSweep.regroup_order(a3, g3) is None  # 20 beta points: no order gives uniform runs of 32
    rng = np.random.default_rng(0)
    assert _TileActionSweep.regroup_order(rng.random(500), rng.random(500)) is None
    # a last, incomplete run is allowed when its pair sorts last
    a4 = np.concatenate([np.full(64, 0.2), np.full(40, 0.9)])[::-1].copy()
    g4 = np.full(104, 0.7)
    order = _TileActionSweep.regroup_order(a4, g4)
    assert order is not None and (a4[order][:64] == 0.2).all() and (a4[order][64:] == 0.9).all()



def test_upload_staging_converts_and_transposes():
    """The pinned-arena upload path of the batched simulation (host logic only: pinned allocation and the device copy
    are stubbed): any dtype / strides in, float64 in the caller's shape out; arenas are reused by the next run."""
    import torch
    from navigation_model_b200.batched_sim import BatchedExperiment
    from navigation_model_b200.simulation import behavioural_components as bc

    class Shim:  # torch without pinned memory
        float64, uint8 = torch.float64, torch.uint8
        from_numpy = staticmethod(torch.from_numpy)

        @staticmethod
        def empty(n, dtype=None, pin_memory=False):
            return torch.empty(n, dtype=dtype)

    exp = BatchedExperiment(H.product_m9(), H.a8(), O.A8_UNITS, [bc.AnxietyComponent])
    rng = np.random.default_rng(0)
    arrays = [rng.random((300000, 8)), rng.integers(0, 9, 2_000_000), rng.random((50, 400000)).T, rng.random(7),
              np.float32(rng.random((200000, 3)))]
    exp._stage_reset()
    for x in arrays:
        t = exp._upload_f64(Shim, x, "cpu")
        assert t.dtype == torch.float64 and tuple(t.shape) == x.shape
        assert np.array_equal(t.numpy(), x.astype(np.float64))
    used = [ar[1] for ar in exp._arenas]
    assert sum(used) >= sum(a.size * 8 for a in arrays if a.nbytes >= (1 << 20))
    exp._stage_reset()
    assert all(ar[1] == 0 for ar in exp._arenas)
    n_arenas = len(exp._arenas)
    exp._upload_f64(Shim, arrays[0], "cpu")
    assert len(exp._arenas) == n_arenas  # no new pinned allocation for the next run


class _FakeBase:
    """CPU stand-in for the device side of a sweep: the likelihood is a fixed function of the parameters, the arg-min is
    torch on the CPU.  Lets sharded_best_fit's host logic (broadcast, re-ordering, sharding, tie-break) run anywhere."""

    def __init__(self):
        import torch
        self.torch_device = torch.device("cpu")

    @staticmethod
    def _to_dev(torch, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    @staticmethod
    def ll(at, bt, gt, _vi=None):
        import torch
        nll = (at - 0.3) ** 2 + (torch.log(bt) - 1.0) ** 2 + (gt - 0.8) ** 2
        return torch.stack([-0.25 * nll, -0.75 * nll])  # two "sessions"

    @staticmethod
    def argmin_device(ll, nll_out=None):
        import torch
        nll = -(ll[0] + ll[1])
        if nll_out is not None:
            nll_out.copy_(nll)
        best = torch.empty(2, dtype=torch.float64)
        i = int(torch.argmin(nll))
        best[0] = nll[i]
        best[1:].view(torch.int64)[0] = i
        return best


def test_best_fit_broadcasts_scalars_and_validates_lengths():
    """best_fit with a scalar alpha evaluated one parameter set only.  Scalars broadcast now, like
    log_likelihood; other length mismatches raise."""
    from navigation_model_b200.sweep import sharded_best_fit
    base = _FakeBase()
    betas = np.logspace(-1, 1.5, 50)
    v, i = sharded_best_fit(base, base.ll, 0.3, betas, 0.8)
    assert i == int(np.argmin((np.log(betas) - 1.0) ** 2)) and i != 0
    with pytest.raises(ValueError):
        sharded_best_fit(base, base.ll, [0.1, 0.2], betas, 0.8)


def test_best_fit_regroups_and_answers_in_the_callers_order():
    from navigation_model_b200.sweep import parameter_grid, regroup_order, sharded_best_fit
    base = _FakeBase()
    a, b, g = parameter_grid(np.linspace(0.1, 1, 7), np.logspace(-1, 1, 64), np.linspace(0.5, 0.9, 3))  # gamma fastest
    assert regroup_order(a, g) is not None
    seen = []

    def ll(at, bt, gt, vi=None):
        seen.append((at.numpy().copy(), gt.numpy().copy()))
        return base.ll(at, bt, gt)

    v, i = sharded_best_fit(base, ll, a, b, g)
    A, G = seen[0][0].reshape(-1, 32), seen[0][1].reshape(-1, 32)
    assert (A == A[:, :1]).all() and (G == G[:, :1]).all()  # the device saw runs of 32 sharing (alpha, gamma)
    nll = (a - 0.3) ** 2 + (np.log(b) - 1.0) ** 2 + (g - 0.8) ** 2
    assert i == int(np.argmin(nll))
    assert sharded_best_fit(base, base.ll, a, b, g, regroup=False)[1] == i
    # ties go to the lowest index of the CALLER's order, whatever order the device saw
    a2, b2, g2 = np.tile(a, 2), np.tile(b, 2), np.tile(g, 2)
    assert sharded_best_fit(base, base.ll, a2, b2, g2)[1] == i


def _gloo_best_fit_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from navigation_model_b200.sweep import parameter_grid, sharded_best_fit
    base = _FakeBase()
    a, b, g = parameter_grid(np.linspace(0.1, 1, 9), np.logspace(-1, 1, 32), np.linspace(0.5, 0.9, 5))
    q.put((rank,) + tuple(sharded_best_fit(base, base.ll, a, b, g)))
    dist.destroy_process_group()


def test_sharded_best_fit_world2_gloo():
    import torch.multiprocessing as mp
    from navigation_model_b200.sweep import parameter_grid
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_best_fit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b, g = parameter_grid(np.linspace(0.1, 1, 9), np.logspace(-1, 1, 32), np.linspace(0.5, 0.9, 5))
    want = int(np.argmin((a - 0.3) ** 2 + (np.log(b) - 1.0) ** 2 + (g - 0.8) ** 2))
    assert [r[2] for r in res] == [want, want]
