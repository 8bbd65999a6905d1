"""ctypes loader of the C oracle (``oracle/liboracle.so``) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

See ``oracle/nm_oracle.c`` for the restated reference lines.  Built by ``make -C oracle`` (also run by
``__graft_entry__.build()``); the built library travels to the GPU box with the repo snapshot.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "nm_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        h = C.CDLL(LIB_PATH)
        h.orc_fit.restype = C.c_int
        h.orc_fit.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int64, C.c_void_p, C.c_int64] + \
            [C.c_void_p] * 3 + [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        h.orc_max_threads.restype = C.c_int
        _lib = h
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fit(alg, tr, alpha, gamma, n_states, beta=None, v0=None, orders=(), want_v=True, threads=0):
    """P parameter sets over one session (transitions ``tr`` from ``np_oracle.extract_transitions``).

    :return: ``(V[P, n_states] or None, ll[P] or None)``
    """
    alpha = np.ascontiguousarray(alpha, dtype=np.float64).reshape(-1)
    gamma = np.ascontiguousarray(gamma, dtype=np.float64).reshape(-1)
    P = alpha.size
    with_ll = beta is not None
    beta_a = np.ascontiguousarray(beta, dtype=np.float64).reshape(-1) if with_ll else None
    s = np.ascontiguousarray(tr["s"], dtype=np.int32)
    s1 = np.ascontiguousarray(tr["s1"], dtype=np.int32)
    r = np.ascontiguousarray(tr["r"], dtype=np.float64)
    ncand = np.ascontiguousarray(tr["ncand"], dtype=np.int32)
    cand = np.ascontiguousarray(tr["cand"], dtype=np.int32)
    obs = np.ascontiguousarray(tr["obs"], dtype=np.int32)
    replay = np.ascontiguousarray(np.concatenate(list(orders)) if len(orders) else np.zeros(0), dtype=np.int64)
    v0a = None if v0 is None else np.ascontiguousarray(v0, dtype=np.float64).reshape(-1)
    V = np.empty((P, n_states)) if want_v else None
    ll = np.empty(P) if with_ll else None
    rc = lib().orc_fit(int(alg), int(with_ll), _ptr(s), _ptr(s1), _ptr(r), _ptr(ncand), _ptr(cand), _ptr(obs),
                       len(s), _ptr(replay), len(replay), _ptr(alpha), _ptr(beta_a), _ptr(gamma), P, int(n_states),
                       _ptr(v0a), _ptr(V), _ptr(ll), int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_fit failed ({rc})")
    return V, ll


def simulate(maze, actions, discrete_actions, kinds, P, init_pose, init_disc, draws, vtable_norm=None, bcost_tables=None,
             threads=0, history=True):
    """N mice x T steps with replayed draws through ``orc_simulate`` (the compiled twin of ``np_oracle.simulate``).

    :param maze: ``np_oracle.OracleMaze``
    :param kinds: component kinds in model order (``np_oracle.COMP_*``)
    :param P: dict of per-mouse parameter arrays (``beta`` + the reference's parameter names; ``p4`` optional)
    :param bcost_tables: ``[N, A]`` cached von-Mises values (``np_oracle.bcost_table`` per mouse); computed here if None
    :return: dict(hist_pos [N,T+1,2], hist_ori [N,T+1], choice [N,T], occupancy [N,X,Y], status [N], final_pose [N,3])
    """
    from . import np_oracle as O
    h = lib()
    if not hasattr(h, "_sim_ready"):
        h.orc_simulate.restype = C.c_int
        h.orc_simulate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                   C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 17 + [C.c_int]
        h._sim_ready = True
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    actions = f64(actions).reshape(-1, 2)
    units = np.asarray(discrete_actions)
    A = len(actions)
    zero_act = np.ascontiguousarray(np.all(units == 0, axis=1), dtype=np.uint8)
    draws = f64(draws)
    N, T = draws.shape
    col = lambda *names: f64(np.stack([np.broadcast_to(np.asarray(P[k], dtype=np.float64), (N,)) for k in names], 1))  # noqa: E731
    anx = bcw = bct = exq = bpq = vtw = vt = None
    if O.COMP_ANXIETY in kinds:
        P = dict(P)
        P.setdefault("p4", np.zeros(N))
        anx = col("W_anx", "p1", "p2", "p3", "p4")
    if O.COMP_BCOST in kinds:
        bcw = f64(np.broadcast_to(P["W_bcost"], (N,)))
        if bcost_tables is None:
            k, gm = np.broadcast_to(P["k"], (N,)), np.broadcast_to(P["gamma"], (N,))
            bcost_tables = np.stack([O.bcost_table(units, k[m], gm[m]) for m in range(N)])
        bct = f64(bcost_tables)
    if O.COMP_EXPERIENCE in kinds:
        exq = col("W_exp", "xi", "kappa_home", "kappa_explo", "theta_explo", "theta_home")
    if O.COMP_BPERSISTENCE in kinds:
        bpq = col("W_bper", "bp", "Wm", "Wnm")
    if O.COMP_VTABLE in kinds:
        vtw = f64(np.broadcast_to(P["W_values"], (N,)))
        vt = f64(vtable_norm).reshape(-1)
    kinds_a = np.ascontiguousarray(kinds, dtype=np.int32)
    tiles = np.ascontiguousarray(maze.tiles, dtype=np.uint8)
    pose = f64(init_pose).reshape(N, 3)
    disc = f64(np.broadcast_to(np.asarray(init_disc, dtype=np.float64), (N, 2)))
    beta = f64(np.broadcast_to(P["beta"], (N,)))
    out = {"hist_pos": np.empty((N, T + 1, 2)) if history else None, "hist_ori": np.empty((N, T + 1)) if history else None,
           "choice": np.empty((N, T), dtype=np.uint8), "occupancy": np.empty((N, maze.X, maze.Y), dtype=np.int32),
           "status": np.empty(N, dtype=np.int32), "final_pose": np.empty((N, 3))}
    rc = h.orc_simulate(_ptr(tiles), maze.X, maze.Y, float(maze.st), _ptr(actions), _ptr(zero_act), A, N, T,
                        _ptr(kinds_a), len(kinds_a), _ptr(beta), _ptr(anx), _ptr(bcw), _ptr(bct), _ptr(exq), _ptr(bpq),
                        _ptr(vtw), _ptr(vt), _ptr(pose), _ptr(disc), _ptr(draws), _ptr(out["hist_pos"]),
                        _ptr(out["hist_ori"]), _ptr(out["choice"]), _ptr(out["occupancy"]), _ptr(out["status"]),
                        _ptr(out["final_pose"]), int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_simulate failed ({rc})")
    return out
