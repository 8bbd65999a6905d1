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
