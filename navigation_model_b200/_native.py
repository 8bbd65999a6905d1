"""ctypes binding of ``libnavmodel_b200.so`` (the C-ABI declared in ``include/navmodel_b200.h``).

There is no CPU fallback: if the library is missing, :func:`lib` raises; compute entry points raise
``RuntimeError`` when no CUDA device is usable.  Error codes are translated to the exception types the
reference raises for the same condition (``ValueError`` / ``RuntimeError``).
"""
import ctypes as C
import os
import threading

import numpy as np

from . import _build

NM_MAX_CAND = 8
NM_OBS_NONE = 0xFF
ALG_TD0, ALG_POSITIVE, ALG_NEGATIVE = 0, 1, 2

#: numpy view of ``nm_step_rec`` (32 bytes)
STEP_REC_DTYPE = np.dtype([("r", "<f8"), ("s", "<u2"), ("s1", "<u2"), ("ncand", "u1"), ("obs", "u1"),
                           ("reserved", "<u2"), ("cand", "<u2", (NM_MAX_CAND,))], align=False)
assert STEP_REC_DTYPE.itemsize == 32

_p = C.c_void_p
_dp = C.POINTER(C.c_double)
_i64 = C.c_int64

#: name -> (restype, argtypes); every symbol declared in include/navmodel_b200.h
SIGNATURES = {
    "nm_version": (C.c_int, []),
    "nm_last_error": (C.c_char_p, []),
    "nm_device_count": (C.c_int, []),
    "nm_mouse_create": (_p, [_p, C.c_double, _p, C.c_int, C.c_int, C.c_int]),
    "nm_mouse_free": (None, [_p]),
    "nm_mouse_state": (C.c_int, [_p, _p]),
    "nm_mouse_num_actions": (C.c_int, [_p]),
    "nm_mouse_endpoints": (C.c_int, [_p, _p]),
    "nm_mouse_move_to": (C.c_int, [_p, C.c_double, C.c_double]),
    "nm_mouse_enable_history": (C.c_int, [_p, C.c_int]),
    "nm_mouse_reset_history": (C.c_int, [_p]),
    "nm_mouse_reset": (C.c_int, [_p, _p, _p]),
    "nm_mouse_history_len": (_i64, [_p]),
    "nm_mouse_history": (C.c_int, [_p, _p, _p]),
    "nm_mouse_adopt_trajectory": (C.c_int, [_p, _p, _p, _i64]),
    "nm_mouse_move_through": (C.c_int, [_p, _p, _i64]),
    "nm_maze_upload": (_p, [_p, _p, C.c_int, C.c_int, C.c_double, C.c_int, _p]),
    "nm_maze_free": (None, [_p]),
    "nm_maze_num_states": (C.c_int, [_p]),
    "nm_maze_cont2disc": (C.c_int, [_p, C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nm_maze_movement_possible_cont": (C.c_int, [_p, C.c_double, C.c_double, C.c_double, C.c_double]),
    "nm_maze_closest_visitable_cont": (C.c_int, [_p, C.c_double, C.c_double, C.POINTER(C.c_int),
                                                 C.POINTER(C.c_int)]),
    "nm_stream_build": (_i64, [_p, _p, _p, _i64, _p, _p]),
    "nm_stream_build_pairs": (_i64, [_p, _p, _p, _p, _p, _p, _i64, _p, _p]),
    "nm_streams_upload": (_p, [_p, _p, _p, C.c_int, C.c_int, C.c_int, _p]),
    "nm_streams_update": (C.c_int, [_p, _p, _p]),
    "nm_streams_free": (None, [_p]),
    "nm_sessions_subsample_device": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_double, _p, _p, _p, _p]),
    "nm_streams_build_device": (_p, [_p, _p, _p, _p, _p, C.c_int, _p, C.c_int, _p, _p]),
    "nm_streams_with_replays": (_p, [_p, _p, _p, _p]),
    "nm_streams_download": (C.c_int, [_p, _p, _p]),
    "nm_sessions_orientations_device": (C.c_int, [_p, _p, _p, _p, C.c_int, _p, C.c_int, _p, _p, _p]),
    "nm_positions_adjust_device": (C.c_int, [_p, _p, _i64, _p, _p]),
    "nm_vtable_sweep": (C.c_int, [C.c_int, _p, _p, _p, _p, _i64, _p, C.c_int, _p, _p]),
    "nm_loglik_sweep": (C.c_int, [C.c_int, _p, _p, _p, _p, _p, _i64, _p, C.c_int, _p, _p, _p]),
    "nm_vtable_max_states": (C.c_int, []),
    "nm_scratch_release": (C.c_int, []),
    "nm_loglik_sweep_mode": (C.c_int, [_p, C.POINTER(C.c_int)]),
    "nm_argmin_nll": (C.c_int, [_p, C.c_int, _i64, _p, _p, _p]),
    "nm_fp64_peak_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, _p, C.POINTER(_i64), _p]),
    "nm_modelbased_sweep": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, _p, _p, _p, _i64, _p, _p, _p, _p, _p]),
    "nm_qtable_sweep": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "nm_modelbased_sweep_mode": (C.c_int, [_p, C.POINTER(C.c_int)]),
    "nm_qtable_sweep_mode": (C.c_int, [_p, C.POINTER(C.c_int)]),
    "nm_distance_rows": (C.c_int, [C.c_int, _p, _p, _i64, C.c_int, _p, _p]),
    "nm_simulate": (C.c_int, [_p, _p, _p]),
    "nm_vtable_trace": (C.c_int, [C.c_int, _p, _p, C.c_int, C.c_double, C.c_double, _p, _p, _p, _p, _p]),
    "nm_vtable_trace_max_states": (C.c_int, []),
    "nm_simulate_fits": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "nm_simulate_scratch_columns": (C.c_int64, [C.c_int64]),
    "nm_sim_args_size": (C.c_int, []),
    "nm_maze_movements_possible_device": (C.c_int, [_p, _p, _i64, _p, _p]),
}

NM_MAX_COMPONENTS = 8
NM_MAX_SUBAREAS = 16
NM_MAX_ORI_BINS = 16
TILE_ZONE_IN, TILE_ZONE_OUT, TILE_CORRIDOR_A, TILE_CORRIDOR_B = 1, 2, 4, 8  # NM_TILE_* (navmodel_b200.h)
COMP_ANXIETY, COMP_BCOST, COMP_EXPERIENCE, COMP_BPERSISTENCE, COMP_VTABLE = 1, 2, 3, 4, 5
DIST_NORMALIZED, DIST_MANHATTAN, DIST_JS = 0, 1, 2


class SimArgs(C.Structure):
    """ctypes mirror of ``nm_sim_args`` (include/navmodel_b200.h)."""
    _fields_ = [
        ("n_mice", C.c_int64), ("n_steps", C.c_int32), ("n_actions", C.c_int32),
        ("h_actions", C.c_void_p), ("h_zero_action", C.c_void_p),
        ("n_components", C.c_int32),
        ("component_kind", C.c_int32 * NM_MAX_COMPONENTS), ("component_active", C.c_int32 * NM_MAX_COMPONENTS),
        ("d_beta", C.c_void_p), ("d_init_pose", C.c_void_p), ("d_init_prev", C.c_void_p),
        ("d_anxiety", C.c_void_p), ("d_bcost_weight", C.c_void_p), ("d_bcost_table", C.c_void_p),
        ("d_experience", C.c_void_p), ("d_bpersistence", C.c_void_p),
        ("d_vtable_weight", C.c_void_p), ("d_vtable", C.c_void_p),
        ("vtable_per_mouse", C.c_int32), ("model_kind", C.c_int32),
        ("d_draws", C.c_void_p), ("d_pcg_state", C.c_void_p),
        ("visit_percentage", C.c_double),
        ("d_final_pose", C.c_void_p), ("d_pcg_state_out", C.c_void_p),
        ("d_hist_pos", C.c_void_p), ("d_hist_ori", C.c_void_p), ("d_hist_choice", C.c_void_p),
        ("d_occupancy", C.c_void_p), ("d_it_visit", C.c_void_p), ("d_count_moving", C.c_void_p),
        ("d_tile_counts", C.c_void_p), ("d_status", C.c_void_p), ("d_occupancy_total", C.c_void_p),
        ("d_model_state", C.c_void_p), ("d_pcg_u32", C.c_void_p), ("d_pcg_u32_out", C.c_void_p),
        ("device_seed", C.c_uint64), ("use_device_seed", C.c_int32), ("reserved1", C.c_int32),
        ("d_tile_flags", C.c_void_p), ("d_zone_metrics", C.c_void_p), ("d_static_intervals", C.c_void_p),
        ("d_subarea_counts", C.c_void_p),
        ("n_ori_bins", C.c_int32), ("ori_distance", C.c_int32), ("ori_tile_filter", C.c_uint32),
        ("reserved2", C.c_int32),
        ("h_ori_edges", C.c_void_p), ("ori_min_lim", C.c_double), ("ori_max_lim", C.c_double),
        ("d_ori_hist", C.c_void_p),
        ("d_init_visits", C.c_void_p), ("d_init_model_state", C.c_void_p),
        ("init_visits_max", C.c_uint32), ("reserved3", C.c_int32),
        ("mouse_index_base", C.c_int64),
        ("d_visit_scratch", C.c_void_p), ("visit_scratch_columns", C.c_int64),
    ]

_lock = threading.Lock()
_lib = None


def library_path():
    """The in-tree library; ``NM_NATIVE_LIB`` points at another build of it (A/B runs of compile-time variants)."""
    return os.environ.get("NM_NATIVE_LIB") or _build.LIB_PATH


def lib():
    """Load (once) and return the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = library_path()
            if not os.path.exists(path):
                raise RuntimeError(
                    f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(navigation_model_b200 has no CPU fallback)")
            handle = C.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            if handle.nm_sim_args_size() != C.sizeof(SimArgs):
                raise RuntimeError(f"{path}: nm_sim_args is {handle.nm_sim_args_size()} bytes, the ctypes mirror "
                                   f"{C.sizeof(SimArgs)}: rebuild the library (__graft_entry__.build())")
            _lib = handle
    return _lib


def last_error():
    msg = lib().nm_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


_EXC = {-1: ValueError, -2: RuntimeError, -3: RuntimeError, -4: MemoryError}


def check(rc):
    """Raise the exception matching a negative return code; pass non-negative values through."""
    if rc is None or rc >= 0:
        return rc
    raise _EXC.get(int(rc), RuntimeError)(last_error())


def check_handle(h):
    if not h:
        msg = last_error()
        raise (ValueError if "must have shape" in msg or "bad argument" in msg or "need " in msg else RuntimeError)(msg)
    return h


def device_count():
    return int(lib().nm_device_count())


def require_device():
    if device_count() <= 0:
        raise RuntimeError("navigation_model_b200: no usable CUDA device (there is no CPU fallback)")


def np_ptr(a):
    """Raw pointer of a C-contiguous numpy array (the array must outlive the call)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def cuda_stream_ptr(stream=None):
    """``cudaStream_t`` (as int) of a torch stream, or of torch's current stream."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


def tensor_ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())
