"""The mouse: continuous pose ``(x, y, orientation)`` plus its relative actions.

Python face of the native ``nm_mouse_*`` object (``csrc/nm_mouse.cpp``), which replaces the
reference's pybind11/Eigen extension ``_navigation_model.Mouse`` (``src/_navigation_model/
bindings.cpp:11-62``, ``mouse.cpp:5-92``) and the thin subclass ``simulation/mouse.py:9-16``.  Same
constructor, read-only properties, methods and error behaviour (``ValueError`` when the action list is
not ``(N, 2)``, ``mouse.cpp:5-9``); getters return copies, as Eigen-by-value does.  The matplotlib
``plot`` helpers (``simulation/mouse.py:18-47``) are out of scope.
"""
import ctypes as C

import numpy as np

from .. import _native


class Mouse:

    def __init__(self, init_pos, init_ori, possible_actions, save_history=True):
        pos = np.ascontiguousarray(np.asarray(init_pos, dtype=np.float64).reshape(-1))
        if pos.size != 2:
            raise ValueError("init_pos must have 2 coordinates")
        acts = np.asarray(possible_actions, dtype=np.float64)
        if acts.size == 0:
            acts = np.zeros((0, 2))
        if acts.ndim != 2:
            raise ValueError("The list of possible actions must have shape (N, 2)")
        acts = np.ascontiguousarray(acts)
        lib = _native.lib()
        self._lib = lib
        self._h = _native.check_handle(lib.nm_mouse_create(_native.np_ptr(pos), float(init_ori), _native.np_ptr(acts),
                                                           acts.shape[0], acts.shape[1], int(bool(save_history))))
        self._n_actions = acts.shape[0]
        self._actions = acts.copy()

    @property
    def possible_actions(self):
        """The ``[A, 2]`` relative actions the mouse was built with (copy)."""
        return self._actions.copy()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.nm_mouse_free(h)

    def _state(self):
        out = np.empty(6)
        _native.check(self._lib.nm_mouse_state(self._h, _native.np_ptr(out)))
        return out

    @property
    def position(self):
        return self._state()[0:2].copy()

    @property
    def orientation(self):
        return float(self._state()[2])

    @property
    def initial_position(self):
        return self._state()[3:5].copy()

    @property
    def initial_orientation(self):
        return float(self._state()[5])

    def _history(self):
        n = int(_native.check(self._lib.nm_mouse_history_len(self._h)))
        pos = np.empty((n, 2))
        ori = np.empty(n)
        _native.check(self._lib.nm_mouse_history(self._h, _native.np_ptr(pos), _native.np_ptr(ori)))
        return pos, ori

    @property
    def history_position(self):
        return list(self._history()[0])

    @property
    def history_orientation(self):
        return [float(o) for o in self._history()[1]]

    @property
    def history(self):
        """``(history_position, history_orientation)`` (``simulation/mouse.py:14-16``)."""
        pos, ori = self._history()
        return list(pos), [float(o) for o in ori]

    def history_arrays(self):
        """History as dense arrays ``(float64[n, 2], float64[n])``."""
        return self._history()

    def enable_history(self, enable=True):
        _native.check(self._lib.nm_mouse_enable_history(self._h, int(bool(enable))))

    def reset_history(self):
        _native.check(self._lib.nm_mouse_reset_history(self._h))

    def reset(self, initial_pos=None, initial_ori=None):
        pos = None if initial_pos is None else np.ascontiguousarray(np.asarray(initial_pos, dtype=np.float64).reshape(2))
        ori = None if initial_ori is None else C.byref(C.c_double(float(initial_ori)))
        _native.check(self._lib.nm_mouse_reset(self._h, _native.np_ptr(pos), ori))

    def get_endpoints(self):
        out = np.empty((self._n_actions, 2))
        _native.check(self._lib.nm_mouse_endpoints(self._h, _native.np_ptr(out)))
        return out

    def move_to(self, x, y=None):
        if y is None:
            x, y = x[0], x[1]
        return bool(_native.check(self._lib.nm_mouse_move_to(self._h, float(x), float(y))))
