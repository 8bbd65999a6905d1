"""Reference import harness -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Makes the *unmodified* reference importable in the build container:

* ``/root/reference/src`` is put on ``sys.path`` (read-only, never copied);
* ``matplotlib`` / ``seaborn`` / ``progressbar`` -- declared by the reference
  (``pyproject.toml:17-23``) but not installed here -- are replaced by empty stub modules
  (only plotting code touches them: ``maze.py:3,61,104``, ``mouse.py:3``,
  ``visualization/plotting.py:3-4``);
* the pybind11/Eigen extension ``_navigation_model`` cannot be compiled here (Eigen is fetched from
  the network by ``setup.py:15-24``), so an Eigen-free restatement of
  ``src/_navigation_model/mouse.cpp:5-92`` is registered under that module name.  It follows the C++
  operation for operation (libm ``cos``/``sin``/``atan2`` through :mod:`math`, Eigen ``isApprox`` with
  the default double precision 1e-12) and is pinned by the reference's own ``test/test_mouse.py``.

Only ``tests/golden/make_golden.py`` and tests that are skipped when ``/root/reference`` is absent
use this module.  ``/root/reference`` does not exist on the GPU box.
"""
import math
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("NM_REFERENCE_ROOT", "/root/reference")
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "navigation_model"))


class _RefMouse:
    """Eigen-free restatement of the reference's C++ ``Mouse`` (mouse.cpp:5-92, bindings.cpp:11-62)."""

    def __init__(self, init_pos, init_ori, possible_actions, save_history=True):
        acts = np.array(possible_actions, dtype=np.float64)
        if acts.size > 0 and (acts.ndim != 2 or acts.shape[1] != 2):
            raise ValueError("The list of possible actions must have shape (N, 2)")  # mouse.cpp:6-9
        if acts.size == 0:
            acts = acts.reshape(0, 2)
        self._init_pos = np.array(init_pos, dtype=np.float64).reshape(2)
        self._pos = self._init_pos.copy()
        self._init_ori = float(init_ori)
        self._ori = float(init_ori)
        self._actions = acts
        self._save = bool(save_history)
        self.reset_history()

    # read-only properties return copies (Eigen by value -> numpy), bindings.cpp:20-33
    position = property(lambda self: self._pos.copy())
    initial_position = property(lambda self: self._init_pos.copy())
    orientation = property(lambda self: self._ori)
    initial_orientation = property(lambda self: self._init_ori)
    history_position = property(lambda self: [p.copy() for p in self._hpos])
    history_orientation = property(lambda self: list(self._hori))

    def enable_history(self, enable=True):  # mouse.cpp:38-41
        self._save = bool(enable)
        self.reset_history()

    def reset(self, initial_pos=None, initial_ori=None):  # mouse.cpp:43-53
        if initial_pos is not None:
            self._init_pos = np.array(initial_pos, dtype=np.float64).reshape(2)
        if initial_ori is not None:
            self._init_ori = float(initial_ori)
        self._pos = self._init_pos.copy()
        self._ori = self._init_ori
        self.reset_history()

    def reset_history(self):  # mouse.cpp:55-62
        self._hpos = []
        self._hori = []
        if self._save:
            self._hpos.append(self._init_pos.copy())
            self._hori.append(self._init_ori)

    def get_endpoints(self):  # mouse.cpp:64-73: rot * a + pos, row by row
        c = math.cos(self._ori)
        s = math.sin(self._ori)
        ends = np.zeros((self._actions.shape[0], 2))
        for i in range(self._actions.shape[0]):
            ax = float(self._actions[i, 0])
            ay = float(self._actions[i, 1])
            # Eigen evaluates the 2x2 product as c*ax + (-s)*ay and s*ax + c*ay, then adds pos
            ends[i, 0] = (c * ax + (-s) * ay) + self._pos[0]
            ends[i, 1] = (s * ax + c * ay) + self._pos[1]
        return ends

    def move_to(self, x, y=None):  # mouse.cpp:75-92
        if y is None:
            new = np.array(x, dtype=np.float64).reshape(2)
        else:
            new = np.array([x, y], dtype=np.float64)
        d0 = new[0] - self._pos[0]
        d1 = new[1] - self._pos[1]
        # Eigen isApprox: |a-b|^2 <= prec^2 * min(|a|^2, |b|^2), prec = 1e-12
        approx = (d0 * d0 + d1 * d1) <= (1e-12 * 1e-12) * min(new[0] * new[0] + new[1] * new[1],
                                                              self._pos[0] * self._pos[0]
                                                              + self._pos[1] * self._pos[1])
        moving = False
        if not approx:
            self._ori = math.atan2(new[1] - self._pos[1], new[0] - self._pos[0])
            self._pos = new.copy()
            moving = True
        if self._save:
            self._hpos.append(self._pos.copy())
            self._hori.append(self._ori)
        return moving


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, item):
        return _Anything()


def install():
    """Register stubs and put the reference on ``sys.path``.  Idempotent.  Returns True on success."""
    if not available():
        return False
    if "navigation_model" in sys.modules and getattr(sys.modules["navigation_model"], "__nm_ref__", False):
        return True
    for name in ("matplotlib", "seaborn", "progressbar"):
        try:
            __import__(name)
        except Exception:
            if name == "matplotlib":
                mpl = _stub("matplotlib")
                mpl.colors = _stub("matplotlib.colors", ListedColormap=_Anything, Normalize=_Anything)
                mpl.patches = _stub("matplotlib.patches", Polygon=_Anything)
                mpl.pyplot = _stub("matplotlib.pyplot")
                mpl.animation = _stub("matplotlib.animation", FuncAnimation=_Anything, FFMpegWriter=_Anything)
                mpl.collections = _stub("matplotlib.collections", LineCollection=_Anything)
                mpl.cm = _stub("matplotlib.cm")
            else:
                _stub(name, ProgressBar=_Anything)
    if "_navigation_model" not in sys.modules:
        _stub("_navigation_model", Mouse=_RefMouse)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import navigation_model  # noqa: F401  (the reference package)
    navigation_model.__nm_ref__ = True
    return True
