"""navigation_model_b200 -- B200-native hot path of elimas9/navigation_model.

Module layout mirrors the reference package (``navigation_model.simulation.{maze,mouse,rl,
exploration_model,behavioural_components,experiment}``, ``.analysis.{data,metrics}``,
``.optimization.fitness``) so that ``import navigation_model_b200 as navigation_model`` (or
:func:`install_as`) is a drop-in for the hot path; the batched, GPU-backed entry points live in
:mod:`navigation_model_b200.sweep` and :mod:`navigation_model_b200.batched_sim`.
"""
import importlib
import sys

__version__ = "0.1.0"


def install_as(name="navigation_model"):
    """Alias this package (and its sub-packages) under ``name`` in ``sys.modules``."""
    me = sys.modules[__name__]
    sys.modules[name] = me
    for sub in ("simulation", "analysis", "optimization"):
        mod = importlib.import_module(f"{__name__}.{sub}")
        sys.modules[f"{name}.{sub}"] = mod
    return me
