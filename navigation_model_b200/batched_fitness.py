"""Batched multi-objective fitness: distances between the fused metrics of N simulated mice and data metrics,
computed on the device.

Grid version of the reference's ``MultiObjectiveFitness`` / ``ResultsDistance`` (``optimization/fitness.py:4-40``,
``analysis/statistics.py:106-137``): objectives are named, each compares one key of the results against the same key
of ``compare_against`` with one of the reference's distance functions (``normalized_distance``, ``manhattan_distance``,
``js_distance``; ``statistics.py:10-103``).  Results are ``[N, L]`` (or ``[N]``) arrays with one row per simulated
mouse -- e.g. ``SimResult.occupancy`` flattened, ``tile_counts``, ``count_moving`` -- and the distance is evaluated for
every row by ``nm_distance_rows``: a sweep over simulation parameters reduces simulation -> metrics -> distance without
per-mouse histograms leaving the GPU.
"""
import numpy as np

from . import _native

_KIND = {"normalized": _native.DIST_NORMALIZED, "normalized_distance": _native.DIST_NORMALIZED,
         "manhattan": _native.DIST_MANHATTAN, "manhattan_distance": _native.DIST_MANHATTAN,
         "js": _native.DIST_JS, "js_distance": _native.DIST_JS}


class BatchedMultiObjectiveFitness:
    """A fitness with many named objectives, evaluated for every row (mouse) at once."""

    def __init__(self, compare_against, device=None):
        """:param compare_against: ``{key: data metric}`` (scalars or 1-D / N-D arrays, flattened)"""
        self._compare_against = compare_against
        self._objectives = {}
        self._device = device

    def add_objective(self, name, distance, key):
        """:param distance: ``"normalized"``, ``"manhattan"`` or ``"js"`` (or the reference function itself)"""
        if name in self._objectives:
            raise RuntimeError(f"Objective {name} already in fitness")
        kind = _KIND.get(getattr(distance, "__name__", distance))
        if kind is None:
            raise ValueError(f"no device implementation of distance {distance!r}")
        if key not in self._compare_against:
            raise KeyError(key)
        self._objectives[name] = (kind, key)

    def __call__(self, results):
        """:param results: ``{key: array or CUDA tensor [N, ...]}``  ->  ``{objective: float64[N]}``"""
        import torch
        _native.require_device()
        dev = torch.device("cuda", torch.cuda.current_device() if self._device is None else int(self._device))
        lib = _native.lib()
        out = {}
        with torch.cuda.device(dev):
            for name, (kind, key) in self._objectives.items():
                x = results[key]
                if not isinstance(x, torch.Tensor):
                    x = torch.from_numpy(np.ascontiguousarray(x))
                x = x.to(dev, dtype=torch.float64)
                N = x.shape[0]
                x = x.reshape(N, -1).contiguous()
                ref = torch.from_numpy(np.ascontiguousarray(np.asarray(self._compare_against[key], dtype=np.float64)
                                                            .reshape(-1))).to(dev)
                if ref.numel() != x.shape[1]:
                    raise ValueError(f"objective {name}: result rows have {x.shape[1]} entries, data has {ref.numel()}")
                d = torch.empty(N, dtype=torch.float64, device=dev)
                _native.check(lib.nm_distance_rows(kind, _native.tensor_ptr(x), _native.tensor_ptr(ref), N, x.shape[1],
                                                   _native.tensor_ptr(d), _native.cuda_stream_ptr()))
                out[name] = d
            torch.cuda.current_stream().synchronize()
        return {k: v.cpu().numpy() for k, v in out.items()}

    def best(self, results, weights=None, offset=0, group=None):
        """The row (mouse / simulated parameter set) with the smallest weighted sum of the objectives.

        With an initialised ``torch.distributed`` process group the rows are this rank's block ``[offset, offset + N)``
        of a sharded population (``BatchedExperiment.run_sharded``): the per-rank winners -- 16 bytes each -- are
        all-gathered (``sweep.reduce_best``: NCCL, or gloo for CPU tensors) and every rank returns the global one,
        lowest index on ties.  Rows whose sum is NaN never win.

        :param weights: ``{objective: factor}`` (default 1 for every objective)
        :return: ``(value, global row index)`` -- ``(nan, -1)`` when no row qualifies
        """
        import torch
        import torch.distributed as dist
        from .sweep import reduce_best
        d = self(results)
        total = None
        for name, v in d.items():
            term = v * float(1.0 if weights is None else weights.get(name, 1.0))
            total = term if total is None else total + term
        best = torch.empty(2, dtype=torch.float64)
        if total is None or total.size == 0 or np.all(np.isnan(total)):
            best[0] = float("nan")
            best[1:].view(torch.int64)[0] = -1
        else:
            i = int(np.nanargmin(total))
            best[0] = float(total[i])
            best[1:].view(torch.int64)[0] = i
        if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl":
            best = best.to(torch.device("cuda", torch.cuda.current_device() if self._device is None
                                        else int(self._device)))
        return reduce_best(best, int(offset), group=group)
