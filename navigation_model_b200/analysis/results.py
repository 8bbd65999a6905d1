"""Aggregation over many result dictionaries -- mirror of the reference's ``analysis/results.py`` (``MultiResult``,
``results.py:4-101``; pinned by the reference's ``test/test_results.py``, which runs drop-in against this class).

The reference copies one sub-list per ``__getitem__``; here a view remembers the PATH into the shared list of result
trees and resolves it when a statistic is asked for, and :meth:`MultiResult.from_batch` wraps the batched results of
``BatchedExperiment`` / ``Analyzer`` (leaves with a leading per-mouse axis) without unstacking them."""
import numpy as np


class MultiResult:
    """Statistics over several result dictionaries of the same shape, addressed to any depth::

        mr = MultiResult({"a": {"b": 1}}, {"a": {"b": 2}})
        mr["a"]["b"].data()          # array([1, 2])
        mr["a"]["b"].mean()          # 1.5
    """

    def __init__(self, *results):
        self._results = []  # shared by every view derived from this object
        self._path = ()
        self._stacked = None  # from_batch: one tree whose leaves carry the result axis in front
        for res in results:
            self.append(res)

    @classmethod
    def from_batch(cls, batched):
        """Wrap ONE result tree whose leaves already stack the individual results along axis 0 (what the batched
        simulation / metric paths return); ``data()`` of a leaf is then the leaf itself."""
        out = cls()
        out._stacked = batched
        return out

    def reset(self):
        """Forget every result (``results.py:26-30``)."""
        del self._results[:]
        self._stacked = None

    def append(self, res):
        """Add one result dictionary, typically the output of ``Analyzer.analyze`` (``results.py:32-41``)."""
        if self._path:
            raise TypeError("results are appended to the root MultiResult, not to a view of a sub-entry")
        if self._stacked is not None:
            raise TypeError("a MultiResult built by from_batch holds its results already stacked")
        self._results.append(res)

    def __getitem__(self, item):
        """View of entry ``item`` (key or index) of every result (``results.py:43-53``)."""
        view = MultiResult.__new__(MultiResult)
        view._results = self._results
        view._stacked = self._stacked
        view._path = self._path + (item,)
        self._resolve(view._path, probe=True)  # a wrong key fails here, like the reference's eager copy
        return view

    def __len__(self):
        return len(self.data()) if self._stacked is not None else len(self._results)

    def _resolve(self, path, probe=False):
        def walk(tree):
            for key in path:
                tree = tree[key]
            return tree
        if self._stacked is not None:
            return walk(self._stacked)
        if probe:
            for res in self._results:
                walk(res)
            return None
        return [walk(res) for res in self._results]

    # -- statistics over the result axis: only meaningful on numerical leaves (``results.py:55-93``) -----------------
    def median(self):
        return np.median(self._resolve(self._path), axis=0)

    def mean(self):
        return np.mean(self._resolve(self._path), axis=0)

    def std(self):
        return np.std(self._resolve(self._path), axis=0)

    def percentile(self):
        """``(25th, 75th)`` percentiles."""
        leaf = self._resolve(self._path)
        return np.percentile(leaf, 25, axis=0), np.percentile(leaf, 75, axis=0)

    def data(self):
        """The stacked values (``results.py:95-101``)."""
        return np.array(self._resolve(self._path))
