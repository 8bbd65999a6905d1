"""Distances between metric results (reference ``analysis/statistics.py:10-183``).

SURVEY section 8f rank 1: the arithmetic behind ``MultiObjectiveFitness`` -- tiny per-result host math.
Same function names, argument conventions and clipping rules as the reference.
"""
import math

import numpy as np
from scipy.special import rel_entr


def _normalised_or_zero(h, axis):
    total = np.sum(h, axis=axis, keepdims=True)
    return 0 if total == 0 else h / total


def js_distance(h1, h2, base=None, *, axis=0, keepdims=False):
    """Jensen-Shannon distance of two histograms given as ``[counts, ...]`` (``:10-54``); an all-zero
    histogram is treated as the zero vector, an infinite distance is reported as 1."""
    p = _normalised_or_zero(np.asarray(h1[0]), axis)
    q = _normalised_or_zero(np.asarray(h2[0]), axis)
    m = (p + q) / 2.0
    js = np.sum(rel_entr(p, m), axis=axis, keepdims=keepdims) + np.sum(rel_entr(q, m), axis=axis, keepdims=keepdims)
    if base is not None:
        js /= np.log(base)
    distance = np.sqrt(js / 2.0)
    if math.isinf(distance):
        distance = 1
    return distance


def kl_divergence(h1, h2):
    """Kullback-Leibler divergence (``:56-64``)."""
    return sum(rel_entr(h1, h2))


def normalized_distance(h1, h2):
    """Euclidean distance of the two histograms normalised by their cardinality, clipped to [0, 1]
    (``:67-92``)."""
    c1, c2 = np.sum(h1), np.sum(h2)
    if c1 == 0 and c2 == 0:
        nd = 0
    elif c1 == 0:
        nd = np.linalg.norm(h2 / c2)
    elif c2 == 0:
        nd = np.linalg.norm(h1 / c1)
    else:
        nd = np.linalg.norm(h1 / c1 - h2 / c2)
    return 1 if nd > 1 else nd


def manhattan_distance(x1, x2):
    """L1 distance (``:95-103``)."""
    return np.linalg.norm(np.array(x1) - np.array(x2), ord=1)


class ResultsDistance:
    """Distance of selected keys of a results dict to fixed reference results, with a history
    (``:106-183``)."""

    def __init__(self, distance_function, compare_against, *keys):
        self._distance_function = distance_function
        self._keys = keys
        self._x2 = [compare_against[k] for k in keys]
        self._distances = []

    def __call__(self, results):
        picked = [results[k] for k in self._keys]
        self._distances.append(self._distance_function(picked, self._x2))
        return self._distances[-1]

    def reset(self):
        self._distances = []

    def data(self):
        return self._distances

    def _summary(self, fn, *args):
        return fn(self._distances, *args)

    def median(self):
        return self._summary(np.median)

    def mean(self):
        return self._summary(np.mean)

    def std(self):
        return self._summary(np.std)

    def percentile(self):
        return self._summary(np.percentile, 25), self._summary(np.percentile, 75)
