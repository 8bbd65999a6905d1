"""Golden distances produced by the UNMODIFIED reference (analysis/statistics.py:10-54, 67-103):

    python tests/golden/make_golden_distances.py

normalized_distance, manhattan_distance and js_distance of 300 seeded integer histograms of 81 bins (empty, identical,
single-bin and all-zero-data cases included) against one data histogram -> distances_golden.npz.  Pins the oracle's
restated formulas (tests/test_metrics_golden.py) and, through them or directly, the device kernel nm_distance_rows
(tests/test_fitness_gpu.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import refharness  # noqa: E402

if not refharness.install():
    raise SystemExit("reference not available at /root/reference")

from navigation_model.analysis.statistics import js_distance, manhattan_distance, normalized_distance  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    N, L = 300, 81
    data = rng.integers(0, 50, L).astype(np.float64)
    x = rng.integers(0, 50, (N, L)).astype(np.float64)
    x[0] = 0.0
    x[1] = data
    x[2, :] = 0.0
    x[2, 5] = 7.0
    zero = np.zeros(L)
    out = {"data": data, "x": x,
           "nd": np.array([normalized_distance(x[n], data) for n in range(N)], dtype=np.float64),
           "l1": np.array([manhattan_distance(x[n], data) for n in range(N)], dtype=np.float64),
           "js": np.array([js_distance((x[n],), (data,)) for n in range(N)], dtype=np.float64),
           "nd_zero_data": np.array([normalized_distance(x[n], zero) for n in range(N)], dtype=np.float64)}
    np.savez_compressed(os.path.join(HERE, "distances_golden.npz"), **out)
    print({k: (np.shape(v), float(np.min(v)), float(np.max(v))) for k, v in out.items()})


if __name__ == "__main__":
    main()
