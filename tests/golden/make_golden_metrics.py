"""Golden fixtures of the trajectory metrics, produced by the UNMODIFIED reference.

    python tests/golden/make_golden_metrics.py        (build container only: needs /root/reference)

Takes the six reference trajectories already stored in sim_golden.npz (positions + orientations of
StandardExperiment.run) and stores what the reference's own analysis/metrics.py functions return for them:
orientation histograms (default bins, bins from the action set, six bins with distance=True, filtered by tile
type) and their signatures, the moving / not-moving run lengths with mov_bouts and compute_static_intervals,
the subarea histogram, the shock-zone metrics under both zone descriptions, the corridor switches and the
tile-type histogram.  Nothing from this repo's product code or oracle computes a stored value."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refharness  # noqa: E402

if not refharness.install():
    raise SystemExit("reference not available at /root/reference")

from navigation_model.analysis import metrics as RM  # noqa: E402  (the reference)
from navigation_model.simulation.maze import Maze  # noqa: E402

M9_FILE = os.path.join(refharness.REFERENCE_ROOT, "test", "data", "maze_layout.txt")
SIZE_TILE = 0.111


def main():
    g = dict(np.load(os.path.join(HERE, "sim_golden.npz"), allow_pickle=False))
    maze = Maze.from_file(M9_FILE, size_tile=SIZE_TILE)
    T = maze.MazeTile
    units = [tuple(int(v) for v in a) for a in g["actions_units"]]
    out = {k: [] for k in ("ori_hist8", "ori_hist_actions", "ori_hist6_distance", "ori_hist_filtered", "ori_signature",
                           "run_lengths", "run_kinds", "n_runs", "bouts", "static_intervals", "subareas_hist",
                           "latency_sub12", "entries_sub12", "latency_sub3", "entries_sub3", "latency_lines5",
                           "entries_lines5", "occupancy_lines5", "occupancy_lines2", "corridor_switches",
                           "tiles_histogram", "n_relative")}
    max_runs = 0
    runs_all = []
    for m in range(len(g["seeds"])):
        pos, ori = g["hist_pos"][m], list(g["hist_ori"][m])
        disc = RM.discretize_positions(pos, maze)
        h8, rel, edges8 = RM.hist_orientations(ori, disc)
        ha, _, edges_a = RM.hist_orientations(ori, disc, possible_actions=units)
        h6, _, edges6 = RM.hist_orientations(ori, disc, n_bins=6, distance=True)
        hf, _, _ = RM.hist_orientations_filtered(ori, disc, maze, [T.CORNER, T.WALL])
        out["ori_hist8"].append(h8)
        out["ori_hist_actions"].append(ha)
        out["ori_hist6_distance"].append(h6)
        out["ori_hist_filtered"].append(hf)
        out["n_relative"].append(len(rel))
        out["ori_signature"].append(RM.relative_orientation_signatures(h8)[1])
        runs = RM.hist_time_moving(RM.motion_analysis(disc)[1])
        runs_all.append(runs)
        max_runs = max(max_runs, len(runs))
        out["n_runs"].append(len(runs))
        out["bouts"].append(RM.mov_bouts(runs)[2])
        out["static_intervals"].append(RM.compute_static_intervals(runs))
        out["subareas_hist"].append(RM.subareas(disc, maze))
        out["latency_sub12"].append(RM.latency_enter_shock_zone(disc, maze, subareas_shock_zone=(1, 2)))
        out["entries_sub12"].append(RM.shock_zone_entries(disc, maze, subareas_shock_zone=(1, 2)))
        out["latency_sub3"].append(RM.latency_enter_shock_zone(disc, maze, subareas_shock_zone=3))
        out["entries_sub3"].append(RM.shock_zone_entries(disc, maze, subareas_shock_zone=3))
        out["latency_lines5"].append(RM.latency_enter_shock_zone(disc, maze, tile_lines=5))
        out["entries_lines5"].append(RM.shock_zone_entries(disc, maze, tile_lines=5))
        out["occupancy_lines5"].append(RM.occupancy_shock_zone(disc, tile_lines=5))
        out["occupancy_lines2"].append(RM.occupancy_shock_zone(disc, tile_lines=2))
        out["corridor_switches"].append(RM.arms(disc, maze))
        out["tiles_histogram"].append(RM.histogram_tiles(disc, maze))
    for runs in runs_all:
        pad = max_runs - len(runs)
        out["run_lengths"].append([r[0] for r in runs] + [0] * pad)
        out["run_kinds"].append([r[1] for r in runs] + [-1] * pad)
    arrays = {k: np.array(v) for k, v in out.items()}
    # long synthetic tile walks (the simulated mice above never commute between the two corridors): random
    # 8-neighbour walks over the visitable tiles, 3000 positions each
    rng = np.random.default_rng(77)
    walks, w_out = [], {k: [] for k in ("walk_switches", "walk_entries_sub12", "walk_entries_lines5",
                                        "walk_latency_sub3", "walk_latency_lines5", "walk_occupancy_lines5",
                                        "walk_subareas_hist", "walk_static_intervals", "walk_bouts")}
    for w in range(4):
        cell = [(1, 1), (7, 1), (4, 8), (0, 8)][w]
        walk = [cell]
        for _ in range(2999):
            if rng.random() < 0.3:
                walk.append(cell)
                continue
            while True:
                d = rng.integers(-1, 2, 2)
                nxt = (int(cell[0] + d[0]), int(cell[1] + d[1]))
                if maze.is_visitable(nxt):
                    cell = nxt
                    break
            walk.append(cell)
        walks.append(walk)
        runs = RM.hist_time_moving(RM.motion_analysis(walk)[1])
        w_out["walk_switches"].append(RM.arms(walk, maze))
        w_out["walk_entries_sub12"].append(RM.shock_zone_entries(walk, maze, subareas_shock_zone=(1, 2)))
        w_out["walk_entries_lines5"].append(RM.shock_zone_entries(walk, maze, tile_lines=5))
        w_out["walk_latency_sub3"].append(RM.latency_enter_shock_zone(walk, maze, subareas_shock_zone=3))
        w_out["walk_latency_lines5"].append(RM.latency_enter_shock_zone(walk, maze, tile_lines=5))
        w_out["walk_occupancy_lines5"].append(RM.occupancy_shock_zone(walk, tile_lines=5))
        w_out["walk_subareas_hist"].append(RM.subareas(walk, maze))
        w_out["walk_static_intervals"].append(RM.compute_static_intervals(runs))
        w_out["walk_bouts"].append(RM.mov_bouts(runs)[2])
    arrays["walks"] = np.array(walks)
    arrays.update({k: np.array(v) for k, v in w_out.items()})
    arrays.update(edges8=np.array(edges8), edges_actions=np.array(edges_a), edges6=np.array(edges6))
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), **arrays)
    print("metrics_golden.npz:", {k: v.shape for k, v in arrays.items()})
    print({k: arrays[k].tolist() for k in ("ori_hist8", "ori_hist6_distance", "entries_lines5", "corridor_switches",
                                           "static_intervals", "latency_sub12", "walk_switches",
                                           "walk_entries_lines5", "walk_latency_sub3")})


if __name__ == "__main__":
    main()
