#!/usr/bin/env python
"""Per-launch hardware counters of the bench kernels -> profiles/ncu_counts.json (read by bench.py's roofline blocks).

On the GPU box (one gpurun call; ncu runs count as one):

    for k in $(python tools/ncu_counts.py keys); do
      python tools/ncu_counts.py target $k && \
      ncu --metrics $(python tools/ncu_counts.py metrics) --clock-control none $(python tools/ncu_counts.py select $k) \
          --csv --log-file gpurun_out/ncu_$k.csv python tools/ncu_counts.py target $k; done

Back in the build container:

    python tools/ncu_counts.py parse        # gpurun_out/ncu_<key>.csv + .units.json -> profiles/ncu_counts.json,
                                            # raw csv copied to profiles/r02_ncu_<key>.csv

Every entry carries the hash of the kernel's sources (bench.source_sha): bench.py ignores an entry whose hash no longer
matches the tree, so a changed kernel can never be paired with old counters.  Instruction and byte counts of these
deterministic kernels on seeded workloads do not depend on the run; bench.py scales them per unit of work and divides by
its own live time.  A `target` run launches the kernel on the bench workload (or a scaled-down population of the same
distribution, stated in the entry) exactly as bench.py does."""
import csv
import datetime
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = ["sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum",
           "lts__t_bytes.sum.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__time_duration.sum",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
           "sm__icc_request_hit_rate.pct", "launch__registers_per_thread", "sass__inst_executed_local_loads",
           "sass__inst_executed_local_stores", "smsp__thread_inst_executed_per_inst_executed.ratio"]

# key -> (kernel-name regex, launches of that kernel to skip (warm-up), source group of bench.KERNEL_SOURCES)
KEYS = {
    "fit_group_positive_configs1": ("nm_fit_group_kernel", 1, "fit"),
    "fit_global_positive_configs1": ("nm_fit_global_kernel", 1, "fit"),
    # both instantiations are launched by every sweep (the one that does not apply returns at once): select by the
    # DEMANGLED name, which carries the template argument
    "mb_arbitrary_config3": ("nm_mb_kernel<.bool.0>", 1, "mb"),
    "mb_grid_config3": ("nm_mb_kernel<.bool.1>", 1, "mb"),
    "sim_config4": ("nm_sim_kernel", 1, "sim"),
    "q_group_configs1": ("nm_qgroup_kernel", 1, "q"),
}


def target(key):
    import torch
    import bench
    from navigation_model_b200 import workloads as W
    from navigation_model_b200.sweep import ConditioningSweep, ModelBasedSweep, QLearningSweep, parameter_grid
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    maze = W.maze_m9()
    what = ""
    if key.startswith("fit_"):
        traj, rew, a, b, g = bench.make_workload(1)
        if "global" in key:
            perm = np.random.default_rng(7).permutation(len(a))
            a, b, g = a[perm], b[perm], g[perm]
        sw = ConditioningSweep(maze, "positive", possible_actions=W.a8(), device=0)
        sw.add_session(traj, rew)
        a_t, b_t, g_t = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (a, b, g))
        for _ in range(2):
            sw.log_likelihood_device(a_t, b_t, g_t)
        torch.cuda.synchronize()
        units = len(a) * sw.updates_per_parameter_set
        what = "bench headline workload (configs[1] grid x session 0)" + (", random order" if "global" in key else "")
    elif key.startswith("mb_"):
        maze20 = W.maze_m20()
        traj, rew = W.synthetic_session(maze20, 0, 1000, start=(3, 3), goal=(15, 12), jitter=0.02)
        mb = ModelBasedSweep(maze20, W.MOVES8, n_sweeps=1, device=0)
        mb.add_session(traj, rew)
        rng = np.random.default_rng(1000)
        if "grid" in key:
            a, g, b = parameter_grid(np.linspace(0.01, 1, 56), np.linspace(0.5, 0.99, 56), np.logspace(-1, 1.5, 32))
            what = "56 x 56 x 32 beta-fastest grid x 999 steps, 400-state maze (bench: same grid x 4999 steps)"
        else:
            Pm = 20000
            a, b, g = rng.uniform(0.01, 1, Pm), np.exp(rng.uniform(np.log(0.1), np.log(30), Pm)), rng.uniform(0.5, 0.99, Pm)
            what = "20000 arbitrary sets x 999 steps, 400-state maze (bench: 100000 sets x 4999 steps, same distribution)"
        for _ in range(2):
            mb.log_likelihood(a, b, g)
        units = len(a) * mb.updates_per_parameter_set
    elif key == "q_group_configs1":
        traj, rew = W.synthetic_session(maze, 0, bench.T_STEPS)
        g, a, b = parameter_grid(np.linspace(0.5, 0.99, 32), np.linspace(0.01, 1, 64), np.logspace(-1, 1.5, 64))
        ql = QLearningSweep(maze, W.MOVES8, device=0)
        ql.add_session(traj, rew)
        for _ in range(2):
            ql.log_likelihood(a, b, g)
        units = len(a) * ql.updates_per_parameter_set
        what = "bench Q-table workload (configs[1] grid x session 0)"
    elif key == "sim_config4":
        from navigation_model_b200.batched_sim import BatchedExperiment
        from navigation_model_b200.simulation import behavioural_components as bc
        rng = np.random.default_rng(1000)
        N, T = 148 * 6 * 128, 1000
        P = bench.sim_population(rng, N)
        exp = BatchedExperiment(maze, W.a8(), W.A8_UNITS,
                                [bc.AnxietyComponent, bc.BiomechanicalCostComponent, bc.ExperienceComponent,
                                 bc.BiomechPersistenceComponent, bc.ValueTableComponent], v_table=rng.random((9, 9)),
                                device=0)
        pos0 = np.array(maze.get_tile_centre(1, 1)) + rng.uniform(-0.03, 0.03, (N, 2))
        ori0 = rng.uniform(-3, 3, N)
        for rep in range(2):
            exp.run(P, T, pos0, ori0, init_disc=(1, 1), device_seed=77 + rep, n_mice=N, occupancy_total=True,
                    shock_zone={"tile_lines": 5}, corridors=(1, 7), static_intervals=True, subareas=True,
                    orientation_histogram={})
        units = N * T
        what = f"{N} mice x {T} steps of the bench population (bench: 1e6 mice x 1e4 steps, same distribution)"
    else:
        raise SystemExit(f"unknown key {key}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"units_per_launch": int(units), "workload": what},
              open(os.path.join(ROOT, "gpurun_out", f"ncu_{key}.units.json"), "w"))


def parse():
    import bench
    out_path = os.path.join(ROOT, "profiles", "ncu_counts.json")
    try:
        out = json.load(open(out_path))
    except Exception:
        out = {}
    for key, (regex, skip, kind) in KEYS.items():
        path = os.path.join(ROOT, "gpurun_out", f"ncu_{key}.csv")
        upath = os.path.join(ROOT, "gpurun_out", f"ncu_{key}.units.json")
        if not (os.path.exists(path) and os.path.exists(upath)):
            continue
        rows = [r for r in csv.reader(open(path)) if len(r) > 10]
        if not any("Metric Name" in r for r in rows):
            print(key, "-- no kernel was profiled (see", path + ")")
            continue
        hdr = next(r for r in rows if "Metric Name" in r)
        iname, ival, iunit, ik = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
        entry = {"key": key, "source_sha": bench.source_sha(kind),
                 "captured": datetime.datetime.fromtimestamp(os.path.getmtime(path)).strftime("%Y-%m-%d %H:%M"),
                 "raw": f"profiles/r02_ncu_{key}.csv"}
        entry.update(json.load(open(upath)))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6,
                 "msecond": 1e-3, "second": 1.0}
        for r in rows:
            if r is hdr or r[iname] not in METRICS:
                continue
            entry["kernel"] = r[ik][:100]
            try:
                v = float(r[ival].replace(",", ""))
            except ValueError:
                continue
            entry[r[iname]] = v * scale.get(r[iunit], 1.0)
        out[key] = entry
        shutil.copy(path, os.path.join(ROOT, "profiles", f"r02_ncu_{key}.csv"))
        print(key, {k: entry.get(k) for k in ("sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
                                             "gpu__time_duration.sum", "units_per_launch")})
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "keys"
    if cmd == "keys":
        print(" ".join(KEYS))
    elif cmd == "metrics":
        print(",".join(METRICS))
    elif cmd == "select":
        regex, skip, _ = KEYS[sys.argv[2]]
        base = "--kernel-name-base demangled " if "<" in regex else ""
        print(f"{base}-k regex:{regex} -s {skip} -c 1")  # substituted unquoted: the shell does not re-parse < > there
    elif cmd == "target":
        target(sys.argv[2])
    elif cmd == "parse":
        parse()
