#!/usr/bin/env python
"""Summarise an `ncu --csv` log (one row per kernel launch x metric) per kernel name.

    python tools/ncu_summary.py gpurun_out/launches.csv [--last-frac 0.25]

Prints launches, total/avg duration, share of the step and the average of every other metric captured.
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    frac = 1.0
    if "--last-frac" in sys.argv:
        frac = float(sys.argv[sys.argv.index("--last-frac") + 1])
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    idi, ki, mi, ui, vi = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        u = r[ui]
        if r[mi].startswith("gpu__time_duration"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        launches.setdefault(int(r[idi]), {"name": r[ki].split("(")[0]})[r[mi]] = v
    ids = sorted(launches)
    ids = ids[int(len(ids) * (1 - frac)):]
    agg = collections.OrderedDict()
    for i in ids:
        L = launches[i]
        a = agg.setdefault(L["name"], collections.defaultdict(float))
        a["n"] += 1
        for k, v in L.items():
            if k != "name":
                a[k] += v
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    short = lambda m: m.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "") \
        .replace("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%") \
        .replace("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%") \
        .replace("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes") \
        .replace("launch__registers_per_thread", "regs").replace("dram__bytes_read.sum", "dram_rd") \
        .replace("dram__bytes_write.sum", "dram_wr").replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%")
    if "--traffic-json" in sys.argv:
        # DRAM bytes (read + write) of one step per bench phase -> profiles/traffic.json
        import json
        steps = float(sys.argv[sys.argv.index("--traffic-json") + 2]) if len(sys.argv) > sys.argv.index("--traffic-json") + 2 else 1.0
        phase_of = {"k_cull": "cull", "k_roots": "roots", "k_fast": "roots", "k_exact": "roots", "k_contact": "contact", "k_emit": "contact", "k_traverse": "traverse", "k_refit": "refit",
                    "k_reduce_points": "reduce", "k_scatter": "reduce", "DeviceScan": "reduce", "k_reset_dirty": "reduce",
                    "k_avg_velocity": "avgvel", "k_boundary": "finalize", "k_final_position": "finalize",
                    "k_morton": "build", "k_hierarchy": "build", "k_scene_bounds": "build", "DeviceRadixSort": "build"}
        out = {}
        for name, a in agg.items():
            for key, ph in phase_of.items():
                if key in name:
                    out[ph] = out.get(ph, 0) + (a.get("dram__bytes_read.sum", 0) + a.get("dram__bytes_write.sum", 0)) / steps
                    break
        json.dump({k: int(v) for k, v in out.items()}, open(sys.argv[sys.argv.index("--traffic-json") + 1], "w"), indent=1)
    print(f"{len(ids)} launches, {tot:.3f} ms total")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = a["n"]
        t = a["gpu__time_duration.sum"]
        extra = "  ".join(f"{short(k)}={a[k] / n:.3g}" for k in a if k not in ("n", "gpu__time_duration.sum"))
        print(f"{t:9.3f} ms {100 * t / tot:5.1f}%  {int(n):4d}x  avg {t / n:8.4f} ms  {name[:44]:44s} {extra}")


if __name__ == "__main__":
    main()
