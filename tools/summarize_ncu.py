#!/usr/bin/env python
"""Turns ncu artefacts brought back in gpurun_out/ into the small text summaries
kept under profiles/.

  summarize_ncu.py launches <launch_list.csv>        per-kernel share table
  summarize_ncu.py full <capture.ncu-rep>            key metrics per captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            ns = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        name = name.split("<")[0]
        c = agg.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += ns
        total += ns
    print("# %s: %d launches, %.1f us of kernel time (ncu: cold cache, "
          "serialised launches; shares only)" % (path, sum(c[0] for c in agg.values()),
                                                 total / 1e3))
    print("%-60s %7s %12s %10s %7s" % ("kernel", "count", "total_us", "avg_us", "share"))
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %7d %12.1f %10.2f %6.1f%%" % (name[:60], cnt, ns / 1e3,
                                                   ns / cnt / 1e3, 100 * ns / total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    print("# %s: %d captured launches (ncu --set full --clock-control none)"
          % (path, len(rows) - 2))
    for r in rows[2:]:
        print("---- " + r[ki][:150])
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print("  %-78s %16s %s" % (k, r[i], units[i]))
        if "dram__bytes_read.sum" in h:
            def val(k):
                i = h.index(k)
                v = float(r[i].replace(",", ""))
                u = units[i]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0,
                            "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
            traffic = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            t = val("gpu__time_duration.sum")
            print("  %-78s %16.1f GB/s (dram bytes %.0f)" % (
                "=> DRAM traffic / duration", traffic / t / 1e9, traffic))


def traffic(path, key, name_filter):
    """Average DRAM bytes (read + write) per captured launch whose kernel name
    contains name_filter -> profiles/traffic.json[key] (key = algo:scale:kind)."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    vals = []
    for r in rows[2:]:
        if name_filter not in r[ki]:
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(k)
            tot += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        vals.append(tot)
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "profiles", "traffic.json")
    table = json.load(open(dst)) if os.path.exists(dst) else {}
    table[key] = sum(vals) / len(vals)
    table[key + ":source"] = "%s (%d launches of %s)" % (
        os.path.basename(path), len(vals), name_filter)
    json.dump(table, open(dst, "w"), indent=1, sort_keys=True)
    print(key, table[key], len(vals))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
