#!/usr/bin/env python
"""Turns `ncu -i X.ncu-rep --page raw --csv` output + a launch-list CSV into the markdown summary committed under
profiles/.  Usage: scripts/ncu_summary.py <launches.csv> <raw1.csv> [<raw2.csv> ...] > profiles/rN_summary.md"""
import collections
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "smem static"),
    ("launch__shared_mem_per_block_dynamic", "smem dynamic"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"## Launch list `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot:.3f} |")
    ours = sum(a[1] for k, a in agg.items() if "spf::" in k)
    print(f"\nour kernels (spf::*) = {ours / tot:.3f} of all device time in the capture; the rest is torch glue (loss, fills, camera inverse).\n")


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f"## `ncu --set full` capture `{path}`\n")
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")].split("(")[0]
        print(f"### {name}\n\n| metric | value |\n|---|---|")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} (`{key}`) | {row[i]} {units[i]} |")
        print()


if __name__ == "__main__":
    launches(sys.argv[1])
    for p in sys.argv[2:]:
        full(p)
