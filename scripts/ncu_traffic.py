#!/usr/bin/env python
"""DRAM bytes per launch of every spf:: kernel in an ncu --set full capture -> profiles/rN_traffic.json (what bench.py
reports as roofline.traffic).   scripts/ncu_traffic.py <report.ncu-rep> <workload> <out.json>"""
import csv
import io
import json
import subprocess
import sys

NAMES = {"project_forward": "project_forward", "emit_kernel": "emit", "tile_sort_pack": "tile_sort_pack",
         "blend_forward": "blend_forward", "blend_backward_log": "blend_backward", "project_backward": "project_backward",
         "scan_kernel": "scan", "pose_reduce": "pose_reduce"}


def main():
    rep, workload, out = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                     "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active")}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        name = next((v for k, v in NAMES.items() if k in r[col["Kernel Name"]]), None)
        if name is None:
            continue
        b = sum(float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        dur = dur / 1e3 if units[col["gpu__time_duration.sum"]] in ("ns", "nsecond") else dur
        e = acc.setdefault(name, {"n": 0, "bytes": 0.0, "us": 0.0, "inst": 0.0, "issue": 0.0})
        e["n"] += 1; e["bytes"] += b; e["us"] += dur
        e["inst"] += float(r[col["smsp__inst_executed.sum"]].replace(",", ""))
        e["issue"] += float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]].replace(",", ""))
    res = {"workload": workload, "source": f"ncu --set full --clock-control none, {rep} (dram__bytes_read.sum + dram__bytes_write.sum per launch)",
           "kernels": {k: {"dram_bytes_per_launch": round(v["bytes"] / v["n"]), "duration_us_under_ncu": round(v["us"] / v["n"], 3),
                           "warp_instructions_per_launch": round(v["inst"] / v["n"]), "issue_slots_busy_pct": round(v["issue"] / v["n"], 1),
                           "launches_captured": v["n"]} for k, v in acc.items()}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
