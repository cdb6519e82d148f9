#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an ncu report (no GPU needed):
joins `ncu --page source --csv` (SASS rows with executed-instruction counts) with the line table of the cubin
(`nvdisasm --print-line-info`), by instruction offset.

    scripts/ncu_lines.py <report.ncu-rep> <kernel regex> <object file with the kernel (.o / .so)> [min share %]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kre, obj = sys.argv[1:4]
    min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    name = next(csv.reader([lines[0]]))[1]
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    ia, ie, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    # the report may hold several launches of the kernel: keep the first block of addresses
    sass = []
    seen = set()
    for r in rows[1:]:
        if len(r) <= ie or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if a in seen:
            break
        seen.add(a)
        sass.append((a, int(r[ie] or 0), int(r[ism] or 0), r[1].strip()))
    base = sass[0][0]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    fn_re = re.compile(kre)
    want = {a - base for a, _, _, _ in sass}
    table = {}
    for cub in os.listdir(tmp):
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        cands, cur, cur_line = {}, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur = cands.setdefault(m.group(1), {}) if fn_re.search(m.group(1)) else None
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m and cur is not None:
                cur[int(m.group(1), 16)] = cur_line
        for fn, tb in cands.items():       # the template instance whose instruction offsets are exactly the profiled ones
            if set(tb) == want:
                table = tb
                break
        if table:
            break
    if not table:
        print("# no function with matching instruction offsets found (object file differs from the profiled build?)")
    agg = {}
    tot_i = tot_s = 0
    for a, n, smp, txt in sass:
        key = table.get(a - base, ("?", 0))
        e = agg.setdefault(key, [0, 0, 0])
        e[0] += n; e[1] += smp; e[2] += 1
        tot_i += n; tot_s += smp
    print(f"# {name}\n# warp instructions executed: {tot_i}, stall samples: {tot_s}, SASS instructions: {len(sass)}")
    src = {}
    print("| line | inst % | samples % | SASS | source |\n|---|---|---|---|---|")
    for (f, l), (n, smp, k) in sorted(agg.items(), key=lambda kv: kv[0]):
        if 100.0 * n / max(tot_i, 1) < min_share and 100.0 * smp / max(tot_s, 1) < min_share:
            continue
        if f not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
            for cand in (p, os.path.join("spfsplatv2_b200/csrc", f)):
                if os.path.exists(cand):
                    src[f] = open(cand).read().splitlines()
                    break
            else:
                src[f] = []
        text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
        print(f"| {f}:{l} | {100.0 * n / max(tot_i, 1):.1f} | {100.0 * smp / max(tot_s, 1):.1f} | {k} | `{text}` |")


if __name__ == "__main__":
    main()
