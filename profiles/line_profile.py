#!/usr/bin/env python
"""Executed instructions and stall samples of one kernel PER SOURCE LINE.

    ncu -i prof.ncu-rep --page source --csv > prof.source.csv            # SASS view, one row per instruction
    python profiles/line_profile.py prof.source.csv mdanalysis_b200/csrc/rdf.o fine_plan_kernelILb1ELi224ELi2 [N]

The source page of an ncu report lists the kernel's SASS in address order; `nvdisasm -gi` of the same
object (built with -lineinfo) gives the source line of every instruction in the same order.  Joining the
two by position attributes the per-instruction counters to file:line (inlined code: the innermost line)."""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(obj, kernel):
    with tempfile.TemporaryDirectory() as tmp:
        import os
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
        import glob
        cubin = glob.glob(tmp + "/*.cubin")[0]
        text = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
    out, cur, active = [], ("?", 0), False
    for ln in text.splitlines():
        if ln.startswith("//--------------------- .text."):
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln) and ";" in ln:
            out.append((cur, ln.split("*/", 1)[1].strip()))
    return out


def main():
    path, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r))
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    s_col = next(h for h in hdr if h.startswith("Warp Stall Sampling (All"))
    i_col = next(h for h in hdr if h.startswith("Instructions Executed"))
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    sass = sass_lines(obj, kernel)
    if len(sass) != len(body):
        print(f"warning: {len(body)} profiled instructions vs {len(sass)} disassembled -- joined by position up to the shorter")
    agg = defaultdict(lambda: [0.0, 0.0])
    for (where, _), r in zip(sass, body):
        try:
            agg[where][0] += float(r[col[i_col]] or 0)
            agg[where][1] += float(r[col[s_col]] or 0)
        except ValueError:
            pass
    ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"-- {kernel}: {ti:.0f} warp instructions, {ts:.0f} stall samples; lines by executed instructions")
    src_cache = {}
    for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src_cache:
            try:
                src_cache[f] = open("mdanalysis_b200/csrc/" + f).read().splitlines()
            except OSError:
                src_cache[f] = []
        code = src_cache[f][n - 1].strip()[:90] if 0 < n <= len(src_cache[f]) else ""
        print(f"{a[0] / max(ti, 1):6.1%} inst {a[1] / max(ts, 1):6.1%} stall  {f}:{n:<5d} {code}")


main()
