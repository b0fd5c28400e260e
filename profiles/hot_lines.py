#!/usr/bin/env python
"""Hottest source lines of an `ncu --page source --csv` dump (needs -lineinfo builds):

    ncu -i prof.ncu-rep --page source --csv > prof.source.csv
    python profiles/hot_lines.py prof.source.csv [N] [--by-inst]

Prints, per source line (SASS, or CUDA-C with `--print-source cuda`), the warp-stall samples and
executed instructions, sorted by samples (or by instructions), with the dominant stall reasons."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    args = [a for a in sys.argv[2:] if not a.startswith("--")]
    by_inst = "--by-inst" in sys.argv
    top = int(args[0]) if args else 30
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next((i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c or "Samples" in c for c in r)), None)
    if hdr_i is None:
        print("no source page in", path)
        return
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    s_col = next((h for h in hdr if h.startswith("Warp Stall Sampling (All")), None) or next(h for h in hdr if "Sampl" in h)
    i_col = next((h for h in hdr if h.startswith("Instructions Executed")), None)
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    src_col = "Source"
    agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float), 0])
    total = 0.0
    for r in rows[hdr_i + 1:]:
        if len(r) != len(hdr):
            continue
        try:
            s = float(r[col[s_col]] or 0)
        except ValueError:
            continue
        key = r[col[src_col]].strip()
        a = agg[key]
        a[0] += s
        if i_col:
            try:
                a[1] += float(r[col[i_col]] or 0)
            except ValueError:
                pass
        for h in stall_cols:
            try:
                a[2][h] += float(r[col[h]] or 0)
            except ValueError:
                pass
        a[3] += 1
        total += s
    total_i = sum(a[1] for a in agg.values())
    print(f"-- hottest lines ({s_col}; total {total:.0f} samples, {total_i:.0f} warp instructions; sorted by {'instructions' if by_inst else 'samples'})")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_inst else 0])[:top]:
        why = ", ".join(f"{h[6:]}={v / max(a[0], 1):.0%}" for h, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3] if v > 0)
        print(f"{a[0] / max(total, 1):6.1%} {a[1]:14.0f} inst  {key[:110]}   [{why}]")


main()
