#!/usr/bin/env python
"""Condensed view of a bench.py JSON line: python profiles/show_bench.py FILE"""
import json
import sys


def walk(d, pre=""):
    for k, v in d.items():
        if isinstance(v, dict):
            if "roofline" in v or "ms" in v or "value" in v or "search_ms" in v:
                keys = ("value", "unit", "ms", "ms_per_step", "frames_per_s", "pairs_per_s", "search_ms", "build_ms", "candidates_per_s",
                        "ms_per_frame", "ms_per_call", "frames_this_rank", "serial_ms", "openmp_ms", "pageable_result_ms", "pinned_result_ms")
                flat = {a: (round(b, 4) if isinstance(b, float) and abs(b) < 1e6 else (f"{b:.4g}" if isinstance(b, float) else b))
                        for a, b in v.items() if a in keys}
                r = v.get("roofline")
                if isinstance(r, dict):
                    flat["frac"] = round(r["frac"], 3)
                    flat["bound"] = r["bound"]
                print(f"{pre}{k}: {flat}")
            walk(v, pre + k + ".")
        elif k in ("error",):
            print(f"{pre}{k}: {v}")


d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "ms_per_step", "n_gpus", "gpu_launches", "bench_wall_s") if k in d})
print("clocks", d.get("clocks"))
walk(d)
