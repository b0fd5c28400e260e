#!/usr/bin/env python
"""Rates of the staged pageable transfers (csrc/hostcopy.cu) against pinned memory and against the
driver's own pageable path (MDAGPU_NO_STAGED_COPY=1), per thread count.

    python profiles/microbench/hostcopy_bench.py            (spawns itself per setting)
"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child():
    import torch
    from mdanalysis_b200.lib import distances as gd
    from mdanalysis_b200.lib import qcprot
    rng = np.random.default_rng(0)
    a = (rng.random((3341, 3)) * 50).astype(np.float32)
    b = (rng.random((3341, 3)) * 50).astype(np.float32)
    res = np.empty((3341, 3341))
    pin = torch.empty((3341, 3341), dtype=torch.float64).pin_memory().numpy()
    traj = rng.random((3000, 10000, 3), dtype=np.float32)              # 360 MB
    tpin = torch.as_tensor(traj).pin_memory().numpy()
    ref = traj[0]
    out = []
    for name, fn, mb in (("d2h 89 MB pageable", lambda: gd.distance_array(a, b, result=res), 89.3),
                         ("d2h 89 MB pinned", lambda: gd.distance_array(a, b, result=pin), 89.3),
                         ("h2d 360 MB pageable", lambda: qcprot.qcp_trajectory(traj, ref), 360.0),
                         ("h2d 360 MB pinned", lambda: qcprot.qcp_trajectory(tpin, ref), 360.0)):
        fn()
        fn()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        dt = (time.perf_counter() - t0) / 5
        out.append(f"{name}: {dt * 1e3:7.2f} ms {mb / dt / 1e3:6.1f} GB/s")
    print(" | ".join(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for env in ({"MDAGPU_NO_STAGED_COPY": "1"}, {"MDAGPU_COPY_THREADS": "1"}, {"MDAGPU_COPY_THREADS": "2"},
                    {"MDAGPU_COPY_THREADS": "4"}, {"MDAGPU_COPY_THREADS": "8"}, {"MDAGPU_COPY_THREADS": "12"},
                    {"MDAGPU_COPY_THREADS": "16"}):
            e = dict(os.environ)
            e.update(env)
            print(env, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e)
