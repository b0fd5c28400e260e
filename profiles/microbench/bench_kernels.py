#!/usr/bin/env python
"""Per-kernel timing at the SURVEY.md section 8(d) shapes (device-resident, CUDA events).

    python profiles/microbench/bench_kernels.py qcp dist fastns rdf contacts [--quick]

Development tool: prints one line per kernel with achieved GB/s (HBM-bound kernels)
or pairs/s.  bench.py is the contract benchmark; this is for fast iteration.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mdanalysis_b200 import _cabi                                    # noqa: E402
from mdanalysis_b200.lib import distances as gd                      # noqa: E402
from mdanalysis_b200.lib import qcprot                               # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(1)
quick = "--quick" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["qcp", "dist", "fastns", "rdf", "contacts"]


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


if "qcp" in which:
    na, F = 10000, (50000 if not quick else 4000)
    refpos = torch.randn((na, 3), generator=g, device=dev) * 20.0
    traj = (refpos[None] + torch.randn((F, na, 3), generator=g, device=dev)).contiguous()
    ref_np = refpos.cpu().numpy()
    w = np.random.default_rng(0).uniform(1, 32, na)
    w /= w.mean()
    for name, fn, bpa in (("rmsd", lambda: qcprot.qcp_trajectory(traj, ref_np), 12),
                          ("rmsd_weighted", lambda: qcprot.qcp_trajectory(traj, ref_np, weights=w), 12),
                          ("rmsd+rot", lambda: qcprot.qcp_trajectory(traj, ref_np, want_rotation=True), 12),
                          ("aligntraj", lambda: qcprot.qcp_trajectory(traj, ref_np, align=True, convention=1, out=traj), 24)):
        _cabi.profile_report()
        _cabi.profile_enable(True)
        med, best = timed(fn)
        prof = _cabi.profile_report()
        _cabi.profile_enable(False)
        kms = max(v[1] / v[0] for k, v in prof.items() if "traj" in k or "stream" in k)
        print(f"qcp {name:14s} F={F} n={na}: call {med:8.3f} ms (best {best:.3f}); main kernel {kms:.3f} ms = "
              f"{F / kms * 1e3:12.0f} frames/s {F * na * bpa / kms / 1e6:8.1f} GB/s algorithmic;  "
              + ", ".join(f"{k}={v[1] / v[0]:.3f}" for k, v in prof.items()), flush=True)
    del traj

if "dist" in which:
    lib = _cabi.load()
    n, F = 3341, (98 if not quick else 8)
    xyz = (torch.rand((F + 1, n, 3), generator=g, device=dev) * 50.0).contiguous()
    ref = xyz[:1].expand(F, n, 3).contiguous()
    conf = xyz[1:].contiguous()
    out = torch.empty((F, n, n), dtype=torch.float64, device=dev)
    sout = torch.empty((F, n * (n - 1) // 2), dtype=torch.float64, device=dev)
    st = _cabi.current_stream(xyz)
    boxes = {"none": (0, None), "ortho": (1, np.array([50, 50, 50], dtype=np.float32)),
             "triclinic": (2, gd._box_args(np.array([50, 50, 50, 70.53571, 109.48542, 70.518196], dtype=np.float32))[1])}
    for name, (mode, b) in boxes.items():
        med, best = timed(lambda: _cabi.check(lib.mdagpu_distance_array(
            _cabi.tensor_ptr(ref), n, _cabi.tensor_ptr(conf), n, mode, _cabi.ptr(b), _cabi.tensor_ptr(out), F, 1, st)))
        pairs = F * n * n
        print(f"distance_array {name:9s}: {med:8.3f} ms  {pairs / med / 1e6:8.1f} Gpairs/s  {pairs * 8 / med / 1e6:8.1f} GB/s", flush=True)
        med, best = timed(lambda: _cabi.check(lib.mdagpu_self_distance_array(
            _cabi.tensor_ptr(conf), n, mode, _cabi.ptr(b), _cabi.tensor_ptr(sout), F, 1, st)))
        pairs = F * (n * (n - 1) // 2)
        print(f"self_distance_array {name:9s}: {med:8.3f} ms  {pairs / med / 1e6:8.1f} Gpairs/s  {pairs * 8 / med / 1e6:8.1f} GB/s", flush=True)
    del out, sout, ref, conf, xyz

if "fastns" in which:
    from mdanalysis_b200.lib.mdamath import triclinic_vectors
    from mdanalysis_b200.lib.nsgrid import FastNS
    nn = 1_000_000 if not quick else 100_000
    for bname, angles in (("triclinic", (60, 60, 90)), ("ortho", (90, 90, 90))):
        vol_factor = 0.70710678 if bname == "triclinic" else 1.0
        L = (nn / 0.1 / vol_factor) ** (1.0 / 3.0)
        box = np.array([L, L, L, *angles], dtype=np.float32)
        tv = torch.as_tensor(triclinic_vectors(box).astype(np.float64), device=dev)
        pts = (torch.rand((nn, 3), generator=g, device=dev, dtype=torch.float64) @ tv).to(torch.float32).contiguous()
        torch.cuda.synchronize()
        for rep in range(3):
            t0 = time.perf_counter()
            grid = FastNS(8.0, pts, box)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            res = grid.self_search()
            npairs = len(res)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            del res
        print(f"fastns {bname:9s} n={nn}: build {1e3 * (t1 - t0):7.2f} ms  self_search {1e3 * (t2 - t1):7.2f} ms  "
              f"{npairs} pairs  {npairs / (t2 - t1) / 1e9:6.2f} Gpairs/s (wall)", flush=True)
        q = pts[: nn // 10].contiguous()
        t1 = time.perf_counter()
        res = grid.search(q)
        npairs = len(res)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"fastns {bname:9s} search nq={len(q)}: {1e3 * (t2 - t1):7.2f} ms  {npairs} pairs", flush=True)
        del res, grid, pts, q

if "rdf" in which:
    from mdanalysis_b200 import FrameBlock
    from mdanalysis_b200.analysis.rdf import InterRDF
    N, F = 33334, (200 if not quick else 20)
    for bname, angles, L in (("ortho", (90, 90, 90), 100.0), ("triclinic", (60, 60, 90), 112.25)):
        box = np.array([L, L, L, *angles], dtype=np.float32)
        from mdanalysis_b200.lib.mdamath import triclinic_vectors
        tv = torch.as_tensor(triclinic_vectors(box).astype(np.float32), device=dev)
        pos = (torch.rand((F, N, 3), generator=g, device=dev) @ tv).contiguous()
        blk = FrameBlock(pos, box)
        InterRDF(blk, None, nbins=150, range=(0.0, 15.0)).run()       # module load, pool growth
        _cabi.profile_report()
        _cabi.profile_enable(True)
        med, best = timed(lambda: InterRDF(blk, None, nbins=150, range=(0.0, 15.0)).run(), reps=3, warm=1)
        prof = _cabi.profile_report()
        _cabi.profile_enable(False)
        print(f"rdf {bname:9s} N={N} F={F}: {med:8.3f} ms  {F / med * 1e3:9.1f} frames/s   kernels(ms, 4 calls): "
              + ", ".join(f"{k}={v[1]:.2f}" for k, v in prof.items()), flush=True)
        del pos, blk

if "contacts" in which:
    from mdanalysis_b200.analysis.contacts import self_contacts
    nc, Fc = (100_000 if not quick else 20_000), 2
    side = (nc / 0.01) ** (1.0 / 3.0)
    cpos = (torch.rand((Fc, nc, 3), generator=g, device=dev) * side).contiguous()
    for name, bx in (("none", None), ("ortho", np.array([side] * 3 + [90.0] * 3, dtype=np.float32))):
        med, best = timed(lambda: self_contacts(cpos, 8.0, box=bx), reps=3, warm=1)
        print(f"contacts {name:6s} n={nc} F={Fc}: {med:8.3f} ms  {Fc / med * 1e3:8.1f} frames/s", flush=True)
