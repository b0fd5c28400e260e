#!/usr/bin/env python
"""bench.py -- headline benchmark of the CUDA geometry backend.

    python bench.py --gpus N --steps K --warmup W            (ours)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

Headline workload (BASELINE.json configs[1], "C2" in SURVEY.md section 8d):
InterRDF O-O on a synthetic 100k-atom water box -- 33 334 O sites, 1000 frames,
150 bins to 15 A, box 100 A.  One *step* = one pass of the hot path over the
1000 frames of a rank (`mdagpu_rdf_accumulate`: cell-list build + fused
pair/histogram kernel) followed by the reduction of the histogram over ranks.
Scaling is weak: every rank owns its own 1000 frames.

Prints ONE JSON line (see the task contract): metric/value/ms_per_step, the
`roofline` of the dominant kernel, `cpu_baseline` (oracle on the host, rank 0,
N = 1), `e2e` (same metric through the public API with HOST buffers, H2D/D2H
inside the timed region), `clocks`, `gpu_launches`, and `secondary`: the other
kernels of the path (distance arrays, QCP RMSD/AlignTraj, FastNS, contacts) with
their own rooflines.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017 + 1                      # 20261017 + config index
N_O, BOX_L, NBINS, RMAX = 33334, 100.0, 150, 15.0
# algorithmic FP32 flop per frame for C2 (stated in DESIGN.md):
# 33 334 queries x 1760 candidates (cell = r_c/2 traversal) x 11 flop
ALGO_FLOP_PER_FRAME = 33334 * 1760 * 11
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # nominal non-tensor FP32 at max clock


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clock/throttle samples taken DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons)}


def ncu_traffic(n_frames, n_atoms):
    """DRAM bytes per launch of the pair-histogram kernel from the committed ncu capture
    (profiles/rdf_kernel_traffic.json), scaled to this run's frames x atoms; None without it."""
    try:
        with open(os.path.join(ROOT, "profiles", "rdf_kernel_traffic.json")) as fh:
            t = json.load(fh)
        per_atom = (t["dram_bytes_read"] + t["dram_bytes_write"]) / (t["frames_per_launch"] * t["atoms"])
        return {"dram_bytes_per_launch": per_atom * n_frames * n_atoms, "unit": "B",
                "algorithmic_bytes_per_launch": 16.0 * n_frames * n_atoms, "source": t["source"]}
    except (OSError, KeyError, ValueError):
        return None


def make_frames(n_frames, n_atoms, rank):
    rng = np.random.default_rng(SEED + 1000 * rank)
    return (rng.random((n_frames, n_atoms, 3), dtype=np.float32) * np.float32(BOX_L))


# ---------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the same path
# ---------------------------------------------------------------------------------------
def cpu_rdf_sample(orc, pos0, box, nq):
    """capped_distance (FastNS, as the reference selects for n*m >= 1e8) + np.histogram for the
    first nq query atoms of one frame against all atoms.  Returns seconds."""
    t0 = time.perf_counter()
    pairs, dist = orc.capped_distance(pos0[:nq], pos0, RMAX, box=box, method="nsgrid")
    np.histogram(dist, bins=NBINS, range=(0.0, RMAX))
    return time.perf_counter() - t0, len(pairs)


def real_reference():
    """The unmodified reference built out of tree into baseline/_ref, or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "MDAnalysis")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import warnings
    warnings.filterwarnings("ignore")
    try:
        import MDAnalysis
        from MDAnalysis.analysis.rdf import InterRDF
    except Exception:
        return None
    return MDAnalysis, InterRDF


def reference_rdf_sample(ref, pos0, box, nq):
    """The reference's own InterRDF (AnalysisBase.run -> capped_distance -> FastNS -> np.histogram,
    analysis/rdf.py:253-316) on one frame: the first nq O sites against all of them."""
    mda, InterRDF = ref
    u = mda.Universe.empty(pos0.shape[0], trajectory=True)
    u.load_new(pos0[None].copy(), order="fac", dimensions=box[None])
    t0 = time.perf_counter()
    r = InterRDF(u.atoms[:nq], u.atoms, nbins=NBINS, range=(0.0, RMAX)).run()
    return time.perf_counter() - t0, r.results.count


def cpu_baseline_block(orc, pos0, box, n_atoms, budget_s=10.0):
    """`cpu_baseline` object: the reference itself when baseline/_ref is there (kind
    'reference'), with the oracle port's rate beside it; otherwise the port (kind 'port')."""
    ref = real_reference()
    nfr, t = 0, 0.0
    while t < budget_s and nfr < 16:
        dt_, _ = cpu_rdf_sample(orc, pos0, box, n_atoms)
        t += dt_
        nfr += 1
    port = nfr / t
    port_note = (f"oracle port (C restatement of lib/nsgrid.pyx + np.histogram): {nfr} full frames in {t:.1f} s")
    if ref is None:
        return {"value": port, "unit": "frames/s", "cores": 1, "kind": "port", "sample": port_note}
    nq = max(1, n_atoms // 8)
    reference_rdf_sample(ref, pos0, box, max(1, nq // 8))         # imports, first-call overheads
    dt_, _ = reference_rdf_sample(ref, pos0, box, nq)
    return {"value": (nq / n_atoms) / dt_, "unit": "frames/s", "cores": 1, "kind": "reference",
            "sample": f"MDAnalysis 2.8.0-dev0 (baseline/_ref) InterRDF.run() on one frame, {nq} of {n_atoms} O sites "
                      f"against all {n_atoms} (1/8 of a frame's pairs) in {dt_:.1f} s; InterRDF / FastNS are serial in "
                      "the reference (no multi-core variant, analysis/rdf.py: not parallelizable)",
            "port_value": port, "port_sample": port_note}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    box = np.array([BOX_L] * 3 + [90.0] * 3, dtype=np.float32)
    pos0 = make_frames(1, args.atoms, 0)[0]
    ref = real_reference()
    if ref is not None:
        nq = max(1, args.atoms // 8)      # 1/8 of a frame's pairs per step: about 3 s
        run = lambda: reference_rdf_sample(ref, pos0, box, nq)
        kind = "reference"
        sample = (f"per step: MDAnalysis 2.8.0-dev0 (baseline/_ref, built from /root/reference) InterRDF.run() on one "
                  f"frame of the workload, {nq} of {args.atoms} O sites against all {args.atoms} (1/8 of a frame's "
                  "pairs): capped_distance -> FastNS -> np.histogram; serial, as the reference has no multi-core "
                  "variant of FastNS / InterRDF")
    else:
        nq = args.atoms                   # one full frame per step (about 1.4 s with the C port)
        run = lambda: cpu_rdf_sample(orc, pos0, box, nq)
        kind = "port"
        sample = (f"per step: one frame of the workload ({args.atoms} x {args.atoms} O sites): FastNS build + search "
                  "+ np.histogram; oracle port of lib/nsgrid.pyx (serial, as the reference)")
    frac = nq / args.atoms
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = (time.perf_counter() - t0) / args.steps
    fps = frac / dt
    line = {
        "impl": "reference", "metric": "InterRDF frames/s (O-O, 33334 sites, 150 bins to 15 A)",
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 InterRDF O-O synthetic water box", "atoms": args.atoms, "box": BOX_L,
                   "nbins": NBINS, "rmax": RMAX},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------
# ours
# ---------------------------------------------------------------------------------------
def time_region(torch, fn, steps, dist_ready):
    import torch.distributed as dist
    if dist_ready:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if dist_ready:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist_ready:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def secondary_benchmarks(torch, hbm_peak, peak_kind, quick):
    """Other kernels of the path at their SURVEY.md section 8d shapes (device-resident, 1 GPU)."""
    from mdanalysis_b200 import _cabi
    from mdanalysis_b200.lib import distances as gd
    from mdanalysis_b200.lib import qcprot
    from mdanalysis_b200.lib.nsgrid import FastNS
    from mdanalysis_b200.analysis.contacts import self_contacts
    out = {}
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # C1: distance_array + self_distance_array, adk-sized (3341 atoms), 98 frames in one launch
    n, F = 3341, 98 if not quick else 8
    xyz = (torch.rand((F + 1, n, 3), generator=g, device=dev) * 50.0).contiguous()
    ref = xyz[:1].expand(F, n, 3).contiguous()
    conf = xyz[1:].contiguous()
    outbuf = torch.empty((F, n, n), dtype=torch.float64, device=dev)
    sout = torch.empty((F, n * (n - 1) // 2), dtype=torch.float64, device=dev)
    lib = _cabi.load()
    st = _cabi.current_stream(xyz)
    boxes = {"none": (0, None), "ortho": (1, np.array([50, 50, 50], dtype=np.float32)),
             "triclinic": (2, gd._box_args(np.array([50, 50, 50, 70.53571, 109.48542, 70.518196], dtype=np.float32))[1])}
    for name, (mode, b) in boxes.items():
        ms = timed(lambda: _cabi.check(lib.mdagpu_distance_array(
            _cabi.tensor_ptr(ref), n, _cabi.tensor_ptr(conf), n, mode, _cabi.ptr(b), _cabi.tensor_ptr(outbuf), F, 1, st)), 3)
        pairs = F * n * n
        gbs = pairs * 8 / ms / 1e6
        out[f"distance_array_{name}"] = {"pairs_per_s": pairs / ms * 1e3, "ms": ms, "shape": [F, n, n],
                                         "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                                      "frac": gbs / hbm_peak, "peak_kind": peak_kind}}
        ms = timed(lambda: _cabi.check(lib.mdagpu_self_distance_array(
            _cabi.tensor_ptr(conf), n, mode, _cabi.ptr(b), _cabi.tensor_ptr(sout), F, 1, st)), 3)
        pairs = F * (n * (n - 1) // 2)
        gbs = pairs * 8 / ms / 1e6
        out[f"self_distance_array_{name}"] = {"pairs_per_s": pairs / ms * 1e3, "ms": ms,
                                              "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                                           "frac": gbs / hbm_peak, "peak_kind": peak_kind}}
    # C1 as a user calls it: one frame pair through the NumPy front end (host arrays in, host array out),
    # beside the reference's own C kernels (oracle/_ref: calc_distances.h compiled as the reference builds
    # it) on one host core and, for the OpenMP build, on all of them
    try:
        from oracle import oracle as orc
        orc.build()
        x0, x1 = xyz[0].cpu().numpy(), xyz[1].cpu().numpy()
        box6 = {"none": None, "ortho": np.array([50, 50, 50, 90, 90, 90], dtype=np.float32),
                "triclinic": np.array([50, 50, 50, 70.53571, 109.48542, 70.518196], dtype=np.float32)}
        res = np.empty((n, n), dtype=np.float64)
        have_ref = orc.have_reference()
        for name, b6 in box6.items():
            gd.distance_array(x0, x1, box=b6, result=res)
            t0 = time.perf_counter()
            for _ in range(3):
                gd.distance_array(x0, x1, box=b6, result=res)
            t_gpu = (time.perf_counter() - t0) / 3
            entry = {"pairs": n * n, "gpu_call_ms": t_gpu * 1e3,
                     "note": "distance_array(ndarray, ndarray, box, result=ndarray): H2D + kernel + D2H of the 89 MB result"}
            if have_ref:
                t0 = time.perf_counter()
                want = orc.distance_array(x0, x1, b6, impl="reference")
                entry["cpu_serial_ms"] = (time.perf_counter() - t0) * 1e3
                if orc.have_reference(openmp=True):
                    t0 = time.perf_counter()
                    orc.distance_array(x0, x1, b6, impl="reference", openmp=True)
                    entry["cpu_openmp_ms"] = (time.perf_counter() - t0) * 1e3
                    entry["cpu_cores"] = len(os.sched_getaffinity(0))
                entry["identical"] = bool(np.array_equal(res, want))
            out[f"c1_single_call_{name}"] = entry
    except Exception as e:                                            # the baseline is optional here
        out["c1_single_call_error"] = repr(e)
    del outbuf, sout, ref, conf, xyz
    torch.cuda.empty_cache()

    # C4: RMSD + AlignTraj (QCP), 10k atoms; frames limited by what one call needs in HBM
    na, Fq = 10000, 50000 if not quick else 2000
    refpos = (torch.randn((na, 3), generator=g, device=dev) * 20.0)
    traj = (refpos[None] + torch.randn((Fq, na, 3), generator=g, device=dev)).contiguous()
    ref_np = refpos.cpu().numpy()
    ms = timed(lambda: qcprot.qcp_trajectory(traj, ref_np), 3)
    gbs = Fq * na * 12 / ms / 1e6
    out["rmsd_qcp"] = {"frames_per_s": Fq / ms * 1e3, "ms": ms, "shape": [Fq, na, 3],
                       "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": gbs / hbm_peak, "peak_kind": peak_kind}}
    ms = timed(lambda: qcprot.qcp_trajectory(traj, ref_np, align=True, convention=1, out=traj), 3)
    gbs = Fq * na * 24 / ms / 1e6
    out["aligntraj_qcp"] = {"frames_per_s": Fq / ms * 1e3, "ms": ms,
                            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                         "frac": gbs / hbm_peak, "peak_kind": peak_kind}}
    # section 8f: all-pairs RMSD matrix (diffusionmap.DistanceMatrix) of the first frames of the same trajectory
    Fp = 2048 if not quick else 256
    fp64_peak = 61.4 * 148 * 1.965e9 / 1e12      # T lane-ops/s, DFMA rate of profiles/microbench/pipes_b200.txt
    for label, sup, ops in (("pairwise_rmsd", False, 2 * 3), ("pairwise_rmsd_superposed", True, 9)):
        ms = timed(lambda: qcprot.pairwise_rmsd(traj[:Fp], superposition=sup), 3)
        npair = Fp * (Fp + 1) // 2
        tops = npair * na * ops / ms / 1e9
        out[label] = {"frame_pairs_per_s": npair / ms * 1e3, "ms": ms, "shape": [Fp, na, 3],
                      "roofline": {"bound": "fp64", "achieved": tops, "peak": fp64_peak, "unit": "T FP64 lane-ops/s",
                                   "frac": tops / fp64_peak, "peak_kind": "measured DFMA issue rate (61.4 lanes/clk/SM at 1965 MHz)",
                                   "note": f"{ops} FP64 instructions per frame pair and atom"}}
    del traj
    torch.cuda.empty_cache()

    # C3: FastNS self_search, 8 A, 1M atoms, triclinic [L,L,L,60,60,90]
    nn = 1_000_000 if not quick else 100_000
    L = (nn / 0.1 / 0.70710678) ** (1.0 / 3.0)
    box = np.array([L, L, L, 60, 60, 90], dtype=np.float32)
    from mdanalysis_b200.lib.mdamath import triclinic_vectors
    tv = torch.as_tensor(triclinic_vectors(box).astype(np.float64), device=dev)
    pts = (torch.rand((nn, 3), generator=g, device=dev, dtype=torch.float64) @ tv).to(torch.float32).contiguous()
    res = FastNS(8.0, pts, box).self_search()          # warm-up: module load, result pool growth
    del res
    torch.cuda.synchronize()
    best_build = best_search = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        grid = FastNS(8.0, pts, box)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        res = grid.self_search()
        npairs = len(res)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        best_build, best_search = min(best_build, t1 - t0), min(best_search, t2 - t1)
        del res
    out["fastns_self_search"] = {"atoms": nn, "pairs": npairs, "build_ms": best_build * 1e3,
                                 "search_ms": best_search * 1e3, "pairs_per_s": npairs / best_search,
                                 "roofline": {"bound": "hbm", "achieved": npairs * 24 / best_search / 1e9, "peak": hbm_peak,
                                              "unit": "GB/s", "frac": npairs * 24 / best_search / 1e9 / hbm_peak,
                                              "peak_kind": peak_kind,
                                              "note": "24 B written per pair; the search itself is FP32/FP64 pipe bound"},
                                 "note": "wall clock of the call (result buffers from the stream-ordered pool), best of 3"}
    del grid, pts
    torch.cuda.empty_cache()

    # C5: fused contacts, 100k atoms, 8 A
    nc, Fc = 100_000 if not quick else 20_000, 2
    side = (nc / 0.01) ** (1.0 / 3.0)
    cpos = (torch.rand((Fc, nc, 3), generator=g, device=dev) * side).contiguous()
    for name, bx in (("none", None), ("ortho", np.array([side] * 3 + [90.0] * 3, dtype=np.float32))):
        ms = timed(lambda: self_contacts(cpos, 8.0, box=bx), 2)
        pairs = Fc * (nc * (nc - 1) // 2)
        out[f"contacts_self_{name}"] = {"atoms": nc, "frames": Fc, "ms": ms, "pair_decisions_per_s": pairs / ms * 1e3}
    # C5 (i): the materialised self_distance_array of one 100k-atom frame (40 GB of float64, written once)
    npair = nc * (nc - 1) // 2
    big = torch.empty((npair,), dtype=torch.float64, device=dev)
    one = cpos[:1].contiguous()
    for name, (mode, b) in (("none", (0, None)), ("ortho", (1, np.array([side] * 3, dtype=np.float32)))):
        ms = timed(lambda: _cabi.check(lib.mdagpu_self_distance_array(
            _cabi.tensor_ptr(one), nc, mode, _cabi.ptr(b), _cabi.tensor_ptr(big), 1, 1, _cabi.current_stream(one))), 2)
        gbs = npair * 8 / ms / 1e6
        out[f"c5_self_distance_array_{name}"] = {"atoms": nc, "pairs": npair, "ms": ms, "pairs_per_s": npair / ms * 1e3,
                                                 "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                                              "frac": gbs / hbm_peak, "peak_kind": peak_kind}}
    del big
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the cuda backend has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist_ready = world > 1
    if dist_ready:
        # stdout carries ONE JSON line: NCCL's banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mdanalysis_b200 import FrameBlock, _cabi
    from mdanalysis_b200.analysis.rdf import InterRDF

    hbm_peak, peak_kind = load_peaks()
    F, N = args.frames, args.atoms
    box = np.array([BOX_L] * 3 + [90.0] * 3, dtype=np.float32)
    host = torch.from_numpy(make_frames(F, N, rank)).pin_memory()
    dev = host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    blk_dev = FrameBlock(dev, box)
    blk_host = FrameBlock(host.numpy(), box)

    result = {}

    def step_device():
        result["rdf"] = InterRDF(blk_dev, None, nbins=NBINS, range=(0.0, RMAX), distributed=dist_ready).run()

    def step_host():
        result["rdf_host"] = InterRDF(blk_host, None, nbins=NBINS, range=(0.0, RMAX), distributed=dist_ready).run()

    for _ in range(max(args.warmup, 3)):
        step_device()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    _cabi.profile_report()
    _cabi.profile_enable(True)
    launches0 = _cabi.launch_count()
    ms = time_region(torch, step_device, args.steps, dist_ready)
    launches = _cabi.launch_count() - launches0
    prof = _cabi.profile_report()
    _cabi.profile_enable(False)
    clock_info = clocks.stop() if rank == 0 else None
    frames_total = world * F * args.steps
    value = frames_total / (ms / 1e3)

    # dominant kernel: fused pair search + histogram
    kname = "rdf_pair_histogram"
    k_launches, k_ms = prof.get(kname, (0, 0.0))
    achieved = (ALGO_FLOP_PER_FRAME * (N / N_O) ** 2 * F * k_launches / (k_ms / 1e3) / 1e12) if k_ms > 0 else 0.0
    roofline = {"bound": "fp32", "kernel": kname, "achieved": achieved, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved / FP32_PEAK_TFLOPS, "traffic": ncu_traffic(F, N),
                "peak_kind": "nominal non-tensor FP32 at 1965 MHz (MEASURED_PEAKS.json has no FP32 figure)",
                "kernel_ms_per_launch": (k_ms / k_launches) if k_launches else None,
                "kernel_share_of_step": (k_ms / ms) if ms > 0 else None,
                "kernels_ms": {k: round(v[1], 4) for k, v in prof.items()}}

    # end to end through the public API with HOST buffers (H2D + D2H inside the region)
    step_host()
    torch.cuda.synchronize()
    if dist_ready:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist_ready:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": frames_total / dt, "unit": "frames/s", "h2d_bytes_per_step": F * N * 12 + (NBINS + 1) * 8,
           "d2h_bytes_per_step": NBINS * 8, "ms_per_step": dt / args.steps * 1e3,
           "api": "InterRDF(FrameBlock(host ndarray), ...).run() -> mdagpu_rdf_accumulate(space=HOST)"}

    same = bool(np.array_equal(result["rdf"].results.count, result["rdf_host"].results.count))

    cpu_baseline = None
    secondary = None
    if rank == 0 and world == 1:
        from oracle import oracle as orc
        orc.build()
        cpu_baseline = cpu_baseline_block(orc, host[0].numpy(), box, N)
        if not args.no_secondary:
            del blk_dev, dev
            torch.cuda.empty_cache()
            secondary = secondary_benchmarks(torch, hbm_peak, peak_kind, args.quick)

    if rank == 0:
        line = {
            "metric": "InterRDF frames/s (O-O, 33334 sites, 150 bins to 15 A)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 filter + f64 exact",
            "data": "synthetic",
            "config": {"workload": "C2 InterRDF O-O synthetic water box", "atoms": N, "frames_per_gpu": F, "box": BOX_L,
                       "nbins": NBINS, "rmax": RMAX, "l2": "inputs larger than L2 (400 MB per pass)",
                       "parallelism": f"frames sharded over {world} GPU(s); NCCL all_reduce of the histogram"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clock_info, "parity_device_vs_host_path": same, "secondary": secondary,
        }
        emit(line)
    if dist_ready:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """stdout carries ONE JSON line.  Libraries below (NCCL's version banner, anything a C
    extension prints) write to file descriptor 1 directly, so fd 1 is pointed at stderr for the
    rest of the process and the JSON line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="frames per GPU (default: the C2 workload)")
    ap.add_argument("--atoms", type=int, default=N_O, help="O sites per frame (default: the C2 workload)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary kernels")
    ap.add_argument("--quick", action="store_true", help="smaller secondary shapes (smoke runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
