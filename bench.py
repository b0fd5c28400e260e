#!/usr/bin/env python
"""bench.py -- benchmark of the CUDA geometry backend, every config of BASELINE.json at every N.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference arm)

Headline (BASELINE.json configs[1], "C2" in SURVEY.md section 8d): InterRDF O-O on a synthetic
100k-atom water box -- 33 334 O sites, 1000 frames, 150 bins to 15 A, box 100 A.  One *step* = one
pass of the hot path over the 1000 frames of a rank (`mdagpu_rdf_accumulate`: cell-list build,
window planner, fused pair/histogram kernel) followed by ONE all_reduce of [histogram, volume,
frames] over the ranks.  The headline scales weakly (every rank owns its own 1000 frames); the
STRONG scaling of the stated 1000-frame total (np.array_split over the ranks, analysis/base.py:588)
is measured in the same run and reported under `strong`.

The JSON line also carries, under `secondary`, the other four configs at the same N, each
frame-sharded like the reference's `_setup_computation_groups` with its real collective on NCCL
(C1 none, C4 all_gather of rmsd[F], C5 all_gather of the contact counts, C3 independent replicas),
each with `roofline`, `e2e` (HOST buffers through the public API, copies inside the timed region)
and -- rank 0, N = 1 -- `cpu_baseline` (the unmodified reference from baseline/_ref on the box's
host cores, bounded samples).
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

SEED0 = 20261017                          # seed = 20261017 + config index
SEED = SEED0 + 1                          # C2
N_O, BOX_L, NBINS, RMAX = 33334, 100.0, 150, 15.0
# algorithmic FP32 flop per frame for C2 (stated in DESIGN.md):
# 33 334 queries x 1760 candidates (cell = r_c/2 traversal) x 11 flop
ALGO_FLOP_PER_FRAME = 33334 * 1760 * 11
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12       # nominal non-tensor FP32 at max clock
# measured FFMA issue rate, profiles/microbench/pipes_b200.txt: 117.76 lane-ops/clk/SM at 1.965 GHz
FP32_MEASURED_TFLOPS = 117.76 * 148 * 2 * 1.965e9 / 1e12
FP64_MEASURED_TOPS = 60.66 * 148 * 1.965e9 / 1e12          # DFMA lane-ops/s (same file)
CANDIDATE_FLOP = 8                         # FP32 filter of a cutoff search: 3 sub + mul + 2 fma (no wrap: image known per cell pair)


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
    """DRAM bytes per step of the pair-histogram kernel from the committed ncu capture
    (profiles/rdf_kernel_traffic.json), scaled to this run's frames x atoms; None without it."""
    try:
        with open(os.path.join(ROOT, "profiles", "rdf_kernel_traffic.json")) as fh:
            t = json.load(fh)
        per_atom = (t["dram_bytes_read"] + t["dram_bytes_write"]) / (t["frames_per_launch"] * t["atoms"])
        return {"dram_bytes_per_launch": per_atom * n_frames * n_atoms, "unit": "B",
                "algorithmic_bytes_per_launch": 16.0 * n_frames * n_atoms, "source": t["source"],
                "note": t.get("note")}
    except (OSError, KeyError, ValueError):
        return None


def make_frames(n_frames, n_atoms, rank):
    rng = np.random.default_rng(SEED + 1000 * rank)
    return (rng.random((n_frames, n_atoms, 3), dtype=np.float32) * np.float32(BOX_L))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------
# the reference on the host cores (cpu_baseline legs and the reference arm)
# ---------------------------------------------------------------------------------------
_REF = None


def real_reference():
    """The unmodified reference built out of tree into baseline/_ref, or None."""
    global _REF
    if _REF is not None:
        return _REF or None
    ref = os.path.join(ROOT, "baseline", "_ref")
    _REF = False
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
    _REF = (MDAnalysis, InterRDF)
    return _REF


def cpu_rdf_sample(orc, pos0, box, nq):
    """capped_distance (FastNS, as the reference selects for n*m >= 1e8) + np.histogram for the
    first nq query atoms of one frame against all atoms.  Returns seconds."""
    t0 = time.perf_counter()
    pairs, dist = orc.capped_distance(pos0[:nq], pos0, RMAX, box=box, method="nsgrid")
    np.histogram(dist, bins=NBINS, range=(0.0, RMAX))
    return time.perf_counter() - t0, len(pairs)


def reference_rdf_sample(ref, pos0, box, nq):
    """The reference's own InterRDF (AnalysisBase.run -> capped_distance -> FastNS -> np.histogram,
    analysis/rdf.py:253-316) on one frame: the first nq O sites against all of them."""
    mda, InterRDF = ref
    u = mda.Universe.empty(pos0.shape[0], trajectory=True)
    u.load_new(pos0[None].copy(), order="fac", dimensions=box[None])
    t0 = time.perf_counter()
    r = InterRDF(u.atoms[:nq], u.atoms, nbins=NBINS, range=(0.0, RMAX)).run()
    return time.perf_counter() - t0, r.results.count


def cpu_baseline_c2(orc, pos0, box, n_atoms, budget_s=8.0):
    """`cpu_baseline` of the headline: the reference itself when baseline/_ref is there (kind
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
# ours: shared plumbing
# ---------------------------------------------------------------------------------------
class Ctx:
    """What every config needs: torch, the process group, this rank, peaks, flags."""

    def __init__(self, torch, dist, rank, world, local_rank, args):
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dist_ready = world > 1
        self.args = args
        self.quick = args.quick
        self.hbm_peak, self.peak_kind = load_peaks()
        self.cpu = rank == 0 and world == 1 and not args.no_cpu
        self.dev = torch.device("cuda", local_rank)
        self._flush = None

    def flush_l2(self):
        """Write 512 MB (4x the 126 MB L2): the next kernel finds none of its inputs cached."""
        if self._flush is None:
            self._flush = self.torch.empty(512 << 20, dtype=self.torch.uint8, device=self.dev)
        self._flush.add_(1)

    def barrier(self):
        if self.dist_ready:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if not self.dist_ready:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if not self.dist_ready:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, reps, warm=1, flush=False):
        """ms per call of fn (device time, CUDA events on torch's current stream = the stream the
        library launches on), MAX over ranks.  flush=True: inputs smaller than L2 -- every call is
        timed on its own with an L2 flush before it."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not flush:
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
        else:
            ms = 0.0
            for _ in range(reps):
                self.flush_l2()
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1)
            ms /= reps
        self.barrier()
        return self.max_over_ranks(ms)

    def wall(self, fn, reps, warm=1):
        """seconds per call by the host clock (end-to-end legs with host buffers), MAX over ranks."""
        for _ in range(warm):
            fn()
        self.torch.cuda.synchronize()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        self.torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        self.barrier()
        return self.max_over_ranks(dt)

    def hbm(self, gbs):
        return {"bound": "hbm", "achieved": gbs, "peak": self.hbm_peak, "unit": "GB/s",
                "frac": gbs / self.hbm_peak, "peak_kind": self.peak_kind + " copy rate (MEASURED_PEAKS.json)"}

    @staticmethod
    def fp32(tflops, note=None):
        r = {"bound": "fp32", "achieved": tflops, "peak": FP32_MEASURED_TFLOPS, "unit": "TFLOP/s",
             "frac": tflops / FP32_MEASURED_TFLOPS, "frac_of_nominal": tflops / FP32_NOMINAL_TFLOPS,
             "peak_kind": "measured FFMA issue rate, 117.76 lane-ops/clk/SM at 1965 MHz (profiles/microbench/"
                          "pipes_b200.txt); nominal 74.45"}
        if note:
            r["note"] = note
        return r


def shard(n_total, world, rank):
    """Contiguous block of `rank` under np.array_split (analysis/base.py:588)."""
    idx = np.array_split(np.arange(n_total), world)[rank]
    return (int(idx[0]), int(idx[-1]) + 1) if len(idx) else (0, 0)


def stencil_candidates(torch, frac, ncells, periodic, half=True):
    """Candidate pairs of a 27-cell-stencil traversal on the grid the library reports: for every
    atom, the atoms in its own and the 26 neighbouring cells (periodic wrap where the grid wraps);
    half=True counts each unordered pair once without the self pair (self_search / contacts)."""
    nx, ny, nz = (int(v) for v in ncells)
    nc = torch.tensor([nx, ny, nz], device=frac.device)
    c = torch.minimum((frac * nc).floor().to(torch.int64).clamp_(min=0), nc - 1)
    H = torch.bincount((c[:, 2] * ny + c[:, 1]) * nx + c[:, 0], minlength=nx * ny * nz).to(torch.float64).view(nz, ny, nx)
    total = 0.0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                S = H
                ok = True
                for dim, (d, n, per) in enumerate(((dz, nz, periodic[2]), (dy, ny, periodic[1]), (dx, nx, periodic[0]))):
                    if d == 0:
                        continue
                    if per:
                        S = torch.roll(S, shifts=d, dims=dim)
                    elif n == 1:
                        ok = False
                    else:
                        Z = torch.zeros_like(S)
                        src = [slice(None)] * 3
                        dst = [slice(None)] * 3
                        src[dim] = slice(0, n - 1) if d == 1 else slice(1, n)
                        dst[dim] = slice(1, n) if d == 1 else slice(0, n - 1)
                        Z[tuple(dst)] = S[tuple(src)]
                        S = Z
                if ok:
                    total += float((H * S).sum())
    n = frac.shape[0]
    return (total - n) / 2.0 if half else total


# ---------------------------------------------------------------------------------------
# C2 (headline): InterRDF
# ---------------------------------------------------------------------------------------
def bench_c2(ctx):
    torch, args = ctx.torch, ctx.args
    from mdanalysis_b200 import FrameBlock, _cabi
    from mdanalysis_b200.analysis.rdf import InterRDF
    F, N = args.frames, args.atoms
    box = np.array([BOX_L] * 3 + [90.0] * 3, dtype=np.float32)
    pageable = make_frames(F, N, ctx.rank)
    host = torch.from_numpy(pageable).pin_memory()
    dev = host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    blk_dev = FrameBlock(dev, box)
    blk_host = FrameBlock(host.numpy(), box)
    blk_page = FrameBlock(pageable, box)
    result = {}

    def run(blk, key):
        result[key] = InterRDF(blk, None, nbins=NBINS, range=(0.0, RMAX), distributed=ctx.dist_ready).run()

    for _ in range(max(args.warmup, 3)):
        run(blk_dev, "rdf")
    clocks = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        clocks.start()
    _cabi.profile_report()
    _cabi.profile_enable(True)
    launches0 = _cabi.launch_count()
    ctx.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run(blk_dev, "rdf")
    e1.record()
    torch.cuda.synchronize()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = _cabi.launch_count() - launches0
    prof = _cabi.profile_report()
    _cabi.profile_enable(False)
    clock_info = clocks.stop() if ctx.rank == 0 else None
    frames_total = ctx.world * F * args.steps
    value = frames_total / (ms / 1e3)

    # dominant kernel: the fused pair search + histogram (one or more launches per step)
    kname = "rdf_pair_histogram"
    k_launches, k_ms = prof.get(kname, (0, 0.0))
    trav_ms = k_ms + prof.get("rdf_window_planner", (0, 0.0))[1]
    flop = ALGO_FLOP_PER_FRAME * (N / N_O) ** 2 * F * args.steps
    achieved = flop / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
    roofline = ctx.fp32(achieved)
    roofline.update({
        "kernel": kname, "traffic": ncu_traffic(F, N),
        "kernel_launches_per_step": k_launches / args.steps,
        "kernel_ms_per_step": k_ms / args.steps,
        "kernel_ms_per_launch": (k_ms / k_launches) if k_launches else None,
        "kernel_share_of_step": (k_ms / ms) if ms > 0 else None,
        "with_planner": {"ms_per_step": trav_ms / args.steps, "achieved": flop / (trav_ms / 1e3) / 1e12 if trav_ms else 0.0,
                         "frac": (flop / (trav_ms / 1e3) / 1e12 / FP32_MEASURED_TFLOPS) if trav_ms else 0.0,
                         "note": "pair kernel + window planner (the planner computes the per-lane windows the pair kernel walks)"},
        "algorithmic_flop_per_frame": ALGO_FLOP_PER_FRAME * (N / N_O) ** 2,
        "kernels_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in prof.items()}})

    # end to end through the public API with HOST buffers (H2D + D2H inside the region)
    def e2e_of(blk, key, api):
        dt = ctx.wall(lambda: run(blk, key), args.steps)
        return {"value": ctx.world * F / dt, "unit": "frames/s", "h2d_bytes_per_step": F * N * 12 + (NBINS + 1) * 8,
                "d2h_bytes_per_step": NBINS * 8, "ms_per_step": dt * 1e3, "api": api}

    e2e = e2e_of(blk_host, "rdf_host", "InterRDF(FrameBlock(pinned host ndarray), ...).run() -> mdagpu_rdf_accumulate(space=HOST)")
    e2e["pageable"] = e2e_of(blk_page, "rdf_page", "the same call on an ordinary (pageable) ndarray")
    same = bool(np.array_equal(result["rdf"].results.count, result["rdf_host"].results.count)
                and np.array_equal(result["rdf"].results.count, result["rdf_page"].results.count))

    # strong scaling of the STATED config: 1000 frames in total, np.array_split over the ranks
    lo, hi = shard(F, ctx.world, ctx.rank)
    blk_strong = FrameBlock(dev[lo:hi], box)
    _cabi.profile_report()
    _cabi.profile_enable(True)
    ms_s = ctx.timed(lambda: run(blk_strong, "rdf_strong"), args.steps, warm=2)
    prof_s = _cabi.profile_report()
    _cabi.profile_enable(False)
    kern_s = sum(v[1] for v in prof_s.values()) / (args.steps + 2)
    strong = {"frames_total": F, "frames_this_rank": hi - lo, "ms_per_step": ms_s, "value": F / (ms_s / 1e3),
              "unit": "frames/s", "kernels_ms_per_step_rank0": round(kern_s, 4),
              "host_and_collective_ms_per_step": round(ms_s - kern_s, 4),
              "note": "the stated 1000 frames split with np.array_split; one all_reduce of [count, volume, frames] per step; "
                      "what does not shrink with N is the host side of run() (launch + Python) and the collective"}

    # hits/s and candidates/s next to frames/s.  A hit = an ordered pair inside
    # the cutoff = what the histogram holds (self pairs included); candidates = the 1760 per query of the
    # cell = r_c / 2 traversal the algorithmic flop count is defined on (the fine-x walk evaluates ~550).
    hits_per_frame = float(np.sum(result["rdf"].results.count)) / (F * (ctx.world if ctx.dist_ready else 1))
    roofline["hits_per_s"] = hits_per_frame * value
    roofline["hits_per_frame"] = hits_per_frame
    roofline["algorithmic_candidates_per_s"] = (ALGO_FLOP_PER_FRAME / 11.0) * (N / N_O) ** 2 * value
    cpu_baseline = None
    if ctx.cpu:
        from oracle import oracle as orc
        orc.build()
        cpu_baseline = cpu_baseline_c2(orc, pageable[0], box, N)
    del blk_dev, blk_strong, dev, host, blk_host, blk_page
    torch.cuda.empty_cache()
    return {"value": value, "ms": ms, "roofline": roofline, "e2e": e2e, "launches": int(launches),
            "clocks": clock_info, "same": same, "strong": strong, "cpu_baseline": cpu_baseline}


def bench_c2_triclinic(ctx):
    """InterRDF in a triclinic cell (`row_tile_kernel`): same sites, 60/60/90 degrees."""
    torch = ctx.torch
    from mdanalysis_b200 import FrameBlock, _cabi
    from mdanalysis_b200.analysis.rdf import InterRDF
    from mdanalysis_b200.lib.mdamath import triclinic_vectors
    F, N = (200 if not ctx.quick else 20), ctx.args.atoms
    L = (N / 0.033334 / 0.70710678) ** (1.0 / 3.0)
    box = np.array([L, L, L, 60, 60, 90], dtype=np.float32)
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(SEED + 77 + ctx.rank)
    tv = torch.as_tensor(triclinic_vectors(box).astype(np.float64), device=ctx.dev)
    pos = (torch.rand((F, N, 3), generator=g, device=ctx.dev, dtype=torch.float64) @ tv).to(torch.float32).contiguous()
    blk = FrameBlock(pos, box)
    fn = lambda: InterRDF(blk, None, nbins=NBINS, range=(0.0, RMAX), distributed=ctx.dist_ready).run()
    _cabi.profile_report()
    _cabi.profile_enable(True)
    ms = ctx.timed(fn, 3, warm=1)
    prof = _cabi.profile_report()
    _cabi.profile_enable(False)
    k_ms = prof.get("rdf_pair_histogram", (0, 0.0))[1] / 4
    flop = ALGO_FLOP_PER_FRAME * (N / N_O) ** 2 * F
    out = {"frames_per_s": ctx.world * F / ms * 1e3, "ms": ms, "frames_per_gpu": F, "box": [float(v) for v in box],
           "roofline": ctx.fp32(flop / (k_ms / 1e3) / 1e12 if k_ms else 0.0,
                                "same algorithmic flop per frame as the orthorhombic headline"),
           "kernels_ms": {k: round(v[1] / 4, 4) for k, v in prof.items()}}
    del pos, blk
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------
# C1: distance_array + self_distance_array, adk (3341 atoms, 98 frames)
# ---------------------------------------------------------------------------------------
def adk_frames():
    """The 98 adk frames through the native DCD reader when the fixture travelled with
    baseline/_ref; otherwise adk-sized synthetic coordinates (stated in the entry)."""
    dcd = os.path.join(ROOT, "baseline", "_ref", "MDAnalysisTests", "data", "adk_dims.dcd")
    if os.path.exists(dcd):
        from mdanalysis_b200.coordinates.dcd import read_dcd
        return np.ascontiguousarray(read_dcd(dcd).positions), "adk_dims.dcd (98 frames, 3341 atoms) read by mdagpu_dcd_readframes"
    rng = np.random.default_rng(SEED0)
    return (rng.random((98, 3341, 3)) * 50.0).astype(np.float32), "synthetic adk-sized coordinates (fixture absent)"


C1_BOXES = {"none": None, "ortho": np.array([50, 50, 50, 90, 90, 90], dtype=np.float32),
            "triclinic": np.array([50, 50, 50, 70.53571, 109.48542, 70.518196], dtype=np.float32)}


def bench_c1(ctx):
    torch = ctx.torch
    from mdanalysis_b200 import _cabi
    from mdanalysis_b200.lib import distances as gd
    lib = _cabi.load()
    xyz_all, source = adk_frames()
    if ctx.quick:
        xyz_all = xyz_all[:8]
    Ft, n = xyz_all.shape[0], xyz_all.shape[1]
    out = {"data": source, "atoms": n, "frames_total": Ft,
           "sharding": "frames (np.array_split); no collective -- results stay on the rank that made them"}

    def one_pass(lo, hi, label):
        F = hi - lo
        res = {}
        if F == 0:
            F = 0
        x0 = torch.as_tensor(xyz_all[:1], device=ctx.dev).expand(max(F, 1), n, 3).contiguous()
        conf = torch.as_tensor(xyz_all[lo:hi] if F else xyz_all[:1], device=ctx.dev).contiguous()
        Fa = conf.shape[0]
        dbuf = torch.empty((Fa, n, n), dtype=torch.float64, device=ctx.dev)
        sbuf = torch.empty((Fa, n * (n - 1) // 2), dtype=torch.float64, device=ctx.dev)
        st = _cabi.current_stream(conf)
        for name, b6 in C1_BOXES.items():
            mode, b = gd._box_args(b6)
            ms = ctx.timed(lambda: _cabi.check(lib.mdagpu_distance_array(
                _cabi.tensor_ptr(x0), n, _cabi.tensor_ptr(conf), n, mode, _cabi.ptr(b), _cabi.tensor_ptr(dbuf), Fa, 1, st)), 3)
            pairs = ctx.sum_over_ranks(F * n * n)
            res[f"distance_array_{name}"] = {"pairs_per_s": pairs / ms * 1e3, "ms": ms,
                                             "roofline": ctx.hbm(pairs * 8 / ms / 1e6 / ctx.world)}
            ms = ctx.timed(lambda: _cabi.check(lib.mdagpu_self_distance_array(
                _cabi.tensor_ptr(conf), n, mode, _cabi.ptr(b), _cabi.tensor_ptr(sbuf), Fa, 1, st)), 3)
            pairs = ctx.sum_over_ranks(F * (n * (n - 1) // 2))
            res[f"self_distance_array_{name}"] = {"pairs_per_s": pairs / ms * 1e3, "ms": ms,
                                                  "roofline": ctx.hbm(pairs * 8 / ms / 1e6 / ctx.world)}
        del dbuf, sbuf, x0, conf
        torch.cuda.empty_cache()
        res["frames_this_rank"] = F
        res["note"] = label
        return res

    out["weak"] = one_pass(0, Ft, f"every rank: all {Ft} frames in one launch per function (outputs 8.75 + 4.37 GB > L2); "
                                  "roofline per GPU, 8 B written per pair")
    if ctx.world > 1:
        lo, hi = shard(Ft, ctx.world, ctx.rank)
        out["strong"] = one_pass(lo, hi, f"the {Ft} frames split over {ctx.world} ranks")

    # as a user calls it: one frame pair through the NumPy front end (host arrays in, host result out)
    x0, x1 = xyz_all[0], xyz_all[min(10, Ft - 1)]
    res_page = np.empty((n, n), dtype=np.float64)
    res_pin = torch.empty((n, n), dtype=torch.float64).pin_memory().numpy()
    e2e = {}
    for name, b6 in C1_BOXES.items():
        ent = {"pairs": n * n}
        for kind, res in (("pageable", res_page), ("pinned", res_pin)):
            dt = ctx.wall(lambda: gd.distance_array(x0, x1, box=b6, result=res), 5, warm=2)
            ent[f"{kind}_result_ms"] = dt * 1e3
            ent[f"{kind}_result_pairs_per_s"] = ctx.world * n * n / dt
        ent["h2d_bytes_per_call"] = 2 * n * 12
        ent["d2h_bytes_per_call"] = n * n * 8
        e2e[name] = ent
    e2e["api"] = "lib.distances.distance_array(ndarray, ndarray, box, result=ndarray): H2D + kernel + D2H of the 89 MB result"
    out["e2e"] = e2e

    if ctx.cpu:
        cb = {"cores": host_cores(), "unit": "pairs/s"}
        try:
            from oracle import oracle as orc
            orc.build()
            if orc.have_reference():
                for name, b6 in C1_BOXES.items():
                    t0 = time.perf_counter()
                    want = orc.distance_array(x0, x1, b6, impl="reference")
                    ts = time.perf_counter() - t0
                    ent = {"serial_ms": ts * 1e3, "serial_pairs_per_s": n * n / ts}
                    if orc.have_reference(openmp=True):
                        orc.distance_array(x0, x1, b6, impl="reference", openmp=True)
                        t0 = time.perf_counter()
                        orc.distance_array(x0, x1, b6, impl="reference", openmp=True)
                        tp = time.perf_counter() - t0
                        ent.update({"openmp_ms": tp * 1e3, "openmp_pairs_per_s": n * n / tp})
                    gd.distance_array(x0, x1, box=b6, result=res_page)
                    ent["identical_to_gpu_result"] = bool(np.array_equal(res_page, want))
                    cb[name] = ent
                cb["kind"] = "reference"
                cb["value"] = cb["none"]["serial_pairs_per_s"]
                cb["sample"] = ("oracle/_ref: the reference's calc_distances.h compiled with its own flags (serial and "
                                "-fopenmp), one call distance_array(frame 0, frame 10), 11.2 M pairs, per box type")
        except Exception as e:                                            # the baseline is optional
            cb["error"] = repr(e)
        out["cpu_baseline"] = cb
    return out


# ---------------------------------------------------------------------------------------
# C4: RMSD + AlignTraj (QCP), 10k atoms x 50k frames
# ---------------------------------------------------------------------------------------
def c4_trajectory(ctx, F, na, seed):
    """frame f = ref @ Q_f + t_f + N(0, 1) with a random rotation Q_f (normalised
    Gaussian quaternion) and translation t_f ~ N(0, 5^2); generated on the device in chunks."""
    torch = ctx.torch
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(SEED0 + 3)                                   # the reference structure is the same on all ranks
    ref = torch.randn((na, 3), generator=g, device=ctx.dev, dtype=torch.float64) * 20.0
    g.manual_seed(seed)
    traj = torch.empty((F, na, 3), dtype=torch.float32, device=ctx.dev)
    for a in range(0, F, 2000):
        b = min(F, a + 2000)
        q = torch.randn((b - a, 4), generator=g, device=ctx.dev, dtype=torch.float64)
        q = q / q.norm(dim=1, keepdim=True)
        w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        Q = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
        t = torch.randn((b - a, 1, 3), generator=g, device=ctx.dev, dtype=torch.float64) * 5.0
        fr = torch.matmul(ref[None], Q) + t
        fr += torch.randn(fr.shape, generator=g, device=ctx.dev, dtype=torch.float64)
        traj[a:b] = fr.to(torch.float32)
    return ref.to(torch.float32).cpu().numpy(), traj


def bench_c4(ctx):
    torch = ctx.torch
    from mdanalysis_b200 import FrameBlock, parallel
    from mdanalysis_b200.analysis.align import AlignTraj
    from mdanalysis_b200.analysis.rms import RMSD
    from mdanalysis_b200.lib import qcprot
    na = 10000
    Ft = 50000 if not ctx.quick else 2000
    ref_np, traj = c4_trajectory(ctx, Ft, na, SEED0 + 3 + 1000 * ctx.rank)
    masses = np.random.default_rng(SEED0 + 3).uniform(1, 32, na)
    w_mass = masses / masses.mean()
    out = {"atoms": na, "frames_total": Ft,
           "data": "synthetic: rotated + translated + noisy copies of a 10k-atom reference (SURVEY.md 8d), generated on the device",
           "sharding": "frames (np.array_split); all_gather of rmsd[F] float64 over NCCL (ndarray_vstack of the "
                       "reference's aggregator); aligned coordinates stay on their rank"}

    def one_pass(lo, hi, n_total, label):
        sub = traj[lo:hi]
        F = hi - lo
        res = {"frames_this_rank": F, "note": label}
        series = {}

        def rmsd_step(weights=None):
            vals, _, _ = qcprot.qcp_trajectory(sub, ref_np, weights=weights, convention=0)
            series["rmsd"] = parallel.allgather_series(vals, n_total) if ctx.dist_ready else vals

        frames_all = ctx.sum_over_ranks(F)
        ms = ctx.timed(rmsd_step, 5, warm=2)
        res["rmsd"] = {"frames_per_s": frames_all / ms * 1e3, "ms": ms,
                       "roofline": ctx.hbm(frames_all * na * 12 / ms / 1e6 / ctx.world)}
        assert series["rmsd"].shape[0] == (n_total if ctx.dist_ready else F)
        ms = ctx.timed(lambda: rmsd_step(w_mass), 5, warm=1)
        res["rmsd_mass_weighted"] = {"frames_per_s": frames_all / ms * 1e3, "ms": ms,
                                     "roofline": ctx.hbm(frames_all * na * 12 / ms / 1e6 / ctx.world)}

        def align_step():
            vals, _, _ = qcprot.qcp_trajectory(sub, ref_np, align=True, convention=1, out=sub)
            series["align"] = parallel.allgather_series(vals, n_total) if ctx.dist_ready else vals

        ms = ctx.timed(align_step, 3, warm=1)
        res["aligntraj"] = {"frames_per_s": frames_all / ms * 1e3, "ms": ms,
                            "roofline": ctx.hbm(frames_all * na * 24 / ms / 1e6 / ctx.world),
                            "note": "in place (in_memory=True): 12 B read + 12 B written per atom"}
        return res

    # AlignTraj rewrites the trajectory in place: keep the RMSD legs first, and re-generate for the strong pass
    out["weak"] = one_pass(0, Ft, Ft * ctx.world, f"every rank: its own {Ft} frames (6.0 GB > L2)")
    rm = qcprot.qcp_trajectory(traj[:64], ref_np, convention=0)[0]
    out["rmsd_after_alignment_first_frames"] = float(rm.mean().item())
    if ctx.world > 1:
        del traj
        torch.cuda.empty_cache()
        ref_np, traj = c4_trajectory(ctx, Ft, na, SEED0 + 3)          # the SAME 50k frames on every rank, each takes its block
        lo, hi = shard(Ft, ctx.world, ctx.rank)
        out["strong"] = one_pass(lo, hi, Ft, f"the stated {Ft} frames split over {ctx.world} ranks")

    # end to end: host trajectory -> RMSD(...).run() -> results.rmsd on the host
    Fe = min(Ft, 5000)
    del traj
    torch.cuda.empty_cache()
    ref_np, small = c4_trajectory(ctx, Fe, na, SEED0 + 3 + 1000 * ctx.rank)
    pin = torch.empty((Fe, na, 3), dtype=torch.float32).pin_memory()
    pin.copy_(small)
    page = pin.numpy().copy()
    del small
    e2e = {"frames": Fe, "h2d_bytes_per_step": Fe * na * 12, "d2h_bytes_per_step": Fe * 8,
           "api": "analysis.rms.RMSD(FrameBlock(host ndarray), ref).run() / analysis.align.AlignTraj(...).run(): "
                  "upload ring + kernel + D2H of the series (AlignTraj: + D2H of the aligned frames)"}
    for kind, arr in (("pinned", pin.numpy()), ("pageable", page)):
        blk = FrameBlock(arr)
        dt = ctx.wall(lambda: RMSD(blk, ref_np).run(), 3, warm=1)
        e2e[f"rmsd_{kind}_frames_per_s"] = ctx.world * Fe / dt
    blk = FrameBlock(pin.numpy())
    dt = ctx.wall(lambda: AlignTraj(blk, ref_np, in_memory=True).run(), 2, warm=1)
    e2e["aligntraj_pinned_frames_per_s"] = ctx.world * Fe / dt
    e2e["aligntraj_d2h_bytes_per_step"] = Fe * na * 12 + Fe * 8
    e2e["value"], e2e["unit"] = e2e["rmsd_pinned_frames_per_s"], "frames/s"
    out["e2e"] = e2e

    if ctx.cpu:
        out["cpu_baseline"] = cpu_baseline_c4(page, ref_np, masses)
    del pin, page
    return out


def cpu_baseline_c4(traj, ref_np, masses):
    """RMSD.run serial + multiprocessing, AlignTraj, bounded frame subsets."""
    ref = real_reference()
    if ref is None:
        return {"kind": "port", "error": "baseline/_ref absent"}
    mda = ref[0]
    from MDAnalysis.analysis import align, rms
    na = traj.shape[1]
    cores = host_cores()

    def universe(xyz):
        u = mda.Universe.empty(na, trajectory=True)
        u.add_TopologyAttr("masses", masses)
        u.load_new(xyz, order="fac")
        return u

    ru = universe(ref_np[None].copy())
    n1 = min(len(traj), 2000)
    u = universe(traj[:n1].copy())
    t0 = time.perf_counter()
    rms.RMSD(u, ru, select="all").run()
    t_ser = time.perf_counter() - t0
    out = {"kind": "reference", "unit": "frames/s", "cores": 1, "value": n1 / t_ser,
           "sample": f"MDAnalysis 2.8.0-dev0 (baseline/_ref): RMSD(u, ref, select='all').run() on {n1} in-memory frames "
                     f"of the same synthetic trajectory, serial: {t_ser:.2f} s"}
    n2 = min(len(traj), 400)
    try:
        u2 = universe(traj[:n2].copy())
        t0 = time.perf_counter()
        rms.RMSD(u2, ru, select="all").run(backend="multiprocessing", n_workers=min(cores, 16))
        t_mp = time.perf_counter() - t0
        out["multiprocessing"] = {"value": n2 / t_mp, "cores": min(cores, 16),
                                  "sample": f"run(backend='multiprocessing', n_workers={min(cores, 16)}) on {n2} frames: {t_mp:.2f} s "
                                            "(pickles the in-memory trajectory to every worker)"}
    except Exception as e:
        out["multiprocessing"] = {"error": repr(e)}
    n3 = min(len(traj), 1000)
    u3 = universe(traj[:n3].copy())
    t0 = time.perf_counter()
    align.AlignTraj(u3, ru, select="all", in_memory=True).run()
    t_al = time.perf_counter() - t0
    out["aligntraj"] = {"value": n3 / t_al, "cores": 1,
                        "sample": f"AlignTraj(u, ref, select='all', in_memory=True).run() on {n3} frames, serial (not "
                                  f"parallelizable in the reference): {t_al:.2f} s"}
    return out


# ---------------------------------------------------------------------------------------
# C3: FastNS self_search / capped_distance, 8 A, 1M atoms, triclinic -- replicas at N > 1
# ---------------------------------------------------------------------------------------
def bench_c3(ctx):
    torch = ctx.torch
    from mdanalysis_b200 import _cabi
    from mdanalysis_b200.lib import distances as gd
    from mdanalysis_b200.lib.mdamath import triclinic_vectors
    from mdanalysis_b200.lib.nsgrid import FastNS
    nn = 1_000_000 if not ctx.quick else 100_000
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(SEED0 + 2 + 1000 * ctx.rank)
    out = {"atoms": nn, "cutoff": 8.0, "sharding": "single frame: replicas only (each rank its own seed), no collective"}
    for label, angles in (("triclinic", (60, 60, 90)), ("ortho", (90, 90, 90))):
        vol = nn / 0.1
        sinf = 0.70710678 if label == "triclinic" else 1.0
        L = (vol / sinf) ** (1.0 / 3.0)
        box = np.array([L, L, L, *angles], dtype=np.float32)
        tv = torch.as_tensor(triclinic_vectors(box).astype(np.float64), device=ctx.dev)
        frac = torch.rand((nn, 3), generator=g, device=ctx.dev, dtype=torch.float64)
        pts = (frac @ tv).to(torch.float32).contiguous()
        grid = FastNS(8.0, pts, box)
        nc, cs, per, tri, _ = grid.grid_info()
        cand = stencil_candidates(torch, frac, nc, per, half=True)
        holder = {}

        def build():
            holder["g"] = FastNS(8.0, pts, box)

        def search():
            holder["r"] = None                                 # hand the previous result back to the pool first
            holder["r"] = holder["g"].self_search()

        build()
        search()                                               # warm-up: result pool growth
        npairs = len(holder["r"])
        _cabi.profile_report()
        _cabi.profile_enable(True)
        ms_build = ctx.timed(build, 3, warm=0, flush=True)
        ms_search = ctx.timed(search, 3, warm=2, flush=True)
        prof = _cabi.profile_report()
        _cabi.profile_enable(False)
        pairs_all = ctx.sum_over_ranks(npairs)
        call_ms = ms_search
        k = prof.get("pair_self_search")
        if k and k[0]:
            # the kernel itself (library events around the launch); the call adds the stream-ordered
            # allocation of the 1.7 GB result, which depends on the state of the pool
            ms_search = ctx.max_over_ranks(k[1] / k[0])
        ent = {"box": [float(v) for v in box], "pairs": npairs, "build_ms": ms_build, "search_ms": ms_search,
               "search_call_ms": call_ms,
               "pairs_per_s": pairs_all / ms_search * 1e3,
               "candidates": cand, "candidates_per_s": cand * ctx.world / ms_search * 1e3,
               "candidate_note": f"pairs inside the half 27-cell stencil of the grid the library reports ({list(map(int, nc))} cells), "
                                 "counted from the cell occupancies -- the set the reference's self_search walks",
               "roofline": ctx.fp32(cand * CANDIDATE_FLOP / (ms_search / 1e3) / 1e12,
                                    f"{CANDIDATE_FLOP} FP32 flop per candidate (filter); the ~10 % that pass are re-evaluated in float64"),
               "output_bound": ctx.hbm(npairs * 24 / ms_search / 1e6),
               "l2": "12 MB of coordinates < L2: every timed call follows a 512 MB L2 flush",
               "kernels_ms": {k: round(v[1] / v[0], 4) for k, v in prof.items() if v[0]}}
        ent["output_bound"]["note"] = "24 B written per pair (2 x int64 + float64 on export; the kernel writes 16 B)"
        del holder["r"]
        out[f"self_search_{label}"] = ent
        if label == "triclinic":
            # end to end: host ndarray -> pairs + distances on the host
            x_host = pts.cpu().numpy()
            res = {}

            def call():
                res["p"], res["d"] = gd.self_capped_distance(x_host, 8.0, box=box, method="nsgrid")

            dt = ctx.wall(call, 2, warm=1)
            np_e = len(res["p"])
            out["e2e"] = {"value": ctx.world * np_e / dt, "unit": "pairs/s", "ms_per_call": dt * 1e3,
                          "h2d_bytes_per_step": nn * 12, "d2h_bytes_per_step": np_e * 24,
                          "api": "lib.distances.self_capped_distance(ndarray, 8.0, box=box, method='nsgrid') -> (pairs int64[n,2], d float64[n]) on the host"}
            res.clear()
            if ctx.cpu:
                out["cpu_baseline"] = cpu_baseline_c3(x_host, box)
            del x_host
        del grid, pts, frac, holder
        torch.cuda.empty_cache()
    return out


def cpu_baseline_c3(x_host, box):
    """The reference's FastNS on a bounded sample: the first 1/10 of the box volume would change the
    geometry, so the sample is a 100k-atom system of the SAME density and cell shape."""
    ref = real_reference()
    n_s = 100_000
    L = (n_s / 0.1 / 0.70710678) ** (1.0 / 3.0)
    sbox = np.array([L, L, L, 60, 60, 90], dtype=np.float32)
    from mdanalysis_b200.lib.mdamath import triclinic_vectors
    rng = np.random.default_rng(SEED0 + 2)
    xs = (rng.random((n_s, 3)) @ triclinic_vectors(sbox).astype(np.float64)).astype(np.float32)
    if ref is None:
        from oracle import oracle as orc
        orc.build()
        t0 = time.perf_counter()
        p, d = orc.self_capped_distance(xs, 8.0, box=sbox, method="nsgrid")
        dt = time.perf_counter() - t0
        return {"kind": "port", "cores": 1, "unit": "pairs/s", "value": len(p) / dt,
                "sample": f"oracle port of FastNS.self_search, {n_s} atoms at the same density: {dt:.1f} s"}
    from MDAnalysis.lib.nsgrid import FastNS as RefNS
    t0 = time.perf_counter()
    gridsearch = RefNS(8.0, xs, sbox)
    t1 = time.perf_counter()
    r = gridsearch.self_search()
    t2 = time.perf_counter()
    p = r.get_pairs()
    d = r.get_pair_distances()
    t3 = time.perf_counter()
    return {"kind": "reference", "cores": 1, "unit": "pairs/s", "value": len(p) / (t3 - t0),
            "build_s": t1 - t0, "search_s": t2 - t1, "export_s": t3 - t2, "pairs": int(len(p)),
            "sample": f"MDAnalysis 2.8.0-dev0 (baseline/_ref) FastNS(8.0).self_search() + get_pairs/get_pair_distances on "
                      f"{n_s} atoms, same density (0.1 / A^3) and cell shape (60/60/90) as the 1M-atom config, 1 repetition "
                      "(the full config takes ~130 s, BASELINE.md section 2); serial only in the reference"}


# ---------------------------------------------------------------------------------------
# C5: self_distance_array contacts, 100k atoms, 16 frames sharded over the GPUs
# ---------------------------------------------------------------------------------------
def bench_c5(ctx):
    torch = ctx.torch
    from mdanalysis_b200 import _cabi, parallel
    from mdanalysis_b200.analysis.contacts import self_contacts
    lib = _cabi.load()
    nc = 100_000 if not ctx.quick else 20_000
    Ft = 16 if not ctx.quick else 4
    side = (nc / 0.01) ** (1.0 / 3.0)
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(SEED0 + 4)
    frames = (torch.rand((Ft, nc, 3), generator=g, device=ctx.dev) * side).contiguous()   # the same 16 frames on every rank
    obox = np.array([side] * 3 + [90.0] * 3, dtype=np.float32)
    out = {"atoms": nc, "frames_total": Ft, "radius": 8.0,
           "sharding": "frames (np.array_split), 2 per GPU at N = 8; all_gather of the per-frame contact counts over NCCL; "
                       "materialised distance arrays never leave their GPU"}
    cs_min = 8.0 * (1.0 + 1e-4) + 1e-3
    ncell = max(1, int(side // cs_min))
    cand = {"ortho": stencil_candidates(torch, (frames[0].double() / side).clamp_(0, 1 - 1e-12), (ncell,) * 3, (ncell >= 3,) * 3),
            "none": stencil_candidates(torch, (frames[0].double() / side).clamp_(0, 1 - 1e-12), (ncell,) * 3, (False,) * 3)}
    lo, hi = shard(Ft, ctx.world, ctx.rank)
    mine = frames[lo:hi].contiguous()
    F = hi - lo
    got = {}
    for name, bx in (("none", None), ("ortho", obox)):
        def step():
            c = self_contacts(mine, 8.0, box=bx)
            got[name] = parallel.allgather_series(c, Ft) if ctx.dist_ready else c

        ms = ctx.timed(step, 3, warm=1, flush=True)
        counts = got[name].cpu().numpy()
        flop = 8 if bx is None else 20
        out[f"contacts_{name}"] = {
            "frames_per_s": Ft / ms * 1e3, "ms": ms, "frames_this_rank": F,
            "contacts_frame0": int(counts[0]), "candidates_per_frame": cand[name],
            "candidates_per_s": cand[name] * Ft / ms * 1e3,
            "nominal_pairs_per_s": Ft * (nc * (nc - 1) // 2) / ms * 1e3,
            "roofline": ctx.fp32(cand[name] * F * flop / (ms / 1e3) / 1e12,
                                 f"evaluated candidates of the cell-list traversal ({ncell}^3 cells, half 27-cell stencil) x {flop} flop; "
                                 "nominal_pairs_per_s = n(n-1)/2 decisions per frame, NOT work done; the step is launch- and "
                                 "sort-bound at this size (build + count of one 100k-atom frame)"),
            "l2": "1.2 MB per frame < L2: every timed call follows a 512 MB L2 flush"}
    # (i) the materialised self_distance_array of one 100k-atom frame per rank (40 GB of float64, written once)
    npair = nc * (nc - 1) // 2
    big = torch.empty((npair,), dtype=torch.float64, device=ctx.dev)
    one = mine[:1].contiguous() if F else frames[:1].contiguous()
    for name, b6 in (("none", None), ("ortho", obox)):
        from mdanalysis_b200.lib import distances as gd
        mode, b = gd._box_args(b6)
        ms = ctx.timed(lambda: _cabi.check(lib.mdagpu_self_distance_array(
            _cabi.tensor_ptr(one), nc, mode, _cabi.ptr(b), _cabi.tensor_ptr(big), 1, 1, _cabi.current_stream(one))), 2)
        out[f"self_distance_array_{name}"] = {"pairs_per_frame": npair, "ms_per_frame": ms,
                                              "pairs_per_s": ctx.world * npair / ms * 1e3,
                                              "roofline": ctx.hbm(npair * 8 / ms / 1e6),
                                              "note": "one frame per rank at a time into a reused 40 GB buffer (weak: aggregate = N x)"}
    # the contact predicate on the materialised array agrees with the fused count (contacts.py:335-337)
    out["materialised_count_equals_fused"] = bool(int((big <= 8.0).sum().item()) == int(self_contacts(one, 8.0, box=obox)[0].item()))
    del big
    torch.cuda.empty_cache()

    # end to end: host frames -> counts on the host
    xh = mine.cpu().numpy() if F else frames[:1].cpu().numpy()
    dt = ctx.wall(lambda: self_contacts(xh, 8.0, box=obox), 3, warm=1)
    out["e2e"] = {"value": ctx.sum_over_ranks(len(xh)) / dt, "unit": "frames/s", "h2d_bytes_per_step": int(xh.nbytes),
                  "d2h_bytes_per_step": 8 * len(xh), "api": "analysis.contacts.self_contacts(ndarray [F, n, 3], 8.0, box) -> counts on the host"}
    if ctx.cpu:
        out["cpu_baseline"] = cpu_baseline_c5(frames[0].cpu().numpy(), obox)
    del frames, mine
    torch.cuda.empty_cache()
    return out


def cpu_baseline_c5(x, obox):
    """materialised self distances at n = 20 000 (serial + OpenMP) and the contact
    count of the full frame through the reference's self_capped_distance."""
    from oracle import oracle as orc
    orc.build()
    out = {"unit": "pairs/s", "cores": host_cores()}
    n_s = 20_000
    xs = np.ascontiguousarray(x[:n_s])
    if orc.have_reference():
        t0 = time.perf_counter()
        orc.self_distance_array(xs, None, impl="reference")
        ts = time.perf_counter() - t0
        out.update({"kind": "reference", "value": n_s * (n_s - 1) / 2 / ts, "serial_s": ts,
                    "sample": f"oracle/_ref (the reference's calc_distances.h): self_distance_array of {n_s} atoms (2.0e8 pairs, 1.6 GB), no box"})
        if orc.have_reference(openmp=True):
            t0 = time.perf_counter()
            orc.self_distance_array(xs, None, impl="reference", openmp=True)
            tp = time.perf_counter() - t0
            out["openmp_pairs_per_s"] = n_s * (n_s - 1) / 2 / tp
    ref = real_reference()
    if ref is not None:
        from MDAnalysis.lib.distances import self_capped_distance
        t0 = time.perf_counter()
        p = self_capped_distance(x, 8.0, box=obox, return_distances=False)
        dt = time.perf_counter() - t0
        out["contacts_full_frame"] = {"frames_per_s": 1.0 / dt, "pairs": int(len(p)), "cores": 1,
                                      "sample": f"MDAnalysis 2.8.0-dev0 self_capped_distance(x[100k], 8.0, box) (auto method): {dt:.2f} s per frame"}
    return out


# ---------------------------------------------------------------------------------------
# section 8f rank 1: batched ingest -- analyses on real MDAnalysis universes (DCD file, MemoryReader)
# ---------------------------------------------------------------------------------------
def bench_ingest(ctx):
    """run_on_gpu(<reference analysis object>) end to end: the reader -> FrameBlock ingest
    (native bulk DCD reader / zero-copy MemoryReader view), staged upload, kernel, results written
    into the reference object.  Needs the reference as the HOST framework (baseline/_ref)."""
    torch = ctx.torch
    ref = real_reference()
    if ref is None or ctx.rank != 0:
        return {"skipped": "baseline/_ref (MDAnalysis as the host framework) absent, or not rank 0"}
    mda = ref[0]
    from MDAnalysis.analysis import rms
    from mdanalysis_b200 import FrameBlock, _cabi, run_on_gpu
    from mdanalysis_b200.analysis.rms import RMSD
    out = {}
    data = os.path.join(ROOT, "baseline", "_ref", "MDAnalysisTests", "data")
    psf, dcd = os.path.join(data, "adk.psf"), os.path.join(data, "adk_dims.dcd")
    if os.path.exists(dcd):
        u = mda.Universe(psf, dcd)
        run_on_gpu(rms.RMSD(u, select="name CA"))
        t0 = time.perf_counter()
        for _ in range(5):
            got = run_on_gpu(rms.RMSD(u, select="name CA"))
        t_gpu = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        blk = FrameBlock.from_atomgroup(u.select_atoms("name CA"))
        t_ing = time.perf_counter() - t0
        t0 = time.perf_counter()
        want = rms.RMSD(u, select="name CA").run()
        t_ref = time.perf_counter() - t0
        out["dcd_adk_rmsd_ca"] = {"frames": len(u.trajectory), "atoms_read": blk.n_atoms, "gpu_call_ms": t_gpu * 1e3,
                                  "frames_per_s": len(u.trajectory) / t_gpu, "ingest_ms": t_ing * 1e3,
                                  "reference_ms": t_ref * 1e3, "reference_frames_per_s": len(u.trajectory) / t_ref,
                                  "max_abs_diff": float(np.abs(got.results.rmsd - want.results.rmsd).max()),
                                  "api": "run_on_gpu(MDAnalysis.analysis.rms.RMSD(Universe(adk.psf, adk_dims.dcd), select='name CA')): "
                                         "mdagpu_dcd_readframes (only the 214 CA atoms of the 98 frames) -> QCP kernel"}
    # a 50k-frame in-memory universe: the ingest must be a view, the upload staged, Python small
    na, F = 10000, (50000 if not ctx.quick else 4000)
    ref_np, traj = c4_trajectory(ctx, F, na, SEED0 + 3)
    arr = _cabi.download(traj)
    del traj
    torch.cuda.empty_cache()
    masses = np.random.default_rng(SEED0 + 3).uniform(1, 32, na)

    def universe(xyz):
        uu = mda.Universe.empty(na, trajectory=True)
        uu.add_TopologyAttr("masses", masses)
        uu.load_new(xyz, order="fac")
        return uu

    u, ru = universe(arr), universe(ref_np[None].copy())
    shares = u.trajectory.get_array() is arr or np.shares_memory(u.trajectory.get_array(), arr)
    run_on_gpu(rms.RMSD(u, ru, select="all"), stop=64)
    t0 = time.perf_counter()
    got = run_on_gpu(rms.RMSD(u, ru, select="all"))
    t_all = time.perf_counter() - t0
    blk = FrameBlock(arr)
    RMSD(blk, ref_np).run()
    t0 = time.perf_counter()
    core = RMSD(blk, ref_np).run()
    t_core = time.perf_counter() - t0
    out["memoryreader_rmsd"] = {
        "frames": F, "atoms": na, "host_bytes": int(arr.nbytes), "universe_shares_the_array": bool(shares),
        "run_on_gpu_ms": t_all * 1e3, "frames_per_s": F / t_all, "gb_per_s_from_host": arr.nbytes / t_all / 1e9,
        "block_call_ms": t_core * 1e3, "overhead_outside_block_call": (t_all - t_core) / t_all,
        "identical_to_block_call": bool(np.array_equal(got.results.rmsd[:, 2], core.results.rmsd[:, 2])),
        "api": "run_on_gpu(MDAnalysis.analysis.rms.RMSD(u, ref_u, select='all')) on a MemoryReader universe (pageable "
               "ndarray): zero-copy ingest, staged H2D (csrc/hostcopy.cu), QCP kernel, results into the reference object; "
               "block_call = analysis.rms.RMSD(FrameBlock(ndarray), ref).run() on the same array"}
    del arr, blk, u
    return out


# ---------------------------------------------------------------------------------------
# section 8f extras kept from round 1: all-pairs RMSD matrix
# ---------------------------------------------------------------------------------------
def bench_pairwise(ctx):
    torch = ctx.torch
    from mdanalysis_b200.lib import qcprot
    na, Fp = 10000, (2048 if not ctx.quick else 256)
    _, traj = c4_trajectory(ctx, Fp, na, SEED0 + 9)
    out = {}
    for label, sup, ops in (("pairwise_rmsd", False, 2 * 3), ("pairwise_rmsd_superposed", True, 9)):
        ms = ctx.timed(lambda: qcprot.pairwise_rmsd(traj, superposition=sup), 3)
        npair = Fp * (Fp + 1) // 2
        tops = npair * na * ops / ms / 1e9
        out[label] = {"frame_pairs_per_s": npair / ms * 1e3, "ms": ms, "shape": [Fp, na, 3],
                      "roofline": {"bound": "fp64", "achieved": tops, "peak": FP64_MEASURED_TOPS, "unit": "T FP64 lane-ops/s",
                                   "frac": tops / FP64_MEASURED_TOPS, "peak_kind": "measured DFMA issue rate (60.66 lanes/clk/SM at 1965 MHz)",
                                   "note": f"{ops} FP64 instructions per frame pair and atom"}}
    del traj
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the cuda backend has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries ONE JSON line: NCCL's banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Ctx(torch, dist, rank, world, local_rank, args)
    t_start = time.perf_counter()
    c2 = bench_c2(ctx)
    secondary = None
    if not args.no_secondary:
        secondary = {}
        wanted = [c.strip().upper() for c in args.configs.split(",")]
        table = (("C1", "C1_distance_arrays_adk", bench_c1), ("C2T", "C2_triclinic_interrdf", bench_c2_triclinic),
                 ("C3", "C3_fastns_1M", bench_c3), ("C4", "C4_rmsd_aligntraj", bench_c4),
                 ("C5", "C5_contacts_100k", bench_c5), ("PW", "pairwise_rmsd_matrix", bench_pairwise),
                 ("ING", "ingest_from_reader", bench_ingest))
        for key, name, fn in table:
            if key not in wanted or (key == "ING" and not ctx.cpu):
                continue
            t0 = time.perf_counter()
            try:
                secondary[name] = fn(ctx)
            except Exception as e:                             # one config must not take the line down
                import traceback
                traceback.print_exc()
                secondary[name] = {"error": repr(e)}
                if ctx.dist_ready:
                    raise
            secondary[name]["bench_wall_s"] = round(time.perf_counter() - t0, 2)
            torch.cuda.empty_cache()

    if rank == 0:
        F, N = args.frames, args.atoms
        line = {
            "metric": "InterRDF frames/s (O-O, 33334 sites, 150 bins to 15 A)", "value": c2["value"], "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": c2["ms"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 filter + f64 exact",
            "data": "synthetic",
            "config": {"workload": "C2 InterRDF O-O synthetic water box", "atoms": N, "frames_per_gpu": F, "box": BOX_L,
                       "nbins": NBINS, "rmax": RMAX, "l2": "inputs larger than L2 (400 MB per pass)",
                       "parallelism": f"frames sharded over {world} GPU(s); one NCCL all_reduce of [histogram, volume, frames]"},
            "roofline": c2["roofline"], "cpu_baseline": c2["cpu_baseline"], "e2e": c2["e2e"],
            "gpu_launches": c2["launches"], "clocks": c2["clocks"], "parity_device_vs_host_path": c2["same"],
            "strong": c2["strong"], "secondary": secondary,
            "bench_wall_s": round(time.perf_counter() - t_start, 1),
        }
        emit(line)
    if ctx.dist_ready:
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
    ap.add_argument("--configs", default="C1,C2T,C3,C4,C5,PW,ING", help="secondary configs to run")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
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
