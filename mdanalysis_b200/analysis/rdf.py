"""InterRDF on the GPU -- mirror of ``MDAnalysis/analysis/rdf.py`` (:86-348).

Per frame the reference does ``capped_distance(g1, g2, r_max, box)`` followed by
``np.histogram`` (rdf.py:268-291).  Here all frames of the block are binned by
ONE call of ``mdagpu_rdf_accumulate``; no pair list is ever materialised.

Exactness: for every frame the host decides which algorithm the reference would
have used (``lib.distances._determine_method``), because brute force and FastNS
differ in their minimum-image arithmetic; the kernel then evaluates the float64
squared distance of that algorithm bit for bit and compares it with thresholds
that are the exact pre-images of np.histogram's float64 edges under sqrt.  The
integer counts therefore equal the reference's.
"""
import numpy as np

from .._cabi import (DEVICE, HOST, MIC_ORTHO_LAZY, MIC_ORTHO_ROUND, MIC_TRICLINIC, check,
                     current_stream, is_tensor, load, ptr, tensor_ptr)
from ..frames import as_block
from ..lib.distances import _determine_method
from ..lib.mdamath import box_volume, triclinic_vectors
from .base import BlockAnalysis


def sqrt_ge_threshold(e):
    """Smallest float64 x with sqrt(x) >= e  (sqrt correctly rounded, monotone)."""
    e = np.float64(e)
    if e <= 0.0:
        return np.float64(0.0)
    x = e * e
    while np.sqrt(x) >= e:
        x = np.nextafter(x, -np.inf)
    while np.sqrt(x) < e:
        x = np.nextafter(x, np.inf)
    return x


def sqrt_le_threshold(e):
    """Largest float64 x with sqrt(x) <= e."""
    e = np.float64(e)
    if e < 0.0:
        return np.float64(-1.0)
    x = e * e
    while np.sqrt(x) <= e:
        x = np.nextafter(x, np.inf)
    while np.sqrt(x) > e:
        x = np.nextafter(x, -np.inf)
    return x


_THRESHOLD_CACHE = {}
_EDGE_CACHE = {}


def histogram_edges(settings):
    """``np.histogram([-1], **settings)`` as the reference calls it for its bin edges
    (analysis/rdf.py:253-257): ``(number of bins, edges)``, cached per setting (the call costs
    0.12 ms, a visible share of a short run)."""
    try:
        key = (settings["bins"] if np.isscalar(settings["bins"]) else None, tuple(map(float, settings["range"])))
    except (KeyError, TypeError, ValueError):
        key = (None,)
    if key[0] is None:
        count, edges = np.histogram([-1], **settings)
        return len(count), edges
    hit = _EDGE_CACHE.get(key)
    if hit is None:
        count, edges = np.histogram([-1], **settings)
        if len(_EDGE_CACHE) > 64:
            _EDGE_CACHE.clear()
        hit = _EDGE_CACHE[key] = (len(count), edges)
    return hit[0], hit[1].copy()


def histogram_thresholds(edges):
    """d2 thresholds equivalent to binning sqrt(d2) with np.histogram on ``edges``.

    Vectorised (the per-edge scalar search cost 1.3 ms per ``run()`` for 150 bins -- as much as
    the rest of the host side of a run) and cached per edge array: ``t[k]`` = smallest float64
    ``x`` with ``sqrt(x) >= edges[k]`` (``sqrt_ge_threshold``), ``t[-1]`` = largest ``x`` with
    ``sqrt(x) <= edges[-1]`` (``sqrt_le_threshold``).
    """
    e = np.ascontiguousarray(edges, dtype=np.float64)
    key = e.tobytes()
    hit = _THRESHOLD_CACHE.get(key)
    if hit is not None:
        return hit.copy()
    x = e * e
    # walk to the boundary: first down while the predecessor still satisfies sqrt >= e, then up
    # while x itself does not (a handful of steps: e * e is within an ulp or two of the boundary)
    for _ in range(64):
        prev = np.nextafter(x, -np.inf)
        with np.errstate(invalid="ignore"):
            move = (np.sqrt(prev) >= e) & (x > 0.0)
        if not move.any():
            break
        x = np.where(move, prev, x)
    for _ in range(64):
        move = np.sqrt(x) < e
        if not move.any():
            break
        x = np.where(move, np.nextafter(x, np.inf), x)
    t = np.where(e <= 0.0, 0.0, x)                        # sqrt_ge_threshold: edges <= 0 -> 0
    t[-1] = sqrt_le_threshold(e[-1])
    if len(_THRESHOLD_CACHE) > 64:
        _THRESHOLD_CACHE.clear()
    _THRESHOLD_CACHE[key] = t.copy()
    return t


class InterRDF(BlockAnalysis):
    """Radial distribution function between two groups of atoms.

    g1, g2   FrameBlock | [F, n, 3] array / CUDA tensor | AtomGroup with the same
             frames; pass the SAME object for a self RDF (pairs (i, i) are then
             counted in bin 0 unless ``exclusion_block=(1, 1)``, as in the reference)
    dimensions  unit cells [F, 6] / [6] when g1 is a bare array
    nbins, range, norm ('rdf' | 'density' | 'none'), exclusion_block: as rdf.py:220-251
    """

    def __init__(self, g1, g2, nbins=75, range=(0.0, 15.0), norm="rdf", exclusion_block=None,
                 exclude_same=None, dimensions=None, distributed=False, **kwargs):
        ids = self._resolve_exclude_same(g1, g2, exclude_same, exclusion_block)
        b1 = as_block(g1, dimensions)
        # distributed=True: this rank holds one shard of the frames; counts, volume
        # and frame count are summed over ranks (NCCL all_reduce on the device
        # histogram) before normalisation -- results.ndarray_sum of the reference.
        self._distributed = distributed
        super().__init__(b1, **kwargs)
        self._same = g2 is g1 or g2 is None
        self._b2 = b1 if self._same else as_block(g2, dimensions)
        if self._b2.n_frames != b1.n_frames:
            raise ValueError("g1 and g2 must have the same number of frames")
        self.norm = str(norm).lower()
        self.rdf_settings = {"bins": nbins, "range": range}
        self._exclusion_block = exclusion_block
        self._exclude_ids = ids
        if ids is not None and (len(ids[0]) != b1.n_atoms or len(ids[1]) != self._b2.n_atoms):
            raise ValueError("exclude_same ids must have one entry per atom of g1 and of g2")
        if self.norm not in ["rdf", "density", "none"]:
            raise ValueError(f"'{self.norm}' is an invalid norm. Use 'rdf', 'density' or 'none'.")

    @staticmethod
    def _resolve_exclude_same(g1, g2, exclude_same, exclusion_block):
        """exclude_same -> (ids of g1's atoms, ids of g2's atoms) as int32, or None.

        The reference takes 'residue' | 'segment' | 'chain' and reads ``resindices`` /
        ``segindices`` / ``chainIDs`` from the AtomGroups (rdf.py:245-247, 280-285); with bare
        arrays or FrameBlocks pass the two id arrays themselves: ``exclude_same=(ids1, ids2)``."""
        if exclude_same is None:
            return None
        if isinstance(exclude_same, str):
            if exclude_same not in ["residue", "segment", "chain"]:
                raise ValueError("The exclude_same argument to InterRDF must be None, 'residue', "
                                 "'segment' or 'chain'.")
            attr = {"residue": "resindices", "segment": "segindices", "chain": "chainIDs"}[exclude_same]
            g2 = g1 if g2 is None else g2
            if not (hasattr(g1, attr) and hasattr(g2, attr)):
                raise TypeError(f"exclude_same='{exclude_same}' needs AtomGroups with .{attr}; with arrays pass "
                                "exclude_same=(ids_of_g1_atoms, ids_of_g2_atoms)")
            a, b = np.asarray(getattr(g1, attr)), np.asarray(getattr(g2, attr))
        else:
            a, b = (np.asarray(x) for x in exclude_same)
        if exclusion_block is not None:
            raise ValueError("The exclude_same argument to InterRDF cannot be used with "
                             "exclusion_block.")
        _, inv = np.unique(np.concatenate([a, b]), return_inverse=True)     # any dtype (chain ids are strings)
        inv = inv.astype(np.int32)
        return np.ascontiguousarray(inv[:len(a)]), np.ascontiguousarray(inv[len(a):])

    def run(self, start=None, stop=None, step=None, frames=None, **kwargs):
        idx = self._slice(start, stop, step, frames)
        self._b2_run = self._b2 if self._same else self._b2.slice_frames(idx)
        return super().run(start, stop, step, frames, **kwargs)

    def _prepare(self, blk):
        ncount, edges = histogram_edges(self.rdf_settings)
        self.results.count = np.zeros(ncount, dtype=np.float64)
        self.results.edges = edges
        self.results.bins = 0.5 * (edges[:-1] + edges[1:])
        self.volume_cum = 0
        self._maxrange = self.rdf_settings["range"][1]

    # -- per-frame method of the reference, grouped into as few launches as possible --
    def _frame_modes(self, blk, na, nb):
        dims = blk.dimensions
        if dims is None:
            raise ValueError("InterRDF on the cuda backend needs unit-cell dimensions "
                             "(box=None frames are not supported)")
        uniq = dims[:1] if blk.constant_box() else dims
        modes = []
        ra = np.empty((na, 3), dtype=np.float32)     # only the lengths matter when a box is given
        rb = np.empty((nb, 3), dtype=np.float32)
        for box in uniq:
            method = _determine_method(ra, rb, self._maxrange, box=box)
            h9 = triclinic_vectors(box).reshape(9)
            if method == "nsgrid":
                tri = bool(h9[3] != 0 or h9[6] != 0 or h9[7] != 0)          # nsgrid.pyx:319-321
                modes.append((1, MIC_TRICLINIC if tri else MIC_ORTHO_LAZY, 1))
            else:
                ortho = bool(np.all(box[3:] == 90.0))                         # util.py:2522
                modes.append((0, MIC_ORTHO_ROUND if ortho else MIC_TRICLINIC, 0 if ortho else 1))
        if len(uniq) == 1:
            modes = modes * blk.n_frames
        return modes

    def _compute(self, blk):
        b2 = blk if self._same else self._b2_run
        na, nb = blk.n_atoms, b2.n_atoms
        nbins = len(self.results.edges) - 1
        lo, hi = float(self.results.edges[0]), float(self.results.edges[-1])
        thr = histogram_thresholds(self.results.edges)
        xa, xb = self._exclusion_block if self._exclusion_block is not None else (0, 0)
        modes = self._frame_modes(blk, na, nb)
        dev = blk.on_device
        if dev:
            import torch
            counts = torch.zeros(nbins, dtype=torch.int64, device=blk.positions.device)
        else:
            counts = np.zeros(nbins, dtype=np.uint64)
        rc2 = float(self._maxrange) * float(self._maxrange)           # nsgrid.pyx:540
        d2_brute = float(sqrt_le_threshold(self._maxrange))           # d <= max_cutoff
        kinds = sorted(set(modes))
        for mode in kinds:
            use_grid, mic, wrap = mode
            # (one mode for every frame -- a constant cell -- needs no per-frame comparison)
            fidx = np.arange(blk.n_frames) if len(kinds) == 1 else np.flatnonzero([m == mode for m in modes])
            whole = len(fidx) == blk.n_frames
            A = blk.positions if whole else blk.positions[fidx if not dev else torch.as_tensor(fidx, device=blk.positions.device)]
            B = None if self._same else (b2.positions if whole else b2.positions[fidx if not dev else torch.as_tensor(fidx, device=blk.positions.device)])
            dims = blk.dimensions if whole else blk.dimensions[fidx]
            const = bool((dims == dims[0]).all())
            box6 = np.ascontiguousarray(dims[:1] if const else dims, dtype=np.float32)
            h9 = np.ascontiguousarray([triclinic_vectors(b).reshape(9) for b in box6], dtype=np.float32)
            d2_max = rc2 if use_grid else d2_brute
            if self._exclude_ids is not None:
                ida, idb = self._exclude_ids
                if dev:
                    A = A.contiguous()
                    Bp = None if B is None else tensor_ptr(B.contiguous())
                    Ap, cp, sp_, strm = tensor_ptr(A), tensor_ptr(counts), DEVICE, current_stream(A)
                else:
                    A = np.ascontiguousarray(A)
                    B = None if B is None else np.ascontiguousarray(B)
                    Ap, Bp, cp, sp_, strm = ptr(A), ptr(B), ptr(counts), HOST, None
                check(load().mdagpu_rdf_accumulate_ids(
                    Ap, na, Bp, nb, len(fidx), ptr(box6), ptr(h9), int(not const), mic, wrap,
                    float(self._maxrange), d2_max, use_grid, ptr(thr), nbins, lo, hi, ptr(ida), ptr(idb),
                    cp, sp_, strm))
                continue
            if dev:
                A = A.contiguous()
                Bp = None if B is None else tensor_ptr(B.contiguous())
                rc = load().mdagpu_rdf_accumulate(
                    tensor_ptr(A), na, Bp, nb, len(fidx), ptr(box6), ptr(h9), int(not const), mic,
                    wrap, float(self._maxrange), d2_max, use_grid, ptr(thr), nbins, lo, hi, int(xa),
                    int(xb), tensor_ptr(counts), DEVICE, current_stream(A))
            else:
                A = np.ascontiguousarray(A)
                B = None if B is None else np.ascontiguousarray(B)
                rc = load().mdagpu_rdf_accumulate(
                    ptr(A), na, ptr(B), nb, len(fidx), ptr(box6), ptr(h9), int(not const), mic, wrap,
                    float(self._maxrange), d2_max, use_grid, ptr(thr), nbins, lo, hi, int(xa),
                    int(xb), ptr(counts), HOST, None)
            check(rc)
        # rdf.py:291 adds ts.volume frame by frame: the same sequential float64 sum, without a
        # Python statement per frame (a constant cell needs one box_volume call)
        if blk.n_frames == 0:
            vols = np.zeros(0)
        elif blk.constant_box():
            vols = np.full(blk.n_frames, box_volume(blk.dimensions[0]), dtype=np.float64)
        else:
            vols = np.array([box_volume(b) for b in blk.dimensions], dtype=np.float64)
        if len(vols):
            self.volume_cum = float(np.add.accumulate(np.concatenate(([float(self.volume_cum)], vols)))[-1])
        if self._distributed:
            counts = self._allreduce(counts, dev, blk)
        self._device_counts = counts if dev else None
        host_counts = counts.cpu().numpy() if dev else counts
        self.results.count += host_counts.astype(np.float64)
        self._nA, self._nB = na, nb

    def _allreduce(self, counts, dev, blk):
        """Sum histogram, cumulative volume and frame count over all ranks in ONE collective:
        ``[count[nbins], volume_cum, n_frames]`` as float64 (counts are integers below 2**53, so
        their float64 sum is exact -- the reference's own ``count`` array is float64, rdf.py:262)."""
        import torch
        from .. import parallel
        if not parallel.is_distributed():
            return counts
        if dev:
            device = counts.device
            local = counts
        else:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.distributed.get_backend() == "nccl" else "cpu"
            local = torch.as_tensor(counts.astype(np.int64)).to(device)
        nb = local.shape[0]
        packed = torch.empty(nb + 2, dtype=torch.float64, device=device)
        packed[:nb] = local
        packed[nb:] = torch.tensor([float(self.volume_cum), float(self.n_frames)], dtype=torch.float64)
        parallel.allreduce_sum(packed)
        tail = packed[nb:].cpu()
        if float(packed[:nb].max()) >= 2.0 ** 53:
            raise OverflowError("histogram counts exceed 2**53: the float64 reduction is not exact")
        self.volume_cum = float(tail[0])
        self.n_frames = int(round(float(tail[1])))
        total = packed[:nb].to(torch.int64)
        return total if dev else total.cpu().numpy().astype(np.uint64)

    def _conclude(self):
        norm = self.n_frames
        if self.norm in ["rdf", "density"]:
            vols = np.power(self.results.edges, 3)
            norm = norm * (4 / 3 * np.pi * np.diff(vols))
        if self.norm == "rdf":
            N = self._nA * self._nB
            if self._exclusion_block:
                xA, xB = self._exclusion_block
                nblocks = self._nA / xA
                N -= xA * xB * nblocks
            box_vol = self.volume_cum / self.n_frames
            norm = norm * (N / box_vol)
        self.results.rdf = self.results.count / norm


class InterRDF_s(BlockAnalysis):
    """Site-specific radial distribution functions (``analysis/rdf.py:351-678``).

    u      FrameBlock | [F, n, 3] array / CUDA tensor holding every atom (with unit cells)
    ags    list of ``[sites1, sites2]`` pairs of index arrays into ``u``
    Results as the reference: ``results.count[k]`` / ``results.rdf[k]`` have shape
    ``(len(sites1), len(sites2), nbins)``, ``results.indices[k] = [sites1, sites2]``,
    ``get_cdf()`` the cumulative counts.  All frames of a pair of groups are binned by one
    call (``mdagpu_rdf_sites_accumulate``: per-site histograms in device memory).
    """

    def __init__(self, u, ags, nbins=75, range=(0.0, 15.0), norm="rdf", density=False, dimensions=None,
                 **kwargs):
        blk = as_block(u, dimensions)
        super().__init__(blk, **kwargs)
        self.ags = [[np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64)] for a, b in ags]
        self.norm = str(norm).lower()
        self.rdf_settings = {"bins": nbins, "range": range}
        if self.norm not in ["rdf", "density", "none"]:
            raise ValueError(f"'{self.norm}' is an invalid norm. Use 'rdf', 'density' or 'none'.")
        if density:
            self.norm = "density"

    def _prepare(self, blk):
        ncount, edges = histogram_edges(self.rdf_settings)
        self.results.count = [np.zeros((len(a), len(b), ncount), dtype=np.float64) for a, b in self.ags]
        self.results.edges = edges
        self.results.bins = 0.5 * (edges[:-1] + edges[1:])
        self.volume_cum = 0
        self._maxrange = self.rdf_settings["range"][1]

    _frame_modes = InterRDF._frame_modes

    def _compute(self, blk):
        nbins = len(self.results.edges) - 1
        lo, hi = float(self.results.edges[0]), float(self.results.edges[-1])
        thr = histogram_thresholds(self.results.edges)
        rc2 = float(self._maxrange) * float(self._maxrange)
        d2_brute = float(sqrt_le_threshold(self._maxrange))
        dev = blk.on_device
        for k, (ia, ib) in enumerate(self.ags):
            A_all, B_all = blk.select(ia).positions, blk.select(ib).positions
            na, nb = len(ia), len(ib)
            if na == 0 or nb == 0:
                continue
            modes = self._frame_modes(blk, na, nb)
            if dev:
                import torch
                counts = torch.zeros((na, nb, nbins), dtype=torch.int64, device=A_all.device)
            else:
                counts = np.zeros((na, nb, nbins), dtype=np.uint64)
            for mode in sorted(set(modes)):
                use_grid, mic, wrap = mode
                fidx = np.flatnonzero([m == mode for m in modes])
                whole = len(fidx) == blk.n_frames
                sel = fidx if not dev else None
                if dev and not whole:
                    import torch
                    sel = torch.as_tensor(fidx, device=A_all.device)
                A = A_all if whole else A_all[sel]
                B = B_all if whole else B_all[sel]
                dims = blk.dimensions if whole else blk.dimensions[fidx]
                const = bool((dims == dims[0]).all())
                box6 = np.ascontiguousarray(dims[:1] if const else dims, dtype=np.float32)
                h9 = np.ascontiguousarray([triclinic_vectors(b).reshape(9) for b in box6], dtype=np.float32)
                d2_max = rc2 if use_grid else d2_brute
                if dev:
                    A, B = A.contiguous(), B.contiguous()
                    rc = load().mdagpu_rdf_sites_accumulate(
                        tensor_ptr(A), na, tensor_ptr(B), nb, len(fidx), ptr(box6), ptr(h9), int(not const), mic,
                        wrap, float(self._maxrange), d2_max, use_grid, ptr(thr), nbins, lo, hi,
                        tensor_ptr(counts), DEVICE, current_stream(A))
                else:
                    A, B = np.ascontiguousarray(A), np.ascontiguousarray(B)
                    rc = load().mdagpu_rdf_sites_accumulate(
                        ptr(A), na, ptr(B), nb, len(fidx), ptr(box6), ptr(h9), int(not const), mic, wrap,
                        float(self._maxrange), d2_max, use_grid, ptr(thr), nbins, lo, hi, ptr(counts), HOST, None)
                check(rc)
            host_counts = counts.cpu().numpy() if dev else counts
            self.results.count[k] += host_counts.astype(np.float64)
        if self.norm == "rdf":
            for box in blk.dimensions:                   # rdf.py:599-600
                self.volume_cum += box_volume(box)

    def _conclude(self):
        norm = self.n_frames
        if self.norm in ["rdf", "density"]:
            vols = np.power(self.results.edges, 3)
            norm = norm * (4 / 3 * np.pi * np.diff(vols))
        if self.norm == "rdf":
            norm = norm * (1 / (self.volume_cum / self.n_frames))
        self.results.indices = []
        self.results.rdf = []
        for k, (ia, ib) in enumerate(self.ags):
            self.results.indices.append([ia, ib])
            self.results.rdf.append(self.results.count[k] / norm)

    def get_cdf(self):
        """Cumulative counts per site pair, averaged over frames (rdf.py:619-640)."""
        self.results.cdf = [np.cumsum(c, axis=2) / self.n_frames for c in self.results.count]
        return self.results.cdf
