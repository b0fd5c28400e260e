"""Run an already-constructed MDAnalysis analysis object on the GPU.

The reference's analyses loop over frames in Python (``analysis/base.py:515-523``) and the
classes on this path that are most worth accelerating (``InterRDF``, ``AlignTraj``,
``Contacts``) refuse every non-serial executor (``_analysis_algorithm_is_parallelizable =
False``, ``base.py:653-654``), so a ``BackendBase`` subclass cannot reach them.  This module
is the caller-level plug instead: user code keeps building the reference object from
AtomGroups exactly as before and replaces ``.run(...)`` by ``run_on_gpu(obj, ...)``::

    rdf = InterRDF(ow, ow, nbins=150, range=(0, 15))       # MDAnalysis, unchanged
    mdanalysis_b200.run_on_gpu(rdf, step=10)               # instead of rdf.run(step=10)
    rdf.results.rdf                                        # filled as the reference fills it

What happens: the frames of the AtomGroups are read in bulk (``FrameBlock.from_atomgroup`` ->
``ReaderBase.timeseries``), the matching class of ``mdanalysis_b200.analysis`` is configured
from the attributes the reference constructor left on the object, all frames go through the
CUDA kernels in one call, and ``results`` / ``frames`` / ``times`` / ``n_frames`` are written
back onto the reference object.  Supported: ``rdf.InterRDF``, ``rdf.InterRDF_s``, ``rms.RMSD``,
``align.AlignTraj`` and ``align.AverageStructure`` (in-memory trajectories), ``contacts.Contacts``,
``diffusionmap.DistanceMatrix`` (metric ``rms.rmsd``, optionally ``partial(rmsd, superposition=True)``).
"""
from __future__ import annotations

import numpy as np

from .frames import FrameBlock


def _frames_of(analysis, start, stop, step, frames):
    n = len(analysis._trajectory)
    if frames is not None:
        if not all(opt is None for opt in (start, stop, step)):
            raise ValueError("start/stop/step cannot be combined with frames")
        idx = np.asarray(frames)
        return np.flatnonzero(idx) if idx.dtype == bool else idx
    return np.arange(n)[slice(start, stop, step)]


def _block(atomgroup, idx):
    """The atoms of ``atomgroup`` for the frames ``idx`` only: the frame range (and the atom
    selection) are pushed down to the reader, nothing else is read."""
    idx = np.asarray(idx, dtype=np.int64)
    if len(idx) > 1 and np.all(np.diff(idx) == idx[1] - idx[0]) and idx[1] > idx[0]:
        st = int(idx[1] - idx[0])
        return FrameBlock.from_atomgroup(atomgroup, start=int(idx[0]), stop=int(idx[-1]) + 1, step=st)
    if len(idx) == 1:
        return FrameBlock.from_atomgroup(atomgroup, start=int(idx[0]), stop=int(idx[0]) + 1, step=1)
    return FrameBlock.from_atomgroup(atomgroup, frames=idx)


def _compact(universe, index_lists):
    """AtomGroup of the union of ``index_lists`` and the lists re-expressed in its numbering."""
    used = np.unique(np.concatenate([np.asarray(ix, dtype=np.int64) for ix in index_lists]))
    return universe.atoms[used], [np.searchsorted(used, np.asarray(ix, dtype=np.int64)) for ix in index_lists]


def _finish(analysis, gpu):
    analysis.n_frames = gpu.n_frames
    analysis.frames = np.asarray(gpu.frames).copy()
    analysis.times = np.asarray(gpu.times).copy()
    return analysis


def _run_interrdf(an, idx, device):
    from .analysis.rdf import InterRDF
    ids = None
    if getattr(an, "exclude_same", None) is not None:       # the attribute name: resindices / segindices / chainIDs
        ids = (np.asarray(getattr(an.g1, an.exclude_same)), np.asarray(getattr(an.g2, an.exclude_same)))
    b1 = _block(an.g1, idx)
    same = an.g2 is an.g1 or (len(an.g1) == len(an.g2) and np.array_equal(an.g1.indices, an.g2.indices))
    b2 = b1 if same else _block(an.g2, idx)
    if device:
        b1 = b1.to_device()
        b2 = b1 if same else b2.to_device()
    gpu = InterRDF(b1, b2, nbins=an.rdf_settings["bins"], range=an.rdf_settings["range"], norm=an.norm,
                   exclusion_block=an._exclusion_block, exclude_same=ids).run()
    for k in ("count", "edges", "bins", "rdf"):
        an.results[k] = gpu.results[k]
    an.volume_cum = gpu.volume_cum
    return _finish(an, gpu)


def _run_interrdf_s(an, idx, device):
    from .analysis.rdf import InterRDF_s
    u = an.ags[0][0].universe
    blk = _block(u.atoms, idx)
    if device:
        blk = blk.to_device()
    sites = [[a.indices, b.indices] for a, b in an.ags]
    gpu = InterRDF_s(blk, sites, nbins=an.rdf_settings["bins"], range=an.rdf_settings["range"], norm=an.norm).run()
    for k in ("count", "edges", "bins", "rdf", "indices"):
        an.results[k] = gpu.results[k]
    an.volume_cum = gpu.volume_cum
    return _finish(an, gpu)


def _run_rmsd(an, idx, device):
    from .analysis.rms import RMSD
    mob_u, ref_u = an.atomgroup.universe, an.reference.universe
    # the reference structure is a snapshot of frame ref_frame, taken before anything else moves a reader
    cur = ref_u.trajectory.ts.frame
    try:
        ref_u.trajectory[an.ref_frame]
        ref_pos = ref_u.atoms.positions.copy()
    finally:
        ref_u.trajectory[cur]
    gsel = getattr(an, "_groupselections_atoms", [])
    # only the atoms that take part are read: the fitted selection and the group selections
    mob_ag, mob_lists = _compact(mob_u, [an.mobile_atoms.indices] + [g["mobile"].indices for g in gsel])
    ref_used, ref_lists = _compact(ref_u, [an.ref_atoms.indices] + [g["reference"].indices for g in gsel])
    blk = _block(mob_ag, idx)
    blk.masses = np.asarray(mob_ag.masses, dtype=np.float64) if hasattr(mob_u.atoms, "masses") else None
    refblk = FrameBlock(ref_pos[ref_used.indices][None],
                        masses=np.asarray(ref_used.masses, dtype=np.float64) if hasattr(ref_u.atoms, "masses") else None)
    if device:
        blk = blk.to_device()
    groups = [{"mobile": m, "reference": r} for m, r in zip(mob_lists[1:], ref_lists[1:])]
    gpu = RMSD(blk, refblk, select={"mobile": mob_lists[0], "reference": ref_lists[0]},
               groupselections=groups or None, weights=an.weights,
               weights_groupselections=an.weights_groupselections, tol_mass=an.tol_mass, ref_frame=0).run()
    an.results.rmsd = gpu.results.rmsd
    return _finish(an, gpu)


def _run_aligntraj(an, idx, device):
    from .analysis.align import AlignTraj
    if an.filename is not None:
        raise NotImplementedError("AlignTraj on the GPU path rewrites an in-memory trajectory (in_memory=True)")
    mob_u, ref_u = an.mobile.universe, an.ref_atoms.universe
    # AlignTraj._prepare fits to the reference atoms as they are when run() is called
    # (analysis/align.py:836-845): snapshot them before the bulk read touches a reader
    refblk = FrameBlock(ref_u.atoms.positions.copy()[None])
    traj = mob_u.trajectory
    arr = traj.get_array()                                   # MemoryReader: the coordinates themselves
    order = traj.stored_order
    blk = _block(mob_u.atoms, idx)
    if device:
        blk = blk.to_device()
    w = an._weights
    gpu = AlignTraj(blk, refblk, select={"mobile": an.mobile_atoms.indices, "reference": an.ref_atoms.indices},
                    weights=None if w is None else np.asarray(w, dtype=np.float64), in_memory=True).run()
    moved = gpu.results.positions
    moved = moved.detach().cpu().numpy() if hasattr(moved, "detach") else np.asarray(moved)
    view = np.transpose(arr, [order.find(c) for c in "fac"])      # [F, n, 3] view of the reader's array
    view[idx] = moved
    an.results.rmsd = np.asarray(gpu.results.rmsd)
    return _finish(an, gpu)


def _run_average_structure(an, idx, device):
    from .analysis.align import AverageStructure
    if an.filename is not None:
        raise NotImplementedError("AverageStructure on the GPU path needs in_memory=True")
    mob_u, ref_u = an.mobile.universe, an.ref_atoms.universe
    traj = mob_u.trajectory
    arr, order = traj.get_array(), traj.stored_order
    blk = _block(mob_u.atoms, idx)
    cur = ref_u.trajectory.ts.frame
    try:
        ref_u.trajectory[an.ref_frame]
        refblk = FrameBlock(ref_u.atoms.positions.copy()[None])
    finally:
        ref_u.trajectory[cur]
    if device:
        blk = blk.to_device()
    w = an._weights
    gpu = AverageStructure(blk, refblk, select={"mobile": an.mobile_atoms.indices, "reference": an.ref_atoms.indices},
                           weights=None if w is None else np.asarray(w, dtype=np.float64), in_memory=True).run()
    moved = blk.positions
    moved = moved.detach().cpu().numpy() if hasattr(moved, "detach") else np.asarray(moved)
    view = np.transpose(arr, [order.find(c) for c in "fac"])
    view[idx] = moved
    an.results.positions = gpu.results.positions
    an.results.rmsd = gpu.results.rmsd
    an.results.universe.load_new(an.results.positions.reshape((1, -1, 3)))      # align.py:1085-1087
    return _finish(an, gpu)


def _run_contacts(an, idx, device):
    from .analysis.contacts import pair_list_distances
    A, B = _block(an.grA, idx), _block(an.grB, idx)
    dims = A.dimensions if an.pbc else None
    Ap, Bp = (A.to_device().positions, B.to_device().positions) if device else (A.positions, B.positions)
    ts = np.empty((A.n_frames, len(an.r0) + 1))
    ts[:, 0] = A.frames
    for k, (mask, r0) in enumerate(zip(an.initial_contacts, an.r0), 1):
        ii, jj = np.nonzero(mask)                                 # row-major: the order of d[mask]
        r = pair_list_distances(Ap, Bp, ii, jj, dims)
        r = r.detach().cpu().numpy() if hasattr(r, "detach") else r
        r0s = r0[mask]
        with np.errstate(invalid="ignore", divide="ignore"):
            for f in range(A.n_frames):
                ts[f, k] = an.fraction_contacts(r[f], r0s, **an.fraction_kwargs)
    an.results.timeseries = ts
    an.n_frames = A.n_frames
    an.frames = np.asarray(A.frames).copy()
    an.times = np.asarray(A.times).copy()
    return an


def _run_distance_matrix(an, idx, device):
    """diffusionmap.DistanceMatrix (diffusionmap.py:226-289): atoms, _metric, _cutoff, _weights."""
    from .analysis.diffusionmap import DistanceMatrix
    blk = _block(an.atoms, idx)
    if device:
        blk = blk.to_device()
    gpu = DistanceMatrix(blk, metric=an._metric, cutoff=an._cutoff, weights=an._weights).run()
    an.results.dist_matrix = gpu.results.dist_matrix
    an._calculated = True
    return _finish(an, gpu)


_RUNNERS = {"InterRDF": _run_interrdf, "DistanceMatrix": _run_distance_matrix, "InterRDF_s": _run_interrdf_s, "RMSD": _run_rmsd,
            "AlignTraj": _run_aligntraj, "AverageStructure": _run_average_structure, "Contacts": _run_contacts}


def run_on_gpu(analysis, start=None, stop=None, step=None, frames=None, device=True):
    """``analysis.run(start, stop, step, frames)`` for a supported MDAnalysis analysis object,
    computed by the CUDA backend.  Returns ``analysis`` with ``results`` filled.

    device=True keeps the frame block on the GPU for the whole run (one H2D copy); False hands
    host arrays to the C ABI (which streams them through its upload ring).
    """
    name = type(analysis).__name__
    runner = _RUNNERS.get(name)
    if runner is None or not type(analysis).__module__.startswith("MDAnalysis.analysis"):
        raise TypeError(f"run_on_gpu(): no GPU path for {type(analysis).__module__}.{name}; "
                        f"supported: {sorted(_RUNNERS)}")
    idx = _frames_of(analysis, start, stop, step, frames)
    return runner(analysis, idx, device)


def _aggregators():
    """How the per-part results of each supported class merge (``ResultsGroup.merge``,
    ``analysis/results.py:185-226``): histograms add up, per-frame series stack."""
    from MDAnalysis.analysis.results import ResultsGroup as RG

    def same(arrs):                      # identical in every part (bin edges and centres)
        return arrs[0]

    def sum_lists(parts):                # InterRDF_s: a list of [n_i, n_j, nbins] arrays per part
        return [np.sum(np.array(x), axis=0) for x in zip(*parts)]

    def first_list(parts):
        return parts[0]

    return {
        "RMSD": {"rmsd": RG.ndarray_vstack},
        "InterRDF": {"count": RG.ndarray_sum, "volume_cum": RG.ndarray_sum, "edges": same, "bins": same},
        "InterRDF_s": {"count": sum_lists, "volume_cum": RG.ndarray_sum, "edges": same, "bins": same,
                       "indices": first_list},
        "Contacts": {"timeseries": RG.ndarray_vstack},
        "AlignTraj": {"rmsd": RG.ndarray_hstack},
    }


def gpu_enabled(analysis):
    """Make a constructed reference analysis acceptable to ``AnalysisBase.run(backend=...)``.

    ``InterRDF``, ``InterRDF_s``, ``AlignTraj`` and ``Contacts`` declare
    ``_analysis_algorithm_is_parallelizable = False`` and ``_configure_backend`` turns every
    non-serial executor away (``analysis/base.py:590-654``).  This is the plug point of SURVEY.md
    section 8(b)4: the object is re-classed to a subclass that sets the flag, names the built-in
    backends as supported (so a custom executor needs ``unsupported_backend=True``, as for ``RMSD``),
    supplies ``_get_aggregator`` (``ndarray_sum`` for counts, ``ndarray_vstack`` / ``hstack`` for
    series) and a ``_conclude`` that works on the merged results.  Usage::

        rdf = gpu_enabled(InterRDF(ow, ow, nbins=150, range=(0, 15)))
        rdf.run(backend=cuda_backend(8), unsupported_backend=True)
    """
    from MDAnalysis.analysis.results import ResultsGroup

    cls = type(analysis)
    name = cls.__name__
    lookup = _aggregators().get(name)
    if lookup is None:
        raise TypeError(f"gpu_enabled(): {cls.__module__}.{name} has no frame-parallel GPU path; "
                        f"supported: {sorted(_aggregators())}")
    if getattr(cls, "_mdagpu_enabled", False):
        return analysis

    def get_supported_backends(klass):
        return ("serial", "multiprocessing", "dask")

    def _get_aggregator(self):
        return ResultsGroup(lookup=dict(lookup))

    def _conclude(self):
        if name in ("InterRDF", "InterRDF_s"):
            self.volume_cum = float(np.sum(self.results.pop("volume_cum")))
            return cls._conclude(self)
        if name == "AlignTraj":
            self._writer.close()
            return None
        if name == "Contacts":
            self.results.timeseries = np.asarray(self.results.timeseries, dtype=float)
            return None
        return cls._conclude(self)

    sub = type(name, (cls,), {"_analysis_algorithm_is_parallelizable": True, "_mdagpu_enabled": True,
                              "get_supported_backends": classmethod(get_supported_backends),
                              "_get_aggregator": _get_aggregator, "_conclude": _conclude,
                              "__module__": cls.__module__})
    analysis.__class__ = sub
    return analysis


def cuda_backend(n_workers: int = 1):
    """A ``BackendBase`` executor for ``AnalysisBase.run(backend=..., unsupported_backend=True)``
    (``analysis/backends.py:43-157``, ``analysis/base.py:590-695``).

    ``apply(func, computations)`` receives ``func = partial(analysis._compute, ...)`` and one
    ``[(result index, frame number), ...]`` array per part (``np.array_split`` of the frames,
    ``base.py:579-588``); each part is computed by ``run_on_gpu`` on a shallow copy of the analysis
    object, which then carries ``frames``, ``times`` and ``results`` for the reference's own
    aggregation (``_get_aggregator``, ``ResultsGroup.merge``).  ``n_workers`` parts run at the same
    time, worker ``k`` on CUDA device ``k mod device_count`` (one host thread per worker: the CUDA
    runtime's current device is per thread, the C ABI releases the GIL, and no process is forked --
    a CUDA context does not survive ``fork``, which is why the reference's ``multiprocessing``
    backend cannot drive this library).  ``RMSD`` gets here as it is; for ``InterRDF``,
    ``InterRDF_s``, ``AlignTraj`` and ``Contacts`` wrap the object with ``gpu_enabled`` first.
    """
    import copy
    from concurrent.futures import ThreadPoolExecutor

    from MDAnalysis.analysis.backends import BackendBase
    from MDAnalysis.analysis.results import Results

    class CudaBackend(BackendBase):
        def apply(self, func, computations):
            import torch
            analysis = getattr(getattr(func, "func", func), "__self__", None)
            if analysis is None:
                raise ValueError("CudaBackend.apply: expected a bound AnalysisBase._compute")
            ndev = max(1, torch.cuda.device_count())
            keep = None
            if getattr(type(analysis), "_mdagpu_enabled", False):
                keep = set(_aggregators()[type(analysis).__name__])

            def work(job):
                k, indexed_frames = job
                part = copy.copy(analysis)
                part.results = Results()
                frames = np.asarray(indexed_frames)[:, 1] if len(indexed_frames) else np.empty(0, dtype=int)
                with torch.cuda.device(k % ndev):
                    run_on_gpu(part, frames=frames)
                if keep is not None:                       # raw, mergeable quantities only
                    if "volume_cum" in keep:
                        part.results["volume_cum"] = np.float64(part.volume_cum)
                    for key in [key for key in part.results if key not in keep]:
                        del part.results[key]
                return part

            jobs = list(enumerate(computations))
            workers = max(1, min(int(self.n_workers), len(jobs)))
            if workers == 1:
                return [work(j) for j in jobs]
            with ThreadPoolExecutor(max_workers=workers) as pool:
                return list(pool.map(work, jobs))

    return CudaBackend(n_workers=n_workers)
