"""Plug the CUDA backend into an installed MDAnalysis.

``MDAnalysis.lib.distances`` selects its low-level module with
``_distances[backend.lower()]`` and ``getattr(module, funcname)``
(``lib/distances.py:139-166``); adding one dictionary entry is therefore all it
takes to make ``backend='cuda'`` work for ``distance_array``,
``self_distance_array``, ``calc_bonds`` / ``calc_angles`` / ``calc_dihedrals``,
``transform_RtoS/StoR`` and ``apply_PBC`` -- no source change, AtomGroup and
AnalysisBase APIs untouched.  ``lib/_distopia.py:59-84`` is the precedent for a
backend module.

``capped_distance`` / ``self_capped_distance`` select by ``method=`` through
local dictionaries (``lib/distances.py:557-560, 1023-1026``), so they cannot be
extended from outside; ``install(patch_capped=True)`` swaps the two module-level
functions for wrappers that keep the reference signature and add ``backend=``.
"""
from __future__ import annotations


def install(patch_capped: bool = False, masquerade=None, capped_default: str = "serial",
            patch_classes: bool = False, patch_analysis: bool = False):
    """Register ``backend='cuda'`` with MDAnalysis.  Returns the patched module.

    masquerade      also install the module under existing key(s) (``'serial'``, or a list like
                    ``['serial', 'openmp']``) so that the reference's own parametrised tests
                    exercise the GPU (option a).
    capped_default  default ``backend=`` of the patched ``capped_distance`` wrappers.
    patch_classes   also point ``MDAnalysis.lib.nsgrid.FastNS`` and
                    ``MDAnalysis.lib.qcprot.CalcRMSDRotationalMatrix`` at the GPU twins (what the
                    reference's ``test_nsgrid.py`` / ``test_qcprot.py`` / ``analysis.rms`` call).
    patch_analysis  ``analysis.diffusionmap.DistanceMatrix.run`` computes the matrix with
                    ``run_on_gpu`` when the metric is ``rms.rmsd`` (the override INTEGRATION.md 3b
                    shows); other metrics keep the reference's Python loop.
    """
    import MDAnalysis.lib.distances as mda_distances

    from .lib import c_distances_cuda
    from .lib import distances as gpu_distances

    mda_distances._distances["cuda"] = c_distances_cuda
    for key in ([masquerade] if isinstance(masquerade, str) else (masquerade or [])):
        mda_distances._distances[key] = c_distances_cuda
    if patch_classes:
        import MDAnalysis.lib.nsgrid as mda_nsgrid
        import MDAnalysis.lib.qcprot as mda_qcprot
        from .lib import nsgrid as gpu_nsgrid
        from .lib import qcprot as gpu_qcprot
        mda_nsgrid.FastNS = gpu_nsgrid.FastNS
        mda_qcprot.CalcRMSDRotationalMatrix = gpu_qcprot.CalcRMSDRotationalMatrix
        mda_qcprot.InnerProduct = gpu_qcprot.InnerProduct
        mda_qcprot.FastCalcRMSDAndRotation = gpu_qcprot.FastCalcRMSDAndRotation
    if patch_analysis:
        import MDAnalysis.analysis.diffusionmap as mda_dmap
        from .adapters import run_on_gpu
        from .analysis.diffusionmap import _metric_kind
        ref_run = mda_dmap.DistanceMatrix.run
        if not getattr(ref_run, "_mdagpu", False):
            def run(self, start=None, stop=None, step=None, frames=None, verbose=None, **kwargs):
                try:
                    _metric_kind(self._metric)
                except NotImplementedError:
                    return ref_run(self, start=start, stop=stop, step=step, frames=frames, verbose=verbose, **kwargs)
                return run_on_gpu(self, start=start, stop=stop, step=step, frames=frames)
            run._mdagpu = True
            run.__doc__ = ref_run.__doc__
            mda_dmap.DistanceMatrix.run = run
    if patch_capped:
        ref_capped = mda_distances.capped_distance
        ref_self_capped = mda_distances.self_capped_distance

        def capped_distance(reference, configuration, max_cutoff, min_cutoff=None, box=None,
                            method=None, return_distances=True, backend=capped_default):
            if backend.lower() == "cuda":
                return gpu_distances.capped_distance(reference, configuration, max_cutoff,
                                                     min_cutoff=min_cutoff, box=box, method=method,
                                                     return_distances=return_distances)
            return ref_capped(reference, configuration, max_cutoff, min_cutoff=min_cutoff, box=box,
                              method=method, return_distances=return_distances)

        def self_capped_distance(reference, max_cutoff, min_cutoff=None, box=None, method=None,
                                 return_distances=True, backend=capped_default):
            if backend.lower() == "cuda":
                return gpu_distances.self_capped_distance(reference, max_cutoff,
                                                          min_cutoff=min_cutoff, box=box,
                                                          method=method,
                                                          return_distances=return_distances)
            return ref_self_capped(reference, max_cutoff, min_cutoff=min_cutoff, box=box,
                                   method=method, return_distances=return_distances)

        capped_distance.__doc__ = ref_capped.__doc__
        self_capped_distance.__doc__ = ref_self_capped.__doc__
        mda_distances.capped_distance = capped_distance
        mda_distances.self_capped_distance = self_capped_distance
    return mda_distances


def uninstall():
    import MDAnalysis.lib.distances as mda_distances
    mda_distances._distances.pop("cuda", None)
