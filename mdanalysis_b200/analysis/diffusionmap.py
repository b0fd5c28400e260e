"""Pairwise RMSD matrices on the GPU -- mirror of ``DistanceMatrix`` in
``MDAnalysis/analysis/diffusionmap.py`` (:147-289).

The reference fills ``results.dist_matrix`` with a Python loop over frame pairs
(``_single_frame``, :248-267), one ``metric(i_ref, j_ref, weights=...)`` call each, the metric
defaulting to ``rms.rmsd`` (no centring, no superposition).  Here the whole matrix is one launch
of ``mdagpu_pairwise_rmsd``.  Only the metrics the kernel implements are accepted:

    metric=rmsd (default)            sqrt(sum(w (a - b)^2) / N)
    metric=rmsd_superposition        rms.rmsd(a, b, weights, superposition=True)  (centre + QCP)

Any other callable raises ``NotImplementedError`` (there is no per-pair Python path in this
package).  ``select`` is an index array (or ``'all'`` / None); with real MDAnalysis AtomGroups a
selection string is resolved through ``select_atoms``.
"""
import functools

import numpy as np

from .._cabi import is_tensor
from ..frames import as_block
from ..lib import qcprot as qcp
from .base import BlockAnalysis
from .rms import _resolve_selection, rmsd


def rmsd_superposition(a, b, weights=None):
    """``rms.rmsd(a, b, weights=weights, superposition=True)``; as a ``metric`` it selects the QCP kernel."""
    return rmsd(a, b, weights=weights, superposition=True)


def _metric_kind(metric):
    if metric is rmsd or getattr(metric, "__name__", None) == "rmsd" and not isinstance(metric, functools.partial):
        return False
    if metric is rmsd_superposition:
        return True
    if isinstance(metric, functools.partial) and getattr(metric.func, "__name__", None) == "rmsd" and not metric.args:
        kw = dict(metric.keywords or {})
        sup = bool(kw.pop("superposition", False))
        center = bool(kw.pop("center", False))
        if not kw and (sup or not center):
            return sup
    raise NotImplementedError("the GPU DistanceMatrix implements metric=rmsd and metric=rmsd_superposition "
                              "(or functools.partial(rmsd, superposition=True)) only")


class DistanceMatrix(BlockAnalysis):
    """All-pairs RMSD matrix of a trajectory (diffusionmap.py:147-289).

    universe  FrameBlock | [F, n, 3] array / CUDA tensor | AtomGroup / Universe
    select    indices of the atoms entering the metric (default all)
    metric    ``rmsd`` (default) or ``rmsd_superposition``
    cutoff    values <= cutoff are stored as 0; the reference's default is the literal
              ``1E0-5`` (= -4.0, diffusionmap.py:237), kept here
    weights   None or an array over the selected atoms
    """

    def __init__(self, universe, select="all", metric=rmsd, cutoff=1E0 - 5, weights=None, **kwargs):
        group = universe.atoms if hasattr(universe, "atoms") and hasattr(universe, "trajectory") and \
            not hasattr(universe, "positions") else universe
        super().__init__(as_block(group), **kwargs)
        self._sel = _resolve_selection(group, select)
        self._superposition = _metric_kind(metric)
        self._cutoff = cutoff
        self._weights = weights
        self._calculated = False

    def _prepare(self, blk):
        self.results.dist_matrix = np.zeros((blk.n_frames, blk.n_frames))

    def _compute(self, blk):
        m = qcp.pairwise_rmsd(blk.positions, select=self._sel, weights=self._weights,
                              superposition=self._superposition, cutoff=self._cutoff)
        self.results.dist_matrix = m.detach().cpu().numpy() if is_tensor(m) else m

    def _conclude(self):
        self._calculated = True

    @property
    def dist_matrix(self):
        import warnings
        warnings.warn("The `dist_matrix` attribute was deprecated in MDAnalysis 2.0.0 and will be removed in "
                      "MDAnalysis 3.0.0. Please use `results.dist_matrix` instead.", DeprecationWarning)
        return self.results.dist_matrix
